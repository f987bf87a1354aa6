#!/usr/bin/env python
"""Headline benchmark: frames/s of association + triangulation + IK on synthetic 8-camera x 32-person
BODY_25 scenes (BASELINE.json metric), B independent clips per GPU advancing one frame per step.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                          # the CPU restatement of the reference

A "step" = every clip of the batch advances by one multi-view frame (all 8 views, all people):
association (affinity + ALS matcher + assignment), IK update of every matched track, triangulation + IK
birth of new tracks, lifecycle. value = clips * steps / device time, inputs resident in HBM. e2e = the same
through mvmc_clips_step_host (pinned host inputs copied H2D and result records copied D2H every step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s assoc+triangulate+IK (8 cams x 32 people x BODY_25)"
WORKLOAD = "synthetic 8 cams x 32 people BODY_25 clips, association + triangulation + IK, one frame per clip per step"
N_VIEWS, N_PEOPLE = 8, 32


def make_inputs(n_clips, n_frames, seed, distinct):
    from multiview_motion_capture_b200 import synthetic as S
    distinct = min(distinct, n_clips)
    base = [S.make_clip(N_VIEWS, N_PEOPLE, n_frames, seed=seed, clip_idx=i) for i in range(distinct)]
    idx = np.arange(n_clips) % distinct
    kps = np.stack([S.body25_to_coco(c["kps25"]) for c in base], 1)[:, idx]      # [F,B,C,P,17,3]
    n_pose = np.stack([c["n_pose"] for c in base], 1)[:, idx]
    K = np.stack([c["K"] for c in base])[idx]
    RT = np.stack([c["RT"] for c in base])[idx]
    return np.ascontiguousarray(kps), np.ascontiguousarray(n_pose.astype(np.int32)), K, RT, base


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        try:
            self.p.terminate()
            self.p.wait(timeout=5)
        except Exception:
            pass
        try:
            self.f.flush()
            self.f.seek(0)
            sm, mx, reasons = [], [], set()
            for line in self.f.read().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.f.name)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------------------
# CPU arm (oracle port), one clip per process, steady-state tracking frames
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, clip_idx, first_frame, n_steps, n_warm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mvmc_oracle as o
    from multiview_motion_capture_b200 import synthetic as S
    n_frames = first_frame + n_steps + n_warm + 1
    c = S.make_clip(N_VIEWS, N_PEOPLE, n_frames, seed=seed, clip_idx=clip_idx)
    kps = S.body25_to_coco(c["kps25"])
    trk = o.Tracker(o.projections(c["K"], c["RT"]), c["K"], c["RT"])
    # steady state: tracks warm-started from the generator's ground truth of the previous frame (births are
    # 10x more expensive and belong to the first frames of a clip, not to the steady state being measured)
    f0 = first_frame - 1
    for pi in range(N_PEOPLE):
        prm = o.PoseParam(c["gt_root"][f0, pi].copy(), c["gt_euler"][f0, pi].copy(), trk.skel.side_bone_lens * c["gt_scale"][pi])
        joints, _ = o.forward_kinematics(trk.skel, prm.root, prm.euler, prm.bone_lens)
        t = o.Track(pi, [f0], [prm], [joints], [[]], state=o.CONFIRMED, hits=3)
        trk.tracks.append(t)
    trk.next_id = N_PEOPLE
    times = []
    for s in range(n_warm + n_steps):
        f = first_frame + s
        t0 = time.perf_counter()
        trk.step(f, kps[f], c["n_pose"][f])
        times.append(time.perf_counter() - t0)
    return times[n_warm:], len(trk.tracks)


def cpu_arm(n_steps, n_warm, cores, seed=1000):
    """Returns (frames_per_s, wall_s_per_step list, cores). Each step: `cores` clips advance one frame in parallel."""
    import multiprocessing as mp
    env_keys = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")
    old = {k: os.environ.get(k) for k in env_keys}
    for k in env_keys:
        os.environ[k] = "1"
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(seed, 100000 + i, 3, n_steps, n_warm) for i in range(cores)])
            wall = time.perf_counter() - t0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    per_clip = np.array([r[0] for r in res])          # [cores, n_steps] seconds per frame
    step_wall = per_clip.max(axis=0)                  # a step ends when the slowest clip has finished its frame
    fps = cores * n_steps / float(step_wall.sum())
    return fps, step_wall.tolist(), float(per_clip.mean()), wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = args.cpu_cores or os.cpu_count() or 1
    fps, step_wall, mean_frame_s, wall = cpu_arm(args.steps, min(args.warmup, 1), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(step_wall)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": cores, "n_views": N_VIEWS, "n_people": N_PEOPLE,
                   "state": "steady-state tracking frames, tracks warm-started from ground truth"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} clips x {args.steps} frames, one process per core, OMP_NUM_THREADS=1, "
                                   f"mean {mean_frame_s:.2f} s per clip-frame (oracle/mvmc_oracle.py, bit-identical to the "
                                   f"reference on its Shelf fixture)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from multiview_motion_capture_b200 import _lib
    from multiview_motion_capture_b200.clips import ClipBatch, ClipStreams
    from multiview_motion_capture_b200._lib import STEP_OUT_DTYPE, check, ptr
    lib = _lib.get_lib()

    B, K, W = args.clips, args.steps, args.warmup + args.preroll
    n_frames = W + K + 1
    t_gen = time.time()
    kps, n_pose, Kc, RT, _ = make_inputs(B, n_frames, seed=1000 + 7919 * rank, distinct=args.distinct)
    t_gen = time.time() - t_gen
    # the batch runs as `groups` groups of clips on their own streams (ClipStreams): one group's IK / affinity / copies fill
    # the SMs another group's ALS launch leaves idle while it drains
    cs = ClipStreams(B, N_VIEWS, N_PEOPLE, groups=args.groups, max_tracks=args.max_tracks, max_new=N_PEOPLE, device=dev)
    cs.set_calib(Kc, RT)
    kps_pin = torch.from_numpy(kps).pin_memory()
    np_pin = torch.from_numpy(n_pose).pin_memory()
    kps_dev = kps_pin.to(dev, non_blocking=True)
    np_dev = np_pin.to(dev, non_blocking=True)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident timing ----------------
    cs.fork()
    for s in range(W):
        cs.step_device(kps_dev[1 + s], np_dev[1 + s], 1 + s)
    cs.join()
    cs.stats(reset=True)
    launches0 = lib.mvmc_launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    cs.fork()
    for s in range(K):
        cs.step_device(kps_dev[1 + W + s], np_dev[1 + W + s], 1 + W + s)
    cs.join()
    summary = None
    if world > 1:
        # the only collective of the path: gather per-rank result summaries (north_star: NCCL only to gather)
        rec = torch.frombuffer(bytearray(8), dtype=torch.float64).to(dev)
        rec[0] = float(B * K)
        gathered = [torch.zeros_like(rec) for _ in range(world)]
        dist.all_gather(gathered, rec)
        summary = gathered
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = lib.mvmc_launch_count() - launches0
    st_run = cs.stats(reset=True)
    value = world * B * K / (ms * 1e-3)

    # ---------------- the kernels timed alone (roofline): ONE launch of all B clips per stage, nothing overlapped ----------------
    # (in the throughput run above kernels of different groups share the SMs, so per-kernel event times there say nothing
    #  about a kernel; this is the same kernel on the same clips, two steps after the same warm-up)
    cb = ClipBatch(B, N_VIEWS, N_PEOPLE, max_tracks=args.max_tracks, max_new=N_PEOPLE, device=dev)
    cb.set_calib(Kc, RT)
    for s in range(W):
        cb.step_device(kps_dev[1 + s], np_dev[1 + s], 1 + s)
    cb.stats(reset=True)
    cb.profile(1)
    if args.als_phases:
        check(lib.mvmc_als_phase_profile(1, None), "mvmc_als_phase_profile")
    KP = min(K, 2)
    for s in range(KP):
        cb.step_device(kps_dev[1 + W + s], np_dev[1 + W + s], 1 + W + s)
    torch.cuda.synchronize(dev)
    if args.als_phases and rank == 0:
        ph = np.zeros(20)
        check(lib.mvmc_als_phase_profile(0, ptr(ph)), "mvmc_als_phase_profile")
        names = ["G=AtA", "inv1", "T=AtXt", "B", "H=BtB", "inv2", "T=BtXtt", "A", "X=ABt", "reduce", "mu-pass", "init", "admm-pass", "admm-wait", "admm-fence", "gj-load", "gj-inv8", "gj-panel", "gj-update", "gj-store"]
        print("k_als phase share of CTA cycles: " + ", ".join(f"{n} {100 * v / ph.sum():.1f}%" for n, v in zip(names, ph)),
              file=sys.stderr, flush=True)
    stage_ms, n_prof = cb.profile(0)
    st = cb.stats(reset=True)
    single_bytes = cb.device_bytes
    cb.close()
    del cb

    # ---------------- FP64 peak probes (CUDA-core DFMA and tensor-core DMMA), timed in this run ----------------
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    blocks = 148 * 8

    def probe(fn, iters, flops_per_iter_block):
        check(fn(blocks, 1000, ptr(sink), stream.cuda_stream), "probe")
        torch.cuda.synchronize(dev)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        check(fn(blocks, iters, ptr(sink), stream.cuda_stream), "probe")
        p1.record(stream)
        torch.cuda.synchronize(dev)
        return blocks * iters * flops_per_iter_block / (p0.elapsed_time(p1) * 1e-3) / 1e12

    fp64_dfma_tflops = probe(lib.mvmc_fp64_probe, 200000, 256 * 16.0)
    fp64_dmma_tflops = probe(lib.mvmc_fp64_tensor_probe, 100000, 8 * 8 * 512.0)

    # ---------------- end-to-end through the host-buffer C-ABI call ----------------
    e2e = None
    if not args.no_e2e:
        cs.reset()
        out_pin = torch.empty(B * STEP_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()

        def host_step(fi):   # every group: H2D of its slice, the step, D2H of its records (mvmc_clips_step_host_async), then sync
            cs.step_host(kps_pin[fi], np_pin[fi], fi, out_pin)
        for s in range(W):
            host_step(1 + s)
        barrier()
        t0 = time.perf_counter()
        for s in range(K):
            host_step(1 + W + s)
        torch.cuda.synchronize(dev)
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        rec = out_pin.numpy().view(STEP_OUT_DTYPE)
        e2e = {"value": world * B * K / float(t_e2e.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(kps_pin[0].numel() * 8 + np_pin[0].numel() * 4),
               "d2h_bytes_per_step": int(out_pin.numel()),
               "alive_tracks_per_clip": float(rec["n_alive"].mean()), "capacity_errors": int((rec["error"] != 0).sum())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    n_prof = max(n_prof, 1)
    als_ms = stage_ms["als"] / n_prof
    ik_ms = stage_ms["ik"] / n_prof
    dom = "k_als" if als_ms >= ik_ms else "k_ik_solve"
    dom_ms = max(als_ms, ik_ms)
    dom_flops = (st["als_flops"] if dom == "k_als" else st["ik_flops"]) / KP
    # k_als runs on the FP64 tensor cores (DMMA), k_ik_solve on the FP64 CUDA cores: each against its own probed peak
    peak_tf = fp64_dmma_tflops if dom == "k_als" else fp64_dfma_tflops
    # algorithmic bytes per launch: inputs (BODY_25 detections as float64 COCO) + outputs (params + joints + assignments)
    bytes_per_clip_frame = N_VIEWS * N_PEOPLE * 17 * 3 * 8 + N_PEOPLE * (68 + 54) * 8 + (N_VIEWS + 1) * N_PEOPLE * 4
    alg_bytes = B * bytes_per_clip_frame
    achieved_tf = dom_flops / (dom_ms * 1e-3) / 1e12
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_summary.json)
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            ns = json.load(f)[dom]
        traffic = {"dram_bytes_per_launch": ns["dram_bytes"], "clips_in_capture": ns["clips"],
                   "dram_bytes_per_clip_frame": ns["dram_bytes"] / ns["clips"], "source": ns["source"]}
    except Exception:
        pass
    roofline = {
        "kernel": dom, "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
        "peak_source": "FP64 tensor-core (mma.sync m8n8k4 f64 = DMMA) probe kernel timed in this run; MEASURED_PEAKS.json holds "
                       "only HBM and bf16 peaks and this path is FP64 (SURVEY.md 8d). FP64 CUDA-core (DFMA) probe: "
                       f"{fp64_dfma_tflops:.1f} TFLOP/s",
        "kernel_ms_per_launch": dom_ms, "algorithmic_flops_per_launch": dom_flops,
        "hbm": {"achieved": alg_bytes / (ms / K * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (ms / K * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_step": alg_bytes,
                "note": "whole step, ALGORITHMIC bytes (inputs + results): about 25 kFLOP per algorithmic byte, so the roofline "
                        "that binds is the FP64 tensor pipe; the kernel's real DRAM traffic is in kernel_dram"},
        "kernel_dram": None if traffic is None else {
            "achieved": traffic["dram_bytes_per_clip_frame"] * B / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": traffic["dram_bytes_per_clip_frame"] * B / (dom_ms * 1e-3) / 1e9 / hbm_peak,
            "note": "DRAM bytes of the dominant kernel per clip-frame from the committed ncu capture x clips of this run / its "
                    "event-timed duration: the second bound of k_als (its n x n iterates stream through HBM every iteration)"},
        "stage_ms_per_step": {k: v / n_prof for k, v in stage_ms.items()},
        "timed": f"each kernel alone: one launch of all {B} clips per stage, {KP} steps after the same warm-up (the throughput run "
                 f"overlaps {args.groups} clip groups on streams)",
        "ik": {"achieved": st["ik_flops"] / KP / (ik_ms * 1e-3) / 1e12 if ik_ms > 0 else None, "peak": fp64_dfma_tflops,
               "unit": "TFLOP/s", "note": "k_ik_solve, SURVEY.md 8d flop formula on the run's own (nfev, njev) counts"},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = args.cpu_cores or os.cpu_count() or 1
        try:
            fps, step_wall, mean_frame_s, wall = cpu_arm(3, 0, cores)
            cpu_baseline = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                            "sample": f"{cores} clips x 3 steady-state frames of the same 8x32 workload, one process per core, "
                                      f"OMP_NUM_THREADS=1, mean {mean_frame_s:.2f} s per clip-frame, wall {wall:.0f} s"}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": B, "distinct_clips": min(args.distinct, B), "n_views": N_VIEWS,
                   "n_people": N_PEOPLE, "max_tracks": args.max_tracks, "preroll_frames": args.preroll,
                   "clip_groups_on_streams": args.groups,
                   "l2": "each step reads a new frame for every clip and sweeps the per-clip ALS workspaces "
                         f"({cs.device_bytes / 2**20:.0f} MiB on device, far larger than the 126 MB L2)",
                   "als_iters_per_clip_frame": st_run["als_iters"] / max(st_run["clip_frames"], 1),
                   "ik_solves_per_clip_frame": st_run["ik_solves"] / max(st_run["clip_frames"], 1) / 2,
                   "input_gen_s": t_gen},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=1184, help="clips per GPU (8 per SM = 4 waves of resident k_als CTAs)")
    ap.add_argument("--distinct", type=int, default=148, help="distinct synthetic clips generated per rank (tiled to --clips)")
    ap.add_argument("--max-tracks", type=int, default=40)
    ap.add_argument("--preroll", type=int, default=4, help="untimed frames before the warm-up (track births happen here)")
    ap.add_argument("--groups", type=int, default=3, help="clip groups stepped concurrently on their own CUDA streams")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--als-phases", action="store_true", help="print k_als's per-phase cycle shares to stderr (diagnostic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--views", type=int, default=0, help="other scene shapes (not the headline metric): cameras ...")
    ap.add_argument("--people", type=int, default=0, help="... and people per scene, e.g. --views 5 --people 4 (Shelf shaped)")
    args = ap.parse_args()
    if args.views or args.people:
        global N_VIEWS, N_PEOPLE, WORKLOAD
        N_VIEWS, N_PEOPLE = args.views or N_VIEWS, args.people or N_PEOPLE
        WORKLOAD = (f"synthetic {N_VIEWS} cams x {N_PEOPLE} people BODY_25 clips (NOT the headline shape), association + "
                    f"triangulation + IK, one frame per clip per step")
        args.max_tracks = min(args.max_tracks, max(8, N_PEOPLE + N_PEOPLE // 4))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
