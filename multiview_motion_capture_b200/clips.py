"""Clip-batch pipeline handle (mvmc_clips): B independent clips advance one frame per step.
Mirrors MvTracker.update_4d (reference: src/motion_capture.py:873-963) for a batch of clips."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import Config, N_COCO, STEP_OUT_DTYPE, check, ptr


class ClipBatch:
    def __init__(self, n_clips, n_views, max_poses, max_tracks=None, max_new=None, device=None, n_inits=3, max_age=0,
                 nfev_update=5, nfev_birth=50):
        lib = _lib.get_lib()
        self.lib = lib
        self.device = torch.device(device) if device is not None else _lib.default_device()
        if self.device.type == "cuda":
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
            torch.cuda.set_device(self.device)
        cfg = Config()
        lib.mvmc_default_config(ctypes.byref(cfg))
        cfg.n_clips, cfg.n_views, cfg.max_poses = n_clips, n_views, max_poses
        cfg.max_tracks = max_tracks if max_tracks is not None else min(64, 2 * max_poses)
        cfg.max_new = max_new if max_new is not None else min(64, max(4, 2 * max_poses))
        cfg.n_inits, cfg.max_age, cfg.nfev_update, cfg.nfev_birth = n_inits, max_age, nfev_update, nfev_birth
        self.cfg = cfg
        self.B, self.C, self.Pmax, self.Tmax = n_clips, n_views, max_poses, cfg.max_tracks
        self._h = ctypes.c_void_p()
        check(lib.mvmc_clips_create(ctypes.byref(cfg), ctypes.byref(self._h)), "mvmc_clips_create")
        self._pinned = self.device.type == "cuda"
        self._out_host = self._host_buffer(n_clips * STEP_OUT_DTYPE.itemsize, np.uint8)
        self._kps_host = self._host_buffer(n_clips * n_views * max_poses * N_COCO * 3, np.float64)
        self._np_host = self._host_buffer(n_clips * n_views, np.int32)

    def _host_buffer(self, count, dtype):
        tdt = {np.uint8: torch.uint8, np.float64: torch.float64, np.int32: torch.int32}[dtype]
        t = torch.empty(count, dtype=tdt)
        if self._pinned:
            t = t.pin_memory()
        return t

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    @property
    def device_bytes(self):
        return self.lib.mvmc_clips_device_bytes(self._h)

    def close(self):
        if self._h:
            self.lib.mvmc_clips_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_calib(self, K, Rt, P=None):
        """K [B,C,3,3], Rt [B,C,3,4] (numpy, float64); P defaults to K @ Rt computed with NumPy exactly as
        load_calib does (src/motion_capture.py:250-272)."""
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(self.B, self.C, 3, 3)
        Rt = np.ascontiguousarray(Rt, dtype=np.float64).reshape(self.B, self.C, 3, 4)
        if P is None:
            P = np.stack([np.stack([K[b, c] @ Rt[b, c] for c in range(self.C)]) for b in range(self.B)])
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(self.B, self.C, 3, 4)
        check(self.lib.mvmc_clips_set_calib(self._h, ptr(K), ptr(Rt), ptr(P), self._stream()), "mvmc_clips_set_calib")
        self.sync()

    def reset(self):
        check(self.lib.mvmc_clips_reset(self._h, self._stream()), "mvmc_clips_reset")

    def sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def step_device(self, kps, n_pose, frame_idx):
        """Device-resident step: kps [B,C,Pmax,17,3] f64, n_pose [B,C] i32 torch tensors on self.device. Async."""
        assert kps.dtype == torch.float64 and n_pose.dtype == torch.int32 and kps.is_contiguous() and n_pose.is_contiguous()
        ops = _lib.torch_ops() if kps.is_cuda else None
        if ops is not None:     # the PyTorch C++ extension over the C-ABI (csrc/torch_ext.cpp)
            ops.clips_step(int(self._h.value), kps, n_pose, int(frame_idx))
        else:
            check(self.lib.mvmc_clips_step(self._h, ptr(kps), ptr(n_pose), int(frame_idx), self._stream()), "mvmc_clips_step")

    def last_out_device(self):
        """uint8 view [B, sizeof(mvmc_step_out)] of the device records of the last step (no copy)."""
        addr = self.lib.mvmc_clips_last_out(self._h)
        return addr

    def step(self, kps, n_pose, frame_idx, want_out=True):
        """Host-buffer step (H2D copy, all kernels, D2H copy of the records, sync). Returns a NumPy structured
        array [B] of STEP_OUT_DTYPE (a view on a reused pinned buffer; copy it to keep it)."""
        kh = self._kps_host.numpy()
        kh[:] = np.asarray(kps, dtype=np.float64).reshape(-1)
        nh = self._np_host.numpy()
        nh[:] = np.asarray(n_pose, dtype=np.int32).reshape(-1)
        return self.step_pinned(frame_idx, want_out)

    def step_pinned(self, frame_idx, want_out=True):
        """Same, taking the inputs already written into self.kps_host / self.n_pose_host."""
        outp = ptr(self._out_host) if want_out else None
        check(self.lib.mvmc_clips_step_host(self._h, ptr(self._kps_host), ptr(self._np_host), int(frame_idx), outp,
                                            self._stream()), "mvmc_clips_step_host")
        if not want_out:
            return None
        rec = self._out_host.numpy().view(STEP_OUT_DTYPE)
        if (rec["error"] != 0).any():
            bad = np.nonzero(rec["error"])[0]
            raise _lib.MvmcError(f"capacity exceeded in clips {bad[:8].tolist()} (raise max_tracks / max_new)")
        return rec

    def step_body25(self, kps25, n_people, frame_idx):
        """Ingest + step (mvmc_clips_step_body25_host): raw OpenPose BODY_25 detections kps25 [B,C,Pmax,25,3] and people
        counts [B,C]; the BODY_25 -> COCO gather runs on the device. Returns the records like step()."""
        if not hasattr(self, "_kps25_host"):
            self._kps25_host = self._host_buffer(self.B * self.C * self.Pmax * 75, np.float64)
        self._kps25_host.numpy()[:] = np.asarray(kps25, dtype=np.float64).reshape(-1)
        self._np_host.numpy()[:] = np.asarray(n_people, dtype=np.int32).reshape(-1)
        check(self.lib.mvmc_clips_step_body25_host(self._h, ptr(self._kps25_host), ptr(self._np_host), int(frame_idx),
                                                   ptr(self._out_host), self._stream()), "mvmc_clips_step_body25_host")
        self.sync()
        rec = self._out_host.numpy().view(STEP_OUT_DTYPE)
        if (rec["error"] != 0).any():
            raise _lib.MvmcError(f"capacity exceeded in clips {np.nonzero(rec['error'])[0][:8].tolist()} (raise max_tracks / max_new)")
        return rec

    @property
    def kps_host(self):
        return self._kps_host.numpy().reshape(self.B, self.C, self.Pmax, N_COCO, 3)

    @property
    def n_pose_host(self):
        return self._np_host.numpy().reshape(self.B, self.C)

    def pack_records(self, cap, clip0=0, rec=None, count=None):
        """Compact records of the tracks solved in the last step (mvmc_clips_pack_records): rec [B,cap,128] f64 and count [B]
        i32 device tensors (allocated when not given). Asynchronous on the current stream."""
        if rec is None:
            rec = torch.empty((self.B, cap, 128), dtype=torch.float64, device=self.device)
            count = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        ops = _lib.torch_ops() if rec.is_cuda else None
        if ops is not None:
            ops.clips_pack_records(int(self._h.value), int(cap), int(clip0), rec, count)
        else:
            check(self.lib.mvmc_clips_pack_records(self._h, int(cap), int(clip0), ptr(rec), ptr(count), self._stream()),
                  "mvmc_clips_pack_records")
        return rec, count

    def set_tracks(self, n_trk, ids, state, hits, tsu, length, param, joints, next_id):
        a = lambda x, dt, shape: np.ascontiguousarray(np.asarray(x, dtype=dt).reshape(shape))
        B, T = self.B, self.Tmax
        args = [a(n_trk, np.int32, (B,)), a(ids, np.int32, (B, T)), a(state, np.int32, (B, T)), a(hits, np.int32, (B, T)),
                a(tsu, np.int32, (B, T)), a(length, np.int32, (B, T)), a(param, np.float64, (B, T, 68)),
                a(joints, np.float64, (B, T, 54)), a(next_id, np.int32, (B,))]
        check(self.lib.mvmc_clips_set_tracks_host(self._h, *[ptr(x) for x in args], self._stream()),
              "mvmc_clips_set_tracks_host")

    STAT_NAMES = ("als_flops", "als_iters", "clip_frames", "ik_solves", "nfev", "njev", "ik_flops", "sum_n2")
    STAGE_NAMES = ("affinity", "als", "assign", "ik", "commit")

    def stats(self, reset=False):
        out = np.zeros(8)
        check(self.lib.mvmc_clips_stats_host(self._h, ptr(out), int(reset), self._stream()), "mvmc_clips_stats_host")
        return dict(zip(self.STAT_NAMES, out.tolist()))

    def profile(self, enable):
        """enable=1 start recording stage events; 0 stop + read; -1 read. Returns (dict stage->ms, n_steps)."""
        out = np.zeros(5)
        n = ctypes.c_int(0)
        check(self.lib.mvmc_clips_profile(self._h, int(enable), ptr(out), ctypes.addressof(n), self._stream()),
              "mvmc_clips_profile")
        return dict(zip(self.STAGE_NAMES, out.tolist())), n.value

    def read_big_groups(self, b=0):
        """The many-pose birth groups of clip b in the last step: {birth index k: [(view, pose id), ...]} (groups of more than
        MVMC_MAX_SEL poses; k = position among the tracks born in that step)."""
        from ._lib import MAX_BIG, MAX_GROUP
        n = ctypes.c_int(0)
        nsel = np.zeros(MAX_BIG, np.int32)
        slot = np.zeros(MAX_BIG, np.int32)
        sel = np.zeros((MAX_BIG, MAX_GROUP, 2), np.int32)
        check(self.lib.mvmc_clips_read_big_groups_host(self._h, int(b), ctypes.addressof(n), ptr(nsel), ptr(slot), ptr(sel), self._stream()),
              "mvmc_clips_read_big_groups_host")
        return {int(slot[g]): [tuple(x) for x in sel[g, :nsel[g]].tolist()] for g in range(n.value)}

    def read_matrices(self, b=0):
        """(dst, sim, xbin, dim_groups) of clip b from the last step."""
        N = self.Tmax + self.C * self.Pmax
        dst = np.zeros((N, N))
        sim = np.zeros((N, N))
        xb = np.zeros((N, N), dtype=np.uint8)
        n = ctypes.c_int(0)
        dg = np.zeros(self.C + 2, dtype=np.int32)
        check(self.lib.mvmc_clips_read_matrices_host(self._h, b, ptr(dst), ptr(sim), ptr(xb), ctypes.addressof(n), ptr(dg),
                                                     self._stream()), "mvmc_clips_read_matrices_host")
        k = n.value
        return (dst.reshape(-1)[:k * k].reshape(k, k).copy(), sim.reshape(-1)[:k * k].reshape(k, k).copy(),
                xb.reshape(-1)[:k * k].reshape(k, k).astype(bool), dg)



class ClipStreams:
    """The same batch of clips as G groups, each a ClipBatch on its own CUDA stream.

    Clips share nothing, so groups can run concurrently: while one group's ALS launch drains (its last CTAs leave SMs
    idle) another group's IK / affinity kernels and host<->device copies fill the machine. Measured on a B200 at 8x32:
    +4 % with two groups, +5 % with three (tools/groups_probe.py). Results are those of ClipBatch, clip for clip."""

    def __init__(self, n_clips, n_views, max_poses, groups=3, device=None, **kw):
        groups = max(1, min(int(groups), n_clips))
        self.B, self.C, self.Pmax = n_clips, n_views, max_poses
        sizes = [n_clips // groups + (1 if g < n_clips % groups else 0) for g in range(groups)]
        self.bounds = [0]
        for sz in sizes:
            self.bounds.append(self.bounds[-1] + sz)
        self.batches = [ClipBatch(sz, n_views, max_poses, device=device, **kw) for sz in sizes]
        self.device = self.batches[0].device
        self.lib = self.batches[0].lib
        self.streams = [torch.cuda.Stream(self.device) for _ in sizes] if self.device.type == "cuda" else [None] * groups
        self.Tmax = self.batches[0].Tmax

    def _each(self):
        for g, cb in enumerate(self.batches):
            yield cb, self.bounds[g], self.bounds[g + 1], self.streams[g]

    @property
    def device_bytes(self):
        return sum(cb.device_bytes for cb in self.batches)

    def close(self):
        for cb in self.batches:
            cb.close()

    def set_calib(self, K, Rt):
        K = np.asarray(K, dtype=np.float64).reshape(self.B, self.C, 3, 3)
        Rt = np.asarray(Rt, dtype=np.float64).reshape(self.B, self.C, 3, 4)
        for cb, lo, hi, _ in self._each():
            cb.set_calib(K[lo:hi], Rt[lo:hi])

    def reset(self):
        for cb in self.batches:
            cb.reset()
        self.sync()

    def fork(self):
        """Group streams wait for the work queued so far on the current stream."""
        if self.device.type == "cuda":
            cur = torch.cuda.current_stream(self.device)
            for st in self.streams:
                st.wait_stream(cur)

    def join(self):
        """The current stream waits for every group."""
        if self.device.type == "cuda":
            cur = torch.cuda.current_stream(self.device)
            for st in self.streams:
                cur.wait_stream(st)

    def sync(self):
        if self.device.type == "cuda":
            for st in self.streams:
                st.synchronize()

    def step_device(self, kps, n_pose, frame_idx):
        """kps [B,C,Pmax,17,3] f64, n_pose [B,C] i32 on the device; asynchronous (bracket a run with fork() / join())."""
        for cb, lo, hi, st in self._each():
            if st is None:
                cb.step_device(kps[lo:hi], n_pose[lo:hi], frame_idx)
            else:
                with torch.cuda.stream(st):
                    cb.step_device(kps[lo:hi], n_pose[lo:hi], frame_idx)

    def step_host(self, kps_pinned, n_pose_pinned, frame_idx, out_pinned):
        """Pinned host tensors kps [B,...] f64, n_pose [B,C] i32, out uint8 [B * sizeof(mvmc_step_out)]: every group copies
        its slice in, steps, copies its records out - all asynchronous on the group's stream; returns after all did."""
        rec = STEP_OUT_DTYPE.itemsize
        for cb, lo, hi, st in self._each():
            check(self.lib.mvmc_clips_step_host_async(cb._h, ptr(kps_pinned[lo:hi]), ptr(n_pose_pinned[lo:hi]), int(frame_idx),
                                                      ptr(out_pinned[lo * rec:hi * rec]),
                                                      st.cuda_stream if st is not None else None), "mvmc_clips_step_host_async")
        self.sync()

    def set_tracks(self, n_trk, ids, state, hits, tsu, length, param, joints, next_id):
        B, T = self.B, self.Tmax
        r = lambda x, dt, shape: np.asarray(x, dtype=dt).reshape(shape)
        n_trk, next_id = r(n_trk, np.int32, (B,)), r(next_id, np.int32, (B,))
        ids, state, hits, tsu, length = (r(x, np.int32, (B, T)) for x in (ids, state, hits, tsu, length))
        param, joints = r(param, np.float64, (B, T, 68)), r(joints, np.float64, (B, T, 54))
        for cb, lo, hi, st in self._each():
            args = (n_trk[lo:hi], ids[lo:hi], state[lo:hi], hits[lo:hi], tsu[lo:hi], length[lo:hi], param[lo:hi], joints[lo:hi],
                    next_id[lo:hi])
            if st is None:
                cb.set_tracks(*args)
            else:
                with torch.cuda.stream(st):
                    cb.set_tracks(*args)
        self.sync()

    def stats(self, reset=False):
        tot = None
        for cb in self.batches:
            s = cb.stats(reset)
            tot = s if tot is None else {k: tot[k] + v for k, v in s.items()}
        return tot
