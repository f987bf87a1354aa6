"""Runs the Shelf fixture through the clip pipeline (GPU, or the emulator with --emu) free-running and
teacher-forced, and compares every frame with the reference goldens. Prints one line per frame + summary.
Diagnostic tool (the pytest versions of these checks live in tests/)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--forced", action="store_true", help="teacher-force the track table from the goldens")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--replicas", type=int, default=1)
    args = ap.parse_args()
    from multiview_motion_capture_b200 import _lib
    if args.emu:
        _lib.use_library(os.path.join(ROOT, "tests", "emu", "libmvmc_emu.so"), device="cpu")
    from multiview_motion_capture_b200.clips import ClipBatch
    import mvmc_oracle as o
    inp = np.load(os.path.join(ROOT, "tests", "golden", "shelf_inputs.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "shelf_ref.npz"))
    kps = o.body25_to_coco(inp["kps25"])
    C, Pm = kps.shape[1:3]
    B = args.replicas
    Tmax = 24
    cb = ClipBatch(B, C, 8, max_tracks=Tmax, max_new=8)
    cb.set_calib(np.repeat(inp["K"][None], B, 0), np.repeat(inp["RT"][None], B, 0))
    kp = np.zeros((B, C, 8, 17, 3))
    last = min(args.frames, int(g["last_frame"]))
    # golden track table replay (id -> last param/joints/state)
    table = {}
    stats = dict(frames=0, xbin_same=0, iters_same=0, alive_same=0, upd_same=0, max_dst=0.0, max_sim=0.0, dj=[], dcost=[],
                 repl_same=0)
    t_all = time.time()
    for f in range(1, last + 1):
        k = f"f{f:04d}_"
        kp[:] = 0
        kp[:, :, :Pm] = kps[f][None]
        if args.forced:
            ids = g[k + "alive_before"].tolist()
            n = len(ids)
            a_ids = np.zeros((B, Tmax), np.int32); st = np.zeros((B, Tmax), np.int32); hits = np.zeros((B, Tmax), np.int32)
            tsu = np.zeros((B, Tmax), np.int32); ln = np.zeros((B, Tmax), np.int32)
            prm = np.zeros((B, Tmax, 68)); jn = np.zeros((B, Tmax, 54))
            for i, tid in enumerate(ids):
                e = table[tid]
                a_ids[:, i] = tid; st[:, i] = e["state"]; hits[:, i] = e["hits"]; tsu[:, i] = e["tsu"]; ln[:, i] = e["len"]
                prm[:, i] = e["param"]; jn[:, i] = e["joints"]
            nid = (max(table) + 1) if table else 0
            cb.set_tracks(np.full(B, n), a_ids, st, hits, tsu, ln, prm, jn, np.full(B, nid))
        t0 = time.time()
        recs = cb.step(kp, np.repeat(inp["n_pose"][f][None], B, 0), f)
        dt = time.time() - t0
        rec = recs[0].copy()
        dst, sim, xb, dg = cb.read_matrices(0)
        gd, gs = g[k + "dst"], g[k + "sim"]
        same_shape = dst.shape == gd.shape
        dd = float(np.abs(dst - gd).max()) if same_shape else float("inf")
        ds = float(np.abs(sim - gs).max()) if same_shape else float("inf")
        xbd = int((xb != g[k + "xbin"].astype(bool)).sum()) if same_shape else -1
        n_alive = int(rec["n_alive"])
        tr = rec["tracks"][:n_alive]
        ids_now = tr["track_id"].tolist()
        upd = tr[tr["updated"] > 0]
        same_alive = ids_now == g[k + "alive_after"].tolist()
        same_upd = upd["track_id"].tolist() == g[k + "upd_ids"].tolist()
        dj = de = -1.0
        if same_upd and len(upd):
            djs = np.abs(upd["joints"].reshape(-1, 18, 3) - g[k + "upd_joints"]).max(axis=(1, 2))
            dj = float(djs.max())
            stats["dj"].extend(djs.tolist())
        if B > 1:
            same_rep = all((recs[b]["tracks"]["joints"][:n_alive] == recs[0]["tracks"]["joints"][:n_alive]).all() and
                           recs[b]["n_alive"] == n_alive for b in range(1, B))
            stats["repl_same"] += int(same_rep)
        stats["frames"] += 1
        stats["xbin_same"] += int(xbd == 0)
        stats["iters_same"] += int(int(rec["als_iters"]) == int(g[k + "als_iters"]))
        stats["alive_same"] += int(same_alive)
        stats["upd_same"] += int(same_upd)
        if args.forced or f == 1:
            stats["max_dst"] = max(stats["max_dst"], dd)
            stats["max_sim"] = max(stats["max_sim"], ds)
        if not args.quiet:
            print(f"f{f} n={rec['n_total']} dst={dd:.1e} sim={ds:.1e} xbin={xbd} it={rec['als_iters']}/{int(g[k+'als_iters'])} "
                  f"alive={ids_now} ok={same_alive} upd_ok={same_upd} dj={dj:.1e} nfev={upd['nfev'].tolist()} {dt*1e3:.1f}ms",
                  flush=True)
        # advance the golden table
        if args.forced:
            for tid in list(table):
                table[tid]["tsu"] += 1
            a_after = g[k + "alive_after"].tolist()
            a_state = g[k + "alive_state"]
            uids = g[k + "upd_ids"].tolist()
            for i, tid in enumerate(uids):
                prm = np.concatenate([g[k + "upd_root"][i], g[k + "upd_euler"][i].reshape(-1), g[k + "upd_blens"][i]])
                table.setdefault(tid, {})
                table[tid].update(param=prm, joints=g[k + "upd_joints"][i].reshape(-1))
            for i, tid in enumerate(a_after):
                table[tid].update(state=int(a_state[i, 0]), hits=int(a_state[i, 1]), tsu=int(a_state[i, 2]),
                                  len=int(a_state[i, 3]))
            for tid in list(table):
                if tid not in a_after:
                    del table[tid]
    n = stats["frames"]
    dj = np.array(stats["dj"]) if stats["dj"] else np.zeros(1)
    print(f"SUMMARY mode={'forced' if args.forced else 'free'} frames={n} xbin_same={stats['xbin_same']} "
          f"iters_same={stats['iters_same']} alive_same={stats['alive_same']} upd_same={stats['upd_same']} "
          f"max_dst={stats['max_dst']:.2e} max_sim={stats['max_sim']:.2e} joints_diff_m: median={np.median(dj):.2e} "
          f"p90={np.percentile(dj, 90):.2e} max={dj.max():.2e} replicas_identical={stats['repl_same']}/{n if B > 1 else 0} "
          f"wall={time.time()-t_all:.1f}s launches={cb.lib.mvmc_launch_count()}")


if __name__ == "__main__":
    main()
