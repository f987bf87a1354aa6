"""TEST INFRASTRUCTURE — generates tests/golden/* by running the REAL reference here.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py shelf   [--frames 300]
    python oracle/make_golden.py synth   (after the synthetic generator exists)

Outputs (committed, small):
  tests/golden/shelf_inputs.npz   the Shelf BODY_25 detections + calibrations, packed
  tests/golden/shelf_ref.npz      what the reference computed on them, per frame

The reference functions exercised are exactly the hot path of SURVEY.md §8(a):
MvTracker.update_4d (/root/reference/src/motion_capture.py:873-963) and everything
below it. Hooks only *record*; they never change a value the reference computes.
"""
import argparse
import glob
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pack_shelf_inputs(max_people=6):
    """OpenPose JSON + calib JSON of the reference's Shelf sample → dense arrays."""
    base = os.path.join(ref_shim.REFERENCE_ROOT, "data", "shelf")
    cams = sorted(os.listdir(os.path.join(base, "kps_opn")))
    K, RT, wh = [], [], []
    for c in cams:
        with open(os.path.join(base, "calibs", f"{c}.json")) as f:
            js = json.load(f)
        K.append(np.array(js["K"], dtype=np.float64).reshape(3, 3))
        RT.append(np.array(js["RT"], dtype=np.float64).reshape(3, 4))
        wh.append(js["imgSize"])
    per_cam = []
    for c in cams:
        paths = sorted(glob.glob(os.path.join(base, "kps_opn", c, "*.json")),
                       key=lambda p: int(os.path.basename(p).split("_")[1]))
        per_cam.append(paths)
    n_frames = min(len(p) for p in per_cam)
    kps = np.zeros((n_frames, len(cams), max_people, 25, 3), dtype=np.float64)
    n_pose = np.zeros((n_frames, len(cams)), dtype=np.int32)
    for ci, paths in enumerate(per_cam):
        for fi in range(n_frames):
            with open(paths[fi]) as f:
                people = json.load(f)["people"]
            assert len(people) <= max_people
            n_pose[fi, ci] = len(people)
            for pi, person in enumerate(people):
                kps[fi, ci, pi] = np.array(person["pose_keypoints_2d"], dtype=np.float64).reshape(25, 3)
    return dict(kps25=kps, n_pose=n_pose, K=np.array(K), RT=np.array(RT), img_wh=np.array(wh, dtype=np.int32))


class Recorder:
    """Wraps a handful of reference callables to log their inputs/outputs."""

    def __init__(self, ref):
        self.ref = ref
        self.solves = []         # dicts, in call order
        self.triangulations = []
        self.als_calls = []
        self.cur_frame = -1
        self.n_inv = 0
        self.next_track_id = 0
        self._install()

    def _install(self):
        ref = self.ref
        rec = self
        import scipy.optimize as so

        real_ls = so.least_squares

        def ls_ik(fun, x0, **kw):
            res = real_ls(fun, x0, **kw)
            rec.solves.append(dict(frame=rec.cur_frame, kind="ik", n=len(x0), max_nfev=kw.get("max_nfev"),
                                   x0=np.array(x0, dtype=np.float64).copy(), x=res.x.copy(), nfev=res.nfev,
                                   njev=res.njev, status=res.status, cost=res.cost, optimality=res.optimality,
                                   f0=fun(np.asarray(x0, dtype=np.float64)).copy()))
            return res

        def ls_tri(fun, x0, **kw):
            res = real_ls(fun, x0, **kw)
            rec.solves.append(dict(frame=rec.cur_frame, kind="tri", n=len(x0), max_nfev=kw.get("max_nfev"),
                                   x0=np.array(x0, dtype=np.float64).copy(), x=res.x.copy(), nfev=res.nfev,
                                   njev=res.njev, status=res.status, cost=res.cost, optimality=res.optimality,
                                   f0=fun(np.asarray(x0, dtype=np.float64)).copy()))
            return res

        ref.ik.least_squares = ls_ik
        ref.mvu.least_squares = ls_tri

        real_tri = ref.mvu.triangulate_point_groups_from_multiple_views_linear

        def tri(proj, pts, min_score, post_optimize=False, n_max_iter=2):
            lin = real_tri(proj, pts, min_score, False, n_max_iter)
            out = real_tri(proj, pts, min_score, post_optimize, n_max_iter)
            rec.triangulations.append(dict(frame=rec.cur_frame, P=np.array(proj).copy(), pts=np.array(pts).copy(),
                                           min_score=min_score, linear=lin.copy(), out=out.copy()))
            return out

        ref.ik.triangulate_point_groups_from_multiple_views_linear = tri

        real_als = ref.mva.match_als
        real_inv = np.linalg.inv

        def counting_inv(a):
            rec.n_inv += 1
            return real_inv(a)

        def als(W, dim_groups, **kw):
            rec.n_inv = 0
            np.linalg.inv = counting_inv
            try:
                match_mat, x_bin = real_als(W, dim_groups, **kw)
            finally:
                np.linalg.inv = real_inv
            rec.als_calls.append(dict(frame=rec.cur_frame, W=np.array(W).copy(), w_dtype=str(np.asarray(W).dtype),
                                      dim_groups=np.array(dim_groups, dtype=np.int32), n_iter=rec.n_inv // 2,
                                      x_bin=np.array(x_bin).copy(), match_mat=np.array(match_mat).copy()))
            return match_mat, x_bin

        ref.mc.match_als = als

        real_init = ref.mc.MvTracklet.__init__

        def tl_init(self_, *a, **kw):
            real_init(self_, *a, **kw)
            self_.golden_id = rec.next_track_id
            rec.next_track_id += 1

        ref.mc.MvTracklet.__init__ = tl_init


def frames_from_packed(ref, packed, frm_idx, calibs):
    """Build the reference's List[FrameData] for one frame, the way `--mode prepare`
    (/root/reference/src/motion_capture.py:974-1005) does, without touching disk."""
    mc = ref.mc
    out = []
    for ci in range(packed["kps25"].shape[1]):
        poses = {}
        for pi in range(int(packed["n_pose"][frm_idx, ci])):
            kps = packed["kps25"][frm_idx, ci, pi]
            coco = ref.pose_def.conversion_openpose_25_to_coco(kps)
            poses[pi] = ref.pose_def.Pose(ref.pose_def.KpsFormat.COCO, keypoints=coco[:, :2],
                                          keypoints_score=coco[:, -1][:, np.newaxis], box=None)
        out.append(mc.FrameData(frm_idx, poses, calibs[ci], view_id=ci + 1))
    return out


def calibs_from_packed(ref, packed):
    calibs = []
    for ci in range(len(packed["K"])):
        K = packed["K"][ci]
        RT = packed["RT"][ci]
        P = K @ RT
        kr_inv = RT[:3, :3].transpose() @ np.linalg.inv(K)
        calibs.append(ref.common.Calib(K=K, Rt=RT, P=P, Kr_inv=kr_inv, img_wh_size=list(packed["img_wh"][ci])))
    return calibs


def warm_start_tracks(ref, rec, tracker, packed, f0):
    """Seed the reference's tracker with one Confirmed tracklet per person whose last PoseShapeParam is the generator's
    ground truth at frame f0 (the steady state bench.py measures; a crowded scene's own frame 1 is the reference's
    non-converging no-track case). The tracklets are real MvTracklet instances built without running __init__'s birth
    solve; everything the reference then computes on them (predict, association, update, lifecycle) is its own code."""
    mc, ik = ref.mc, ref.ik
    skel = tracker.skeleton
    init = dict(ids=[], param=[], joints=[])
    n_people = packed["gt_root"].shape[1]
    for pi in range(n_people):
        prm = ik.PoseShapeParam(packed["gt_root"][f0, pi].copy(), packed["gt_euler"][f0, pi].copy(),
                                skel.ref_side_bone_lens.copy() * packed["gt_scale"][pi])
        locs, _ = ik.foward_kinematics(skel, prm)
        pose = ref.pose_def.Pose(ref.pose_def.KpsFormat.BASIC_18, keypoints=locs, keypoints_score=np.ones((len(locs), 1)),
                                 box=None)
        t = object.__new__(mc.MvTracklet)
        t.frame_idxs, t.cam_poses_2d, t.cam_projs, t.cam_calibs = [f0], [[]], [[]], [[]]
        t.skel = skel
        t.poses = [(f0, prm, pose)]
        t.time_since_update, t.hits, t.state, t.max_age, t.n_inits = 0, 3, mc.TrackState.Confirmed, 0, 3
        t.golden_id = rec.next_track_id
        rec.next_track_id += 1
        tracker.tracklets.append(t)
        init["ids"].append(t.golden_id)
        init["param"].append(np.concatenate([prm.root, prm.euler_angles.reshape(-1), prm.bone_lens]))
        init["joints"].append(np.array(locs))
    return init


def run_reference(packed, first_frame, last_frame, verbose=True, warm=False):
    """Drive the reference tracker over frames [first_frame, last_frame] and record goldens."""
    ref = ref_shim.load()
    rec = Recorder(ref)
    mc = ref.mc
    calibs = calibs_from_packed(ref, packed)
    tracker = mc.MvTracker(ref.ik.load_skeleton())
    gold = {}
    if warm:
        init = warm_start_tracks(ref, rec, tracker, packed, first_frame - 1)
        gold["init_ids"] = np.array(init["ids"], dtype=np.int32)
        gold["init_param"] = np.array(init["param"])
        gold["init_joints"] = np.array(init["joints"])
        gold["init_state"] = np.array([[mc.TrackState.Confirmed.value, 3, 0, 1]] * len(init["ids"]), dtype=np.int32)
    times = []
    import contextlib
    import io
    for frm_idx in range(first_frame, last_frame + 1):
        rec.cur_frame = frm_idx
        d_frames = frames_from_packed(ref, packed, frm_idx, calibs)
        with contextlib.redirect_stdout(io.StringIO()):
            d_frames = [mc.filter_bad_pose(f, 0.01, 4, 5) for f in d_frames]
        kept = [sorted(f.poses.keys()) for f in d_frames]
        n_als_before = len(rec.als_calls)
        n_solve_before = len(rec.solves)
        alive_before = [t.golden_id for t in tracker.tracklets]
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()) as buf, np.errstate(all="ignore"):
            # association result is not returned by update_4d; capture it through the module hook
            real_assoc = mc.associate_tracking
            holder = {}

            def assoc(tlets, frames, min_pixel_error_hard_threshold):
                r = real_assoc(tlets, frames, min_pixel_error_hard_threshold)
                holder["m"] = r
                return r

            mc.associate_tracking = assoc
            try:
                tracker.update_4d(frm_idx, d_frames, None)
            finally:
                mc.associate_tracking = real_assoc
        times.append(time.perf_counter() - t0)
        m = holder["m"]
        als = rec.als_calls[n_als_before]
        assert len(rec.als_calls) == n_als_before + 1
        key = f"f{frm_idx:04d}_"
        gold[key + "kept"] = np.array([[1 if p in k else 0 for p in range(packed["kps25"].shape[2])] for k in kept],
                                      dtype=np.uint8)
        gold[key + "alive_before"] = np.array(alive_before, dtype=np.int32)
        gold[key + "dim_groups"] = als["dim_groups"]
        gold[key + "dst"] = np.array(m.dst_mat)
        gold[key + "sim"] = np.array(m.sim_mat)
        gold[key + "als_iters"] = np.int32(als["n_iter"])
        gold[key + "xbin"] = als["x_bin"].astype(np.uint8)
        gold[key + "match_mat"] = als["match_mat"].astype(np.uint8)
        # matches: rows of (track_idx or -1, view, pose_id)
        rows = []
        for gi, (t_idx, sm) in enumerate(m.spatial_time_matches.items()):
            for v, p in zip(sm.view_idxs, sm.pose_ids):
                rows.append((0, gi, t_idx, v, p))
        for gi, sm in enumerate(m.spatial_matches):
            for v, p in zip(sm.view_idxs, sm.pose_ids):
                rows.append((1, gi, -1, v, p))
        gold[key + "matches"] = np.array(rows, dtype=np.int32).reshape(-1, 5)
        gold[key + "printed"] = np.array(buf.getvalue().count("more than one pose per view"), dtype=np.int32)
        alive_after = tracker.tracklets
        gold[key + "alive_after"] = np.array([t.golden_id for t in alive_after], dtype=np.int32)
        gold[key + "alive_state"] = np.array([[t.state.value, t.hits, t.time_since_update, len(t)]
                                              for t in alive_after], dtype=np.int32).reshape(-1, 4)
        updated = [t for t in alive_after if t.frame_idxs[-1] == frm_idx]
        gold[key + "upd_ids"] = np.array([t.golden_id for t in updated], dtype=np.int32)
        gold[key + "upd_root"] = np.array([t.poses[-1][1].root for t in updated]).reshape(-1, 3)
        gold[key + "upd_euler"] = np.array([t.poses[-1][1].euler_angles for t in updated]).reshape(-1, 18, 3)
        gold[key + "upd_blens"] = np.array([t.poses[-1][1].bone_lens for t in updated]).reshape(-1, 11)
        gold[key + "upd_joints"] = np.array([t.poses[-1][2].keypoints for t in updated]).reshape(-1, 18, 3)
        gold[key + "upd_views"] = np.array([[1 if v in [vp[0] for vp in t.cam_poses_2d[-1]] else 0
                                             for v in range(len(calibs))] for t in updated],
                                           dtype=np.uint8).reshape(-1, len(calibs))
        gold[key + "upd_pose_ids"] = np.array(
            [[dict((vp[0], pid) for vp, pid in zip(t.cam_poses_2d[-1], _pose_ids(t, d_frames))).get(v, -1)
              for v in range(len(calibs))] for t in updated], dtype=np.int32).reshape(-1, len(calibs))
        # solves of this frame
        sl = rec.solves[n_solve_before:]
        gold[key + "solve_meta"] = np.array([[{"ik": 0, "tri": 1}[s["kind"]], s["n"], s["max_nfev"], s["nfev"],
                                              s["njev"], s["status"]] for s in sl], dtype=np.int32).reshape(-1, 6)
        gold[key + "solve_cost"] = np.array([s["cost"] for s in sl], dtype=np.float64)
        for si, s in enumerate(sl):
            gold[key + f"solve{si}_x0"] = s["x0"]
            gold[key + f"solve{si}_x"] = s["x"]
        if verbose:
            print(f"frame {frm_idx}: T={len(alive_before)} n={als['dim_groups'][-1]} als_it={als['n_iter']} "
                  f"solves={len(sl)} alive_after={len(alive_after)} dead={len(tracker.dead_tracklets)} "
                  f"{times[-1]*1e3:.0f} ms", flush=True)
    tri_rec = [t for t in rec.triangulations]
    gold["tri_count"] = np.int32(len(tri_rec))
    for ti, t in enumerate(tri_rec):
        gold[f"tri{ti}_frame"] = np.int32(t["frame"])
        gold[f"tri{ti}_P"] = t["P"]
        gold[f"tri{ti}_pts"] = t["pts"]
        gold[f"tri{ti}_linear"] = t["linear"]
        gold[f"tri{ti}_out"] = t["out"]
    all_tlets = tracker.tracklets + tracker.dead_tracklets
    all_tlets = sorted(all_tlets, key=lambda tl: -len(tl))
    gold["final_ids"] = np.array([t.golden_id for t in all_tlets], dtype=np.int32)
    gold["final_len"] = np.array([len(t) for t in all_tlets], dtype=np.int32)
    gold["final_first_frame"] = np.array([t.frame_idxs[0] for t in all_tlets], dtype=np.int32)
    gold["final_state"] = np.array([t.state.value for t in all_tlets], dtype=np.int32)
    gold["frame_times_s"] = np.array(times)
    gold["first_frame"] = np.int32(first_frame)
    gold["last_frame"] = np.int32(last_frame)
    return gold, tracker


def _pose_ids(tlet, d_frames):
    """pose ids (dict keys) of the 2D poses used in the tracklet's last update, by identity."""
    ids = []
    for v, pose in tlet.cam_poses_2d[-1]:
        pid = [k for k, p in d_frames[v].poses.items() if p is pose]
        ids.append(pid[0] if pid else -1)
    return ids


def make_ik3d_golden(n_records=10):
    """The 3D-target IK variants (src/inverse_kinematics.py:280-336, dead behind `use_only_reproj = True`) run by the
    REFERENCE on the solves of the Shelf golden: for each recorded update / birth, triangulate (post_optimize) then
    solve_pose and solve_pose_bone_lens with the reference's own functions. -> tests/golden/ik3d_ref.npz"""
    ref = ref_shim.load()
    ik = ref.ik
    skel = ik.load_skeleton()
    inp = np.load(os.path.join(GOLD, "shelf_inputs.npz"))
    g = np.load(os.path.join(GOLD, "shelf_ref.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import GoldenTable, fkey
    tab = GoldenTable(g)
    coco = ref.pose_def.conversion_openpose_25_to_coco
    out, n = {}, 0
    solver_cls = ik.PoseSolver
    for f in (1, 2, 3, 30, 31):
        k = fkey(f)
        ids_before = tab.seek(f)
        for u, tid in enumerate(g[k + "upd_ids"].tolist()):
            if n >= n_records:
                break
            views = np.nonzero(g[k + "upd_views"][u])[0]
            pids = g[k + "upd_pose_ids"][u][views]
            cam_kps = [coco(inp["kps25"][f, v, p]) for v, p in zip(views, pids)]
            Ps = [inp["K"][v] @ inp["RT"][v] for v in views]
            birth = tid not in ids_before
            ps = solver_cls(skel, None, [c.copy() for c in cam_kps], Ps, ref.pose_def.KpsFormat.COCO)   # adds the mid spine, builds the index maps
            obs3d = ref.mvu.triangulate_point_groups_from_multiple_views_linear(ps.cam_projs, ps.cam_poses_2d, 0.01, True)
            if birth:
                root = 0.5 * (obs3d[ps.obs_kps_idx_map[ref.pose_def.KpsType.L_Hip], :3] + obs3d[ps.obs_kps_idx_map[ref.pose_def.KpsType.R_Hip], :3])
                init = ik.PoseShapeParam(root, np.zeros((18, 3)), skel.ref_side_bone_lens.copy())
                nfev = 50
            else:
                x = tab.table[tid]["param"]
                init = ik.PoseShapeParam(x[:3].copy(), x[3:57].reshape(18, 3).copy(), x[57:].copy())
                nfev = 5
            calls = []
            real = ik.least_squares

            def spy(fun, x0, **kw):
                r = real(fun, x0, **kw)
                calls.append(r)
                return r
            ik.least_squares = spy
            try:
                p1 = ik.solve_pose(skel, obs3d, ps.obs_kps_idxs, ps.skel_kps_idxs, init, nfev)
                p2 = ik.solve_pose_bone_lens(skel, obs3d, ps.obs_kps_idxs, ps.skel_kps_idxs, p1, nfev)
            finally:
                ik.least_squares = real
            locs, _ = ik.foward_kinematics(skel, p2)
            pre = f"r{n}_"
            out[pre + "frame"], out[pre + "birth"], out[pre + "nfev_cap"] = np.int32(f), np.int32(birth), np.int32(nfev)
            out[pre + "cam_kps"], out[pre + "P"] = np.array(cam_kps), np.array(Ps)
            out[pre + "obs3d"] = obs3d
            out[pre + "obs_idx"], out[pre + "skel_idx"] = np.array(ps.obs_kps_idxs, dtype=np.int32), np.array(ps.skel_kps_idxs, dtype=np.int32)
            pk = lambda q: np.concatenate([q.root.flatten(), q.euler_angles.flatten(), q.bone_lens.flatten()])
            out[pre + "x0"], out[pre + "x1"], out[pre + "x2"], out[pre + "joints"] = pk(init), pk(p1), pk(p2), np.array(locs)
            out[pre + "meta"] = np.array([[c.nfev, c.njev, c.status] for c in calls], dtype=np.int32)
            out[pre + "cost"] = np.array([c.cost for c in calls])
            print(f"record {n}: frame {f} track {tid} birth {birth} views {len(views)} meta {out[pre + 'meta'].tolist()} cost {out[pre + 'cost']}")
            n += 1
    out["count"] = np.int32(n)
    np.savez_compressed(os.path.join(GOLD, "ik3d_ref.npz"), **out)
    print("wrote ik3d_ref.npz", n)


def make_alt_matcher_golden():
    """The reference's alternative matchers (dead code there) run by the REFERENCE on golden frames:
    match_objects_across_views (src/motion_capture.py:166-241) and MvTracker.tracklet_to_poses_association (:852-871).
    -> tests/golden/altmatch_ref.npz"""
    ref = ref_shim.load()
    mc = ref.mc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import GoldenTable, fkey
    import contextlib
    import io
    out, n = {}, 0
    src_idx, dst_idx = ref.pose_def.get_common_kps_idxs(ref.pose_def.KpsFormat.COCO, ref.pose_def.KpsFormat.BASIC_18)
    out["common_coco"], out["common_b18"] = np.array(src_idx, dtype=np.int32), np.array(dst_idx, dtype=np.int32)
    for name, frames in (("shelf", (1, 2, 50, 120, 250)), ("synth_c8p6", (2, 3)), ("warm_c8p16", (3,)), ("warm_c8p32", (3,))):
        inp = np.load(os.path.join(GOLD, f"{name}_inputs.npz"))
        g = np.load(os.path.join(GOLD, f"{name}_ref.npz"))
        packed = {k: inp[k] for k in inp.files}
        calibs = calibs_from_packed(ref, packed)
        tab = GoldenTable(g)
        for f in frames:
            d_frames = frames_from_packed(ref, packed, f, calibs)
            with contextlib.redirect_stdout(io.StringIO()):
                d_frames = [mc.filter_bad_pose(fr, 0.01, 4, 5) for fr in d_frames]
            pre = f"r{n}_"
            out[pre + "scene"], out[pre + "frame"] = np.array(name), np.int32(f)
            for ti, thr in enumerate((200.0, 25.0)):
                try:
                    grps = mc.match_objects_across_views(f, d_frames, False, thr, 0.01)
                    rows = [(gi, vid - 1, pid) for gi, gr in enumerate(grps) for vid, (pid, _) in zip(gr.view_ids, gr.id_poses)]
                    out[pre + f"raises{ti}"] = np.int32(0)
                except ValueError:     # a pose pair without a commonly visible joint: NaN cost, SciPy's assignment raises
                    rows = []
                    out[pre + f"raises{ti}"] = np.int32(1)
                out[pre + f"groups{ti}"] = np.array(rows, dtype=np.int32).reshape(-1, 3)
                out[pre + f"thr{ti}"] = np.float64(thr)
            # 3D ray association against the reference's own tracks before this frame
            joints = tab.joints(f)

            class _T:
                def __init__(self, j):
                    self.last_pose_3d = ref.pose_def.Pose(ref.pose_def.KpsFormat.BASIC_18, j.reshape(18, 3), np.ones((18, 1)), None)
            tl = [_T(j) for j in joints]
            rows, costs = [], []
            for v, fr in enumerate(d_frames):
                for dmax in (0.1,):
                    m = mc.MvTracker.tracklet_to_poses_association(tl, fr, max_dst=dmax)
                    rows += [(v, t, p) for t, p in m]
                for t, tlet in enumerate(tl):
                    for pid, pose in fr.poses.items():
                        costs.append((v, t, pid, mc.MvTracker.tracklet_to_pose_2d_cost(tlet, pose, fr.calib)))
            out[pre + "ray_matches"] = np.array(rows, dtype=np.int32).reshape(-1, 3)
            out[pre + "ray_costs"] = np.array(costs, dtype=np.float64).reshape(-1, 4)
            print(f"record {n}: {name} frame {f}: groups {len(set(out[pre + 'groups0'][:, 0].tolist()))} / "
                  f"{len(set(out[pre + 'groups1'][:, 0].tolist()))}, ray matches {len(rows)}")
            n += 1
    out["count"] = np.int32(n)
    np.savez_compressed(os.path.join(GOLD, "altmatch_ref.npz"), **out)
    print("wrote altmatch_ref.npz", n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["shelf", "synth", "warm", "ik3d", "altmatch"])
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--out", default=None)
    ap.add_argument("--scene", default=None, help="warm: one scene of synthetic.WARM_SCENES (default all)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if args.what == "ik3d":
        make_ik3d_golden()
        return
    if args.what == "altmatch":
        make_alt_matcher_golden()
        return
    if args.what == "warm":
        # tracked (steady-state) frames at the BASELINE shapes: the reference's tracker is seeded from the generator's
        # ground truth at frame `first - 1` (warm_start_tracks) and then runs frames first..last itself
        sys.path.insert(0, ROOT)
        from multiview_motion_capture_b200 import synthetic
        for name, spec in synthetic.WARM_SCENES.items():
            if args.scene and name != args.scene:
                continue
            packed = synthetic.make_warm_scene(name)
            first, last = spec["first"], spec["last"]
            slim = {k: (v[:last + 1] if k in ("kps25", "n_pose", "gt_person", "gt_joints", "gt_root", "gt_euler") else v)
                    for k, v in packed.items()}
            np.savez_compressed(os.path.join(GOLD, f"warm_{name}_inputs.npz"), **slim)
            gold, _ = run_reference(packed, first, last, warm=True)
            np.savez_compressed(os.path.join(GOLD, f"warm_{name}_ref.npz"), **gold)
            print("wrote warm", name, "total s:", gold["frame_times_s"].sum())
        return
    if args.what == "shelf":
        packed = pack_shelf_inputs()
        np.savez_compressed(os.path.join(GOLD, "shelf_inputs.npz"), **packed)
        last = min(args.frames, packed["kps25"].shape[0] - 1)
        gold, _ = run_reference(packed, 1, last)
        out = args.out or os.path.join(GOLD, "shelf_ref.npz")
        np.savez_compressed(out, **gold)
        print("wrote", out, "total s:", gold["frame_times_s"].sum())
    else:
        sys.path.insert(0, ROOT)
        from multiview_motion_capture_b200 import synthetic
        for name, kw in synthetic.GOLDEN_SCENES.items():
            packed = synthetic.make_clip(**kw)
            np.savez_compressed(os.path.join(GOLD, f"synth_{name}_inputs.npz"), **packed)
            gold, _ = run_reference(packed, 1, packed["kps25"].shape[0] - 1)
            np.savez_compressed(os.path.join(GOLD, f"synth_{name}_ref.npz"), **gold)
            print("wrote synth", name)


if __name__ == "__main__":
    main()
