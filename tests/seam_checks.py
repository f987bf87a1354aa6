"""Device-agnostic checks of the same-signature Python seams against the reference goldens: the drop-in
modules are imported under the reference's module names and called with the reference's argument types."""
import os
import sys

import numpy as np

import mvmc_oracle as o
from helpers import ROOT, GoldenTable, fkey, golden, golden_matches

DROPIN = os.path.join(ROOT, "multiview_motion_capture_b200", "dropin")


def dropin_modules():
    """The drop-in package's modules under the reference's names (common, pose_def, motion_capture, ...)."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    for name in ("common", "pose_def", "motion_capture", "mv_math_util", "mv_association", "inverse_kinematics"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(DROPIN):
            del sys.modules[name]
    import common, inverse_kinematics, motion_capture, mv_association, mv_math_util, pose_def  # noqa: E401
    return dict(common=common, pose_def=pose_def, mc=motion_capture, mva=mv_association, mvu=mv_math_util, ik=inverse_kinematics)


def frames_of(mods, inp, f, kept=None):
    """List[FrameData] of golden frame f, built like --mode prepare does; `kept` [C,P] drops the filtered poses."""
    common, pose_def = mods["common"], mods["pose_def"]
    kps = o.body25_to_coco(inp["kps25"])
    out = []
    for v in range(kps.shape[1]):
        K, Rt = inp["K"][v], inp["RT"][v]
        calib = common.Calib(K=K, Rt=Rt, P=K @ Rt, Kr_inv=Rt[:3, :3].T @ np.linalg.inv(K), img_wh_size=list(inp["img_wh"][v]))
        poses = {}
        for p in range(int(inp["n_pose"][f, v])):
            if kept is None or kept[v, p]:
                poses[p] = pose_def.Pose(pose_def.KpsFormat.COCO, keypoints=kps[f, v, p, :, :2].copy(),
                                         keypoints_score=kps[f, v, p, :, 2:3].copy(), box=None)
        out.append(common.FrameData(f, poses, calib, view_id=v + 1))
    return out


class _Tlet:
    def __init__(self, pose_def, joints):
        self.last_pose_3d = pose_def.Pose(pose_def.KpsFormat.BASIC_18, joints.reshape(18, 3), np.ones((18, 1)), None)


def check_associate_tracking(name, frames):
    """associate_tracking(tlets, frames, thr) -> SpatialTimeMatch: matches, matrices and X_bin of the reference."""
    mods = dropin_modules()
    mc = mods["mc"]
    inp, g = golden(name)
    tab = GoldenTable(g)
    for f in frames:
        k = fkey(f)
        tlets = [_Tlet(mods["pose_def"], j) for j in tab.joints(f)]
        m = mc.associate_tracking(tlets, frames_of(mods, inp, f, g[k + "kept"]), 10)
        tm, ng = golden_matches(g, f)
        assert {t: list(zip(s.view_idxs, s.pose_ids)) for t, s in m.spatial_time_matches.items()} == tm, (name, f)
        assert [list(zip(s.view_idxs, s.pose_ids)) for s in m.spatial_matches] == ng, (name, f)
        assert np.array_equal(m.match_mat, g[k + "xbin"].astype(bool)), (name, f)
        assert m.dst_mat.dtype == g[k + "dst"].dtype and m.sim_mat.dtype == g[k + "sim"].dtype
        assert np.abs(m.dst_mat - g[k + "dst"]).max() <= 1e-7 and np.abs(m.sim_mat.astype(np.float64) - g[k + "sim"]).max() <= 1e-9
        for s in list(m.spatial_matches) + list(m.spatial_time_matches.values()):
            for v, p, gi in zip(s.view_idxs, s.pose_ids, s.cost_matrix_idxs):
                assert m.find_matrix_idx_from_view_pose_id(v, p) == gi
                assert m.find_spatial_match(v, p) is not None


def check_match_als(name, frames):
    """mv_association.match_als(W, dimGroup) -> (match_mat, X_bin) on the reference's own similarity matrices."""
    mva = dropin_modules()["mva"]
    _, g = golden(name)
    for f in frames:
        k = fkey(f)
        mm, xb = mva.match_als(g[k + "sim"], g[k + "dim_groups"].tolist())
        assert np.array_equal(xb, g[k + "xbin"].astype(bool)), (name, f)
        assert np.array_equal(mm, g[k + "match_mat"].astype(bool)), (name, f)
        assert np.array_equal(mva.transform_closure(g[k + "xbin"]), g[k + "match_mat"]), (name, f)


def check_solver_and_fk(name, frame):
    """PoseSolver(...).solve() and foward_kinematics(skel, param) with the reference's argument types."""
    mods = dropin_modules()
    ik = mods["ik"]
    inp, g = golden(name)
    k = fkey(frame)
    skel = ik.load_skeleton()
    prm = ik.PoseShapeParam(g[k + "upd_root"][0], g[k + "upd_euler"][0], g[k + "upd_blens"][0])
    locs, _ = ik.foward_kinematics(skel, prm)          # two values, like the reference
    assert np.abs(locs - g[k + "upd_joints"][0]).max() <= 1e-12
    kps = o.body25_to_coco(inp["kps25"])
    views = np.nonzero(g[k + "upd_views"][0])[0]
    pids = g[k + "upd_pose_ids"][0][views]
    cam_kps = [kps[frame, v, p] for v, p in zip(views, pids)]
    Ps = [inp["K"][v] @ inp["RT"][v] for v in views]
    birth = len(g[k + "alive_before"]) == 0
    init = None
    if not birth:
        tab = GoldenTable(g)
        tid = int(g[k + "upd_ids"][0])
        tab.seek(frame)
        x = tab.table[tid]["param"]
        init = ik.PoseShapeParam(x[:3], x[3:57].reshape(18, 3), x[57:])
    param, pose = ik.PoseSolver(skel, init, cam_kps, Ps, obs_kps_format=mods["pose_def"].KpsFormat.COCO).solve()
    assert pose.keypoints.shape == (18, 3) and param.euler_angles.shape == (18, 3)
    assert np.abs(pose.keypoints - g[k + "upd_joints"][0]).max() <= 3e-2
