"""The drop-in boundary: the reference's CLI flags, its per-frame input pickles and its tracklets.pkl
layout, exercised end to end through the kernel emulator on a few Shelf frames (CPU tier)."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from helpers import EMU_LIB, ROOT, fkey, golden

DROPIN = os.path.join(ROOT, "multiview_motion_capture_b200", "dropin")


def _write_openpose_tree(tmp, n_frames):
    """OpenPose-1.3-style JSON + calibration JSON files rebuilt from the packed Shelf fixture."""
    inp, _ = golden("shelf")
    C = inp["kps25"].shape[1]
    for c in range(C):
        os.makedirs(tmp / "kps" / str(c), exist_ok=True)
        os.makedirs(tmp / "calibs", exist_ok=True)
        with open(tmp / "calibs" / f"{c}.json", "w") as f:
            json.dump({"K": inp["K"][c].reshape(-1).tolist(), "RT": inp["RT"][c].reshape(-1).tolist(),
                       "imgSize": inp["img_wh"][c].tolist()}, f)
        for fr in range(n_frames):
            people = [{"pose_keypoints_2d": inp["kps25"][fr, c, p].reshape(-1).tolist()} for p in range(int(inp["n_pose"][fr, c]))]
            with open(tmp / "kps" / str(c) / f"{c}_{fr:012d}_keypoints.json", "w") as f:
                json.dump({"version": 1.3, "people": people}, f)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, emu):
    tmp = tmp_path_factory.mktemp("dropin")
    _write_openpose_tree(tmp, 4)
    env = dict(os.environ, MVMC_LIBRARY=EMU_LIB)   # TEST ONLY: route the CLI subprocess to the kernel emulator
    run = lambda *a: subprocess.run([sys.executable, os.path.join(DROPIN, "motion_capture.py"), *a], env=env, check=True,
                                    capture_output=True, text=True)
    run("--mode", "prepare", "--opn_kps_dir", str(tmp / "kps"), "--calib_dir", str(tmp / "calibs"), "--out_data_dir",
        str(tmp / "dframes"))
    run("--mode", "run", "--video_dir", "", "--data_dir", str(tmp / "dframes"), "--output_dir", str(tmp / "out"),
        "--max_frames", "3")
    return tmp


def test_prepare_writes_reference_layout_pickles(workdir):
    sys.path.insert(0, DROPIN)
    try:
        files = sorted(os.listdir(workdir / "dframes"))
        assert files == [f"{i:06d}.pkl" for i in range(4)]
        with open(workdir / "dframes" / "000001.pkl", "rb") as f:
            frames = pickle.load(f)
        inp, _ = golden("shelf")
        assert [fr.view_id for fr in frames] == [1, 2, 3, 4, 5]
        for v, fr in enumerate(frames):
            assert type(fr).__module__ == "common" and type(fr.calib).__module__ == "common"
            assert sorted(fr.poses) == list(range(int(inp["n_pose"][1, v])))
            assert np.array_equal(fr.calib.P, inp["K"][v] @ inp["RT"][v])
            for pid, pose in fr.poses.items():
                assert type(pose).__module__ == "pose_def" and pose.pose_type.name == "COCO"
                assert pose.keypoints.shape == (17, 2) and pose.keypoints_score.shape == (17, 1)
                assert np.array_equal(pose.to_kps_array()[0], inp["kps25"][1, v, pid, 0])   # Nose stays slot 0
    finally:
        sys.path.remove(DROPIN)


def test_run_mode_writes_tracklets_pickle(workdir):
    """tracklets.pkl: {"tracklets": [...]} sorted by -len, attribute layout of src/motion_capture.py:321-340, and the
    accessors the reference's viz_tracklets uses (p[0], p[-1].keypoints, p[-1].pose_type)."""
    sys.path.insert(0, DROPIN)
    try:
        import motion_capture  # noqa: F401  (classes pickled as __main__.* when run as a script, as in the reference)
        sys.modules["__main__"].MvTracklet = motion_capture.MvTracklet
        sys.modules["__main__"].TrackState = motion_capture.TrackState
        with open(workdir / "out" / "tracklets.pkl", "rb") as f:
            data = pickle.load(f)
    finally:
        sys.path.remove(DROPIN)
    tl = data["tracklets"]
    _, g = golden("shelf")
    assert len(tl) == len(g[fkey(3) + "alive_after"]) == 2
    assert [len(t) for t in tl] == sorted([len(t) for t in tl], reverse=True)
    for t in tl:
        assert t.frame_idxs == [1, 2, 3]
        assert len(t.cam_poses_2d) == len(t.cam_projs) == len(t.poses) == 3 and len(t.cam_calibs) == 1
        assert t.state.name == "Confirmed" and t.hits == 3 and t.time_since_update == 0
        assert (t.max_age, t.n_inits) == (0, 3)
        assert type(t.skel).__module__ == "inverse_kinematics" and t.skel.n_joints == 18
        for (frm, prm, pose), views in zip(t.poses, t.cam_poses_2d):
            assert type(prm).__name__ == "PoseShapeParam"
            assert prm.root.shape == (3,) and prm.euler_angles.shape == (18, 3) and prm.bone_lens.shape == (11,)
            assert pose.pose_type.name == "BASIC_18" and pose.keypoints.shape == (18, 3) and pose.keypoints_score.shape == (18, 1)
            assert len(views) >= 2 and all(isinstance(v, int) and p.pose_type.name == "COCO" for v, p in views)
        p = t.poses[-1]
        assert p[0] == 3 and p[-1].keypoints[:, :3].shape == (18, 3)
    # same people as the reference: 3D joints of the last frame within the solver's noise envelope
    got = sorted(tl, key=lambda t: t.poses[0][2].keypoints[0, 0])
    ref = g[fkey(3) + "upd_joints"]
    ref = ref[np.argsort(g[fkey(1) + "upd_joints"][:, 0, 0])]
    for t, r in zip(got, ref):
        assert np.abs(t.poses[-1][2].keypoints - r).max() < 5e-2
