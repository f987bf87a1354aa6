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
    # TEST ONLY: the launcher binds the kernel emulator, then runs the CLI script as __main__
    run = lambda *a: subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_dropin_emu.py"), *a], check=True,
                                    capture_output=True, text=True)
    run("--mode", "prepare", "--opn_kps_dir", str(tmp / "kps"), "--calib_dir", str(tmp / "calibs"), "--out_data_dir",
        str(tmp / "dframes"))
    run("--mode", "run", "--video_dir", "", "--data_dir", str(tmp / "dframes"), "--output_dir", str(tmp / "out"),
        "--max_frames", "3")
    # the same clip from the reference's per-frame pickles alone (no packed clip.npz beside them)
    os.makedirs(tmp / "pkl_only")
    for f in os.listdir(tmp / "dframes"):
        if f.endswith(".pkl"):
            os.link(tmp / "dframes" / f, tmp / "pkl_only" / f)
    run("--mode", "run", "--video_dir", "", "--data_dir", str(tmp / "pkl_only"), "--output_dir", str(tmp / "out_pkl"),
        "--max_frames", "3")
    return tmp


def test_prepare_writes_reference_layout_pickles(workdir):
    sys.path.insert(0, DROPIN)
    try:
        files = sorted(os.listdir(workdir / "dframes"))
        assert files == [f"{i:06d}.pkl" for i in range(4)] + ["clip.npz"]
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


def test_native_json_parser_equals_json_load(workdir, emu):
    """mvmc_parse_openpose_host / _files_host give exactly the doubles json.load gives (same correctly rounded
    conversion), for every person of every file; the packed clip.npz equals the fixture it was written from."""
    from multiview_motion_capture_b200 import ingest
    inp, _ = golden("shelf")
    paths = sorted((workdir / "kps" / "2").glob("*.json"))
    kps, cnt = ingest.parse_openpose_files(paths, 6, threads=3)
    for i, p in enumerate(paths):
        with open(p) as f:
            people = json.load(f)["people"]
        assert cnt[i] == len(people)
        for q, person in enumerate(people):
            assert np.array_equal(kps[i, q].reshape(-1), np.array(person["pose_keypoints_2d"], dtype=np.float64))
        assert not kps[i, len(people):].any()
    one, n = ingest.parse_openpose_text(open(paths[1], "rb").read(), 6)
    assert n == cnt[1] and np.array_equal(one, kps[1])
    # odd but valid JSON: other keys, nested values, whitespace, exponents
    txt = b'{ "version":1.3,"extra":{"a":[1,{"b":"x\\"y"}]}, "people" : [ {"person_id":[-1], "pose_keypoints_2d":[' + \
          b",".join([b"1e-3", b"-2.5E+1", b"0"] * 25) + b'], "face_keypoints_2d":[]} ] }'
    one, n = ingest.parse_openpose_text(txt, 2)
    assert n == 1 and np.array_equal(one[0], np.tile([1e-3, -25.0, 0.0], (25, 1)))
    clip = ingest.load_clip_npz(workdir / "dframes" / "clip.npz")
    assert np.array_equal(clip["n_pose"], inp["n_pose"][:4])
    P = clip["kps25"].shape[2]
    assert np.array_equal(clip["kps25"], inp["kps25"][:4, :, :P]) and np.array_equal(clip["K"], inp["K"])


def test_packed_and_pickle_inputs_give_the_same_tracklets(workdir):
    """`--mode run` from clip.npz (BODY_25 -> COCO on the device) and from the reference's pickles: identical outputs; and
    the tracklets.npz side format holds the same numbers as tracklets.pkl."""
    from multiview_motion_capture_b200.tracklets_io import load_tracklets_npz
    a = load_tracklets_npz(workdir / "out" / "tracklets.npz")
    b = load_tracklets_npz(workdir / "out_pkl" / "tracklets.npz")
    assert len(a) == len(b) == 2
    for ta, tb in zip(a, b):
        assert ta.frame_idxs == tb.frame_idxs and ta.views == tb.views and (ta.state, ta.hits) == (tb.state, tb.hits)
        for pa, pb in zip(ta.poses, tb.poses):
            assert np.array_equal(pa[1].euler_angles, pb[1].euler_angles) and np.array_equal(pa[-1].keypoints, pb[-1].keypoints)
    sys.path.insert(0, DROPIN)
    try:
        import motion_capture
        sys.modules["__main__"].MvTracklet = motion_capture.MvTracklet
        sys.modules["__main__"].TrackState = motion_capture.TrackState
        with open(workdir / "out" / "tracklets.pkl", "rb") as f:
            tl = pickle.load(f)["tracklets"]
    finally:
        sys.path.remove(DROPIN)
    for t, ta in zip(tl, a):
        assert t.frame_idxs == ta.frame_idxs and t.state.value == ta.state and t.hits == ta.hits
        assert [[v for v, _ in fr] for fr in t.cam_poses_2d] == [[v for v, _ in fr] for fr in ta.views]
        for p, pa in zip(t.poses, ta.poses):      # the viz accessors: p[0], p[-1].keypoints, p[-1].pose_type
            assert p[0] == pa[0] and np.array_equal(p[-1].keypoints, pa[-1].keypoints) and pa[-1].pose_type == "BASIC_18"
            assert np.array_equal(p[1].root, pa[1].root) and np.array_equal(p[1].bone_lens, pa[1].bone_lens)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference sources")
def test_tracklets_pickle_loads_in_the_reference_and_feeds_its_viz_accessors(workdir):
    """Round trip through the reference's own classes (container only): tracklets.pkl is unpickled with the REFERENCE's
    modules (`__main__.MvTracklet` -> its motion_capture.MvTracklet, its pose_def.Pose, inverse_kinematics.PoseShapeParam,
    common.Calib), then the expressions of viz_tracklets / plot_poses_3d_reprojects (src/motion_capture.py:1177-1198,
    src/pose_viz.py:69-160) are evaluated on it: (p[0], p[-1]) pairs, get_pose_bones_index(pose.pose_type), bone end
    points and re-projections."""
    script = r"""
import sys, pickle, numpy as np
sys.path.insert(0, %r)
import ref_shim
ref = ref_shim.load()
import __main__
__main__.MvTracklet = ref.mc.MvTracklet
__main__.TrackState = ref.mc.TrackState
with open(%r, 'rb') as f:
    all_tlets = pickle.load(f)['tracklets']
assert all(type(t) is ref.mc.MvTracklet for t in all_tlets)
tracks = [[(p[0], p[-1]) for p in tlet.poses] for tlet in all_tlets[:10]]          # viz_tracklets :1195
bones = ref.pose_def.get_pose_bones_index(tracks[0][0][1].pose_type)                 # pose_viz.py:100
assert len(bones) > 10
P = all_tlets[0].cam_projs[0][0]
for trk in tracks:
    for frm, pose in trk:
        assert isinstance(frm, int) and type(pose) is ref.pose_def.Pose
        for b in bones:
            p0, p1 = pose.keypoints[b[0], :3], pose.keypoints[b[1], :3]              # pose_viz.py:142
        n = len(pose.keypoints)
        rep = P @ np.concatenate([pose.keypoints, np.ones((n, 1))], axis=-1).T       # pose_viz.py:153
        rep = (rep[:2] / rep[2]).T
        assert np.isfinite(rep).all() and (np.abs(rep) < 5000).all()
t = all_tlets[0]
assert t.is_confirmed() and len(t) == 3 and t.last_pose_3d.keypoints.shape == (18, 3)
assert type(t.poses[0][1]) is ref.ik.PoseShapeParam and type(t.cam_calibs[0][0]) is ref.common.Calib
print('ok', len(all_tlets))
""" % (os.path.join(ROOT, "oracle"), str(workdir / "out" / "tracklets.pkl"))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True)
    assert r.returncode == 0 and "ok 2" in r.stdout, r.stderr[-2000:]
