"""CPU tier: host-side pieces around the path that need neither a GPU nor the emulator - calibration formats, the file sort
key of prepare mode, the tracklets.npz side format (incl. the optional 2D poses), the sharding helpers."""
import json
import os
import pickle
import sys
from enum import Enum
from pathlib import Path

import numpy as np
import pytest

from helpers import ROOT, golden


def test_load_calib_json_and_pkl(tmp_path):
    """src/motion_capture.py:250-272: .json (K, RT, imgSize) and .pkl (K, R, t; 1920 x 1080) calibrations."""
    from multiview_motion_capture_b200.ingest import load_calib_arrays
    inp, _ = golden("shelf")
    K, RT = inp["K"][0], inp["RT"][0]
    with open(tmp_path / "0.json", "w") as f:
        json.dump({"K": K.reshape(-1).tolist(), "RT": RT.reshape(-1).tolist(), "imgSize": [1032, 776]}, f)
    with open(tmp_path / "1.pkl", "wb") as f:
        pickle.dump({"K": K.tolist(), "R": RT[:, :3].tolist(), "t": RT[:, 3].tolist()}, f)
    k0, r0, wh0 = load_calib_arrays(tmp_path / "0.json")
    k1, r1, wh1 = load_calib_arrays(tmp_path / "1.pkl")
    assert np.array_equal(k0, K) and np.array_equal(r0, RT) and list(wh0) == [1032, 776]
    assert np.array_equal(k1, K) and np.array_equal(r1, RT) and list(wh1) == [1920, 1080]
    with pytest.raises(ValueError):
        load_calib_arrays(tmp_path / "2.txt")


def test_prepare_sort_key_is_numeric_like_the_reference():
    """src/motion_capture.py:995 sorts a camera's files by int(stem.split('_')[1]): names that are not zero padded must not
    be paired across cameras in lexicographic order."""
    from multiview_motion_capture_b200.ingest import _frame_key
    names = [f"0_{i}_keypoints.json" for i in (10, 9, 100, 2)]
    assert [p.name for p in sorted(map(Path, names), key=_frame_key)] == [f"0_{i}_keypoints.json" for i in (2, 9, 10, 100)]
    odd = sorted(map(Path, ["b.json", "a.json"]), key=_frame_key)       # no numeric field: falls back to the stem
    assert [p.name for p in odd] == ["a.json", "b.json"]


class _State(Enum):
    Tentative = 1
    Confirmed = 2


class _P:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _fake_tracklets(rng):
    out = []
    for L, views in ((3, 2), (1, 5)):
        t = _P(frame_idxs=list(range(4, 4 + L)), state=_State.Confirmed, hits=L, time_since_update=0, max_age=0, n_inits=3, poses=[],
               cam_poses_2d=[], pose_ids_2d=[])
        for i in range(L):
            prm = _P(root=rng.normal(size=3), euler_angles=rng.normal(size=(18, 3)), bone_lens=rng.uniform(size=11))
            t.poses.append((4 + i, prm, _P(keypoints=rng.normal(size=(18, 3)), keypoints_score=np.ones((18, 1)))))
            t.cam_poses_2d.append([(v, _P(keypoints=rng.normal(size=(17, 2)), keypoints_score=rng.uniform(size=(17, 1)))) for v in range(views)])
            t.pose_ids_2d.append(list(range(10, 10 + views)))
        out.append(t)
    return out


def test_tracklets_npz_round_trip(tmp_path):
    """tracklets_io: every number of the tracklets survives the .npz side format, with and without the 2D poses; the
    rebuilt objects answer the accessors viz_tracklets uses (p[0], p[-1].keypoints, pose_type)."""
    from multiview_motion_capture_b200.tracklets_io import load_tracklets_npz, save_tracklets_npz
    tl = _fake_tracklets(np.random.default_rng(0))
    for keep in (False, True):
        save_tracklets_npz(tmp_path / "t.npz", tl, keep_2d=keep)
        back = load_tracklets_npz(tmp_path / "t.npz")
        assert [len(t) for t in back] == [3, 1]
        for a, b in zip(tl, back):
            assert a.frame_idxs == b.frame_idxs and b.state == a.state.value and b.hits == a.hits
            assert b.views == [[(v, pid) for (v, _), pid in zip(fr, ids)] for fr, ids in zip(a.cam_poses_2d, a.pose_ids_2d)]
            for pa, pb in zip(a.poses, b.poses):
                assert pa[0] == pb[0] and np.array_equal(pa[1].euler_angles, pb[1].euler_angles) and np.array_equal(pa[1].root, pb[1].root)
                assert np.array_equal(pa[-1].keypoints, pb[-1].keypoints) and pb[-1].pose_type == "BASIC_18"
            assert (b.cam_poses_2d is not None) == keep
            if keep:
                for fa, fb in zip(a.cam_poses_2d, b.cam_poses_2d):
                    for (va, qa), (vb, qb) in zip(fa, fb):
                        assert va == vb and np.array_equal(qa.keypoints, qb.keypoints) and np.array_equal(qa.keypoints_score, qb.keypoints_score)
        assert back[0].last_pose_3d is back[0].poses[-1][-1]


def test_shard_clips_round_robin():
    from multiview_motion_capture_b200 import sharding
    for n, w in ((4096, 8), (10, 3), (3, 4)):
        owned = [sharding.shard_clips(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(owned).tolist()) == list(range(n))
        for r, idx in enumerate(owned):
            assert all(sharding.clip_owner(int(c), w) == r for c in idx)
    with pytest.raises(ValueError):
        sharding.shard_clips(4, 4, 4)


def test_scene_stream_does_not_depend_on_the_sharding():
    """A clip's synthetic data depends on (seed, clip id) only: generated alone, in another batch, or by the device-side
    generator's scene set-up, clip 5 has the same cameras, people and detections."""
    from multiview_motion_capture_b200 import synthetic as S
    a = S.SceneStream(4, 8, 16, seed=5, clip_offset=2)
    b = S.SceneStream(1, 8, 16, seed=5, clip_ids=[5])
    for _ in range(3):
        da, db = a.next(), b.next()
    assert np.array_equal(da["kps25"][3], db["kps25"][0]) and np.array_equal(da["n_pose"][3], db["n_pose"][0])
    assert np.array_equal(a.K[3], b.K[0]) and np.array_equal(a.gt_params(np.ones(11))[3], b.gt_params(np.ones(11))[0])
    d = S.DeviceSceneStream(np.array([5]), 8, 16, seed=5, device="cpu")
    assert np.array_equal(d.K[0], b.K[0])
