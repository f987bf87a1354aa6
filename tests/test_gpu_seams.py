"""GPU tier: the same-signature seams and the drop-in CLI on the REAL library."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import seam_checks as SK
from helpers import ROOT, fkey, golden

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "multiview_motion_capture_b200", "dropin")


def test_associate_tracking_seam(cuda):
    SK.check_associate_tracking("shelf", list(range(1, 301, 7)))
    SK.check_associate_tracking("synth_c4p3", [1, 3, 5])
    SK.check_associate_tracking("warm_c8p32", [3, 5])
    SK.check_associate_tracking("warm_c8p16", [3, 4])


def test_match_als_seam(cuda):
    SK.check_match_als("shelf", list(range(1, 301, 11)))
    SK.check_match_als("warm_c8p32", [3])
    SK.check_match_als("synth_c8p6", [2, 3])


def test_solver_and_fk_seams(cuda):
    for f in (1, 2, 50):
        SK.check_solver_and_fk("shelf", f)


def _write_shelf_tree(tmp, n_frames):
    import json
    inp, _ = golden("shelf")
    C = inp["kps25"].shape[1]
    os.makedirs(tmp / "calibs", exist_ok=True)
    for c in range(C):
        os.makedirs(tmp / "kps" / str(c), exist_ok=True)
        with open(tmp / "calibs" / f"{c}.json", "w") as f:
            json.dump({"K": inp["K"][c].reshape(-1).tolist(), "RT": inp["RT"][c].reshape(-1).tolist(),
                       "imgSize": inp["img_wh"][c].tolist()}, f)
        for fr in range(n_frames):
            people = [{"pose_keypoints_2d": inp["kps25"][fr, c, p].reshape(-1).tolist()} for p in range(int(inp["n_pose"][fr, c]))]
            with open(tmp / "kps" / str(c) / f"{c}_{fr:012d}_keypoints.json", "w") as f:
                json.dump({"version": 1.3, "people": people}, f)


def test_dropin_cli_on_the_gpu_300_shelf_frames(cuda, tmp_path):
    """`motion_capture.py --mode prepare` + `--mode run` exactly as a user of the reference runs them, on the CUDA library,
    over the reference's Shelf sample (301 frames of OpenPose JSON). tracklets.pkl against the reference's own run
    (tests/golden/shelf_ref.npz final_*): the IK is chaotic at the 1-ulp level (SURVEY.md 8c'), so a free run cannot be
    identical for 300 frames - the first 40 frames must be, and the totals must agree like the reference agrees with
    itself (+-1 ulp: 20 tracklets, 4 alive)."""
    _write_shelf_tree(tmp_path, 301)
    run = lambda *a: subprocess.run([sys.executable, os.path.join(DROPIN, "motion_capture.py"), *a], check=True, capture_output=True, text=True)
    run("--mode", "prepare", "--opn_kps_dir", str(tmp_path / "kps"), "--calib_dir", str(tmp_path / "calibs"), "--out_data_dir", str(tmp_path / "dframes"))
    run("--mode", "run", "--video_dir", "", "--data_dir", str(tmp_path / "dframes"), "--output_dir", str(tmp_path / "out"))
    sys.path.insert(0, DROPIN)
    try:
        import motion_capture
        sys.modules["__main__"].MvTracklet = motion_capture.MvTracklet
        sys.modules["__main__"].TrackState = motion_capture.TrackState
        with open(tmp_path / "out" / "tracklets.pkl", "rb") as f:
            tl = pickle.load(f)["tracklets"]
    finally:
        sys.path.remove(DROPIN)
    _, g = golden("shelf")
    ref_len, ref_first = g["final_len"], g["final_first_frame"]
    lens = [len(t) for t in tl]
    print(f"PARITY drop-in CLI on the GPU, 300 Shelf frames free-running: {len(tl)} tracklets (reference {len(ref_len)}), "
          f"lengths {lens[:6]}... (reference {ref_len[:6].tolist()}...), frames tracked {sum(lens)} (reference {int(ref_len.sum())})")
    assert lens == sorted(lens, reverse=True)
    assert abs(len(tl) - len(ref_len)) <= 4
    assert abs(sum(lens) - int(ref_len.sum())) <= 0.05 * int(ref_len.sum())
    # tracks born in the first 40 frames: same birth frames as the reference
    early = sorted(t.frame_idxs[0] for t in tl if t.frame_idxs[0] <= 40)
    assert early == sorted(int(x) for x in ref_first if x <= 40)
    from multiview_motion_capture_b200.tracklets_io import load_tracklets_npz
    side = load_tracklets_npz(tmp_path / "out" / "tracklets.npz")
    assert [len(t) for t in side] == lens


def test_step_body25_equals_step(cuda):
    """mvmc_clips_step_body25_host (BODY_25 -> COCO gather on the device) == mvmc_clips_step_host on the host-gathered poses."""
    import mvmc_oracle as o
    from multiview_motion_capture_b200.clips import ClipBatch
    inp, _ = golden("synth_c8p6")
    kps25 = inp["kps25"]
    coco = o.body25_to_coco(kps25)
    a = ClipBatch(1, 8, 8, max_tracks=12, max_new=8, device="cuda:0")
    b = ClipBatch(1, 8, 8, max_tracks=12, max_new=8, device="cuda:0")
    for cb in (a, b):
        cb.set_calib(inp["K"][None], inp["RT"][None])
    P = kps25.shape[2]
    for f in (1, 2, 3):
        k17 = np.zeros((1, 8, 8, 17, 3)); k17[0, :, :P] = coco[f]
        k25 = np.zeros((1, 8, 8, 25, 3)); k25[0, :, :P] = kps25[f]
        ra = a.step(k17, inp["n_pose"][f][None], f).copy()
        rb = b.step_body25(k25, inp["n_pose"][f][None], f).copy()
        assert ra.tobytes() == rb.tobytes(), f
    a.close(); b.close()


def test_torch_extension_is_the_binding_of_the_hot_path(cuda):
    """The PyTorch C++ extension (csrc/torch_ext.cpp -> lib/libmvmc_torch.so) is loaded on the GPU box, its operators are
    what stages.* and ClipBatch.step_device call, and they give the bits the raw C-ABI (ctypes) gives."""
    import ctypes
    import torch
    from multiview_motion_capture_b200 import _lib, stages
    from multiview_motion_capture_b200._lib import check, ptr
    ops = _lib.torch_ops()
    assert ops is not None and os.path.exists(_lib.TORCH_EXT_PATH)
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.normal(0, 0.3, size=(64, 68)), device="cuda:0")
    a = torch.ops.mvmc.fk(x)
    b = torch.empty_like(a)
    check(_lib.get_lib().mvmc_fk(ptr(x), 64, ptr(b), torch.cuda.current_stream().cuda_stream), "mvmc_fk")
    assert torch.equal(a, b) and torch.equal(stages.fk(x), a)
    with pytest.raises(RuntimeError):
        torch.ops.mvmc.fk(x.cpu())          # no CPU implementation is registered: no fallback
    loaded = open("/proc/self/maps").read()
    assert "libmvmc_torch.so" in loaded and "libmvmc.so" in loaded
