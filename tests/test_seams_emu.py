"""CPU tier: the same-signature seams of SURVEY.md 8b (associate_tracking, match_als, transform_closure, PoseSolver,
foward_kinematics) through the kernel emulator on a few Shelf frames; and the A/B swap itself: the REAL reference driven
frame by frame with our associate_tracking swapped in (container only - needs /root/reference)."""
import os
import sys

import numpy as np
import pytest

import seam_checks as SK
from helpers import ROOT, fkey, golden


def test_associate_tracking_seam(emu):
    SK.check_associate_tracking("shelf", [1, 40])
    SK.check_associate_tracking("synth_c4p3", [3])


def test_match_als_seam(emu):
    SK.check_match_als("shelf", [1, 9])


def test_solver_and_fk_seams(emu):
    SK.check_solver_and_fk("shelf", 2)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the A/B swap needs the reference sources")
def test_reference_runs_with_our_association_swapped_in(emu):
    """each stage can be swapped into the reference individually. Here the reference's MvTracker runs
    Shelf frames 1..4 with `motion_capture.associate_tracking` replaced by ours: same track ids, lifecycle and
    parameters (bit for bit: the IK is still the reference's own) as the all-reference golden run."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden as MG
    import ref_shim
    ours = SK.dropin_modules()["mc"].associate_tracking
    ref = ref_shim.load()            # re-imports the reference's modules under their names
    inp, g = golden("shelf")
    packed = {k: inp[k] for k in inp.files}
    calibs = MG.calibs_from_packed(ref, packed)
    tracker = ref.mc.MvTracker(ref.ik.load_skeleton())
    real = ref.mc.associate_tracking
    ref.mc.associate_tracking = ours
    try:
        import contextlib, io
        for f in range(1, 5):
            d_frames = MG.frames_from_packed(ref, packed, f, calibs)
            with contextlib.redirect_stdout(io.StringIO()):
                d_frames = [ref.mc.filter_bad_pose(fr, 0.01, 4, 5) for fr in d_frames]
                tracker.update_4d(f, d_frames, None)
            k = fkey(f)
            upd = [t for t in tracker.tracklets if t.frame_idxs[-1] == f]
            assert len(tracker.tracklets) == len(g[k + "alive_after"])
            assert np.array_equal(np.array([[t.state.value, t.hits, t.time_since_update, len(t)] for t in tracker.tracklets]).reshape(-1, 4),
                                  g[k + "alive_state"])
            for i, t in enumerate(upd):
                assert np.array_equal(t.poses[-1][1].root, g[k + "upd_root"][i]), f
                assert np.array_equal(t.poses[-1][1].euler_angles, g[k + "upd_euler"][i]), f
    finally:
        ref.mc.associate_tracking = real
