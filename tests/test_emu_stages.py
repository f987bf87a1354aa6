"""CPU tier: the REAL kernel sources (csrc/*.cu) compiled for the fiber emulator in tests/emu, checked stage by
stage against the reference goldens on small subsets. The GPU tier (test_gpu_*.py) runs the same checks through
libmvmc.so on full sets. Nothing here is a product path."""
import numpy as np
import pytest

import stage_checks as SC
from helpers import golden

DEV = "cpu"


def test_fundamental(emu):
    SC.check_fundamental(DEV)


def test_affinity_shelf_subset(emu):
    # frame 1 = float32 no-track path; others = tracks + 2D poses (up to 4 people)
    SC.check_affinity(DEV, "shelf", [1, 2, 3, 120, 299])


def test_affinity_synthetic(emu):
    SC.check_affinity(DEV, "synth_c4p3", [1, 2, 5], Pmax=4, Tmax=8)


def test_als_shelf_subset(emu):
    SC.check_als(DEV, "shelf", [1, 9], N=64, rmax=16)


def test_assign_all_shelf_frames(emu):
    _, g = golden("shelf")
    SC.check_assign(DEV, "shelf", list(range(1, int(g["last_frame"]) + 1)))


@pytest.mark.parametrize("name,Pmax,Tmax", [("synth_c4p3", 4, 8), ("synth_c8p6", 8, 12), ("synth_c8p12", 12, 16)])
def test_assign_synthetic(emu, name, Pmax, Tmax):
    _, g = golden(name)
    SC.check_assign(DEV, name, list(range(1, int(g["last_frame"]) + 1)), Pmax=Pmax, Tmax=Tmax)


def test_triangulate(emu):
    assert SC.check_triangulate(DEV, "shelf", limit=3) <= 1e-6


def test_fk(emu):
    SC.check_fk(DEV, M=8)


def test_ik_well_posed_matches_oracle(emu):
    """On a well-conditioned parameter subset the trust-region trajectory is stable, and the kernel must follow
    the oracle's restated SciPy TRF step for step (SURVEY.md 8c' item 3)."""
    probs = SC.ik_problems("shelf", [5])[:1]
    SC.check_ik_well_posed(DEV, probs, nfevs=(5,))


def test_pipeline_teacher_forced_two_frames(emu):
    """mvmc_clips_* end to end (prepare -> affinity -> ALS -> assign -> gather -> IK -> commit) on two tracking frames
    of a synthetic golden scene, 2 replicated clips: identical X_bin / matches / ids / lifecycle, replicas identical."""
    from pipeline_checks import run_golden_clip
    st = run_golden_clip(DEV, "synth_c4p3", 4, 8, forced=True, B=2, max_new=4, frames=[2, 6])
    assert st["xbin"] == st["alive"] == st["upd"] == st["iters"] == st["frames"] == 2
    assert st["replicas"] == 2
    assert np.median(st["dj"]) < 2e-2


def test_clip_streams_equal_clip_batch(emu):
    """The grouped host-side runner (ClipStreams) returns the records of ClipBatch, clip for clip."""
    from pipeline_checks import check_clip_streams_equal_clip_batch
    check_clip_streams_equal_clip_batch(DEV, B=2, frames=(2,))


def test_ik_3d_target_variants(emu):
    """SURVEY.md 8f-4 through the emulator: two births and two updates of tests/golden/ik3d_ref.npz."""
    print(SC.check_ik_targets(DEV, limit=4))


def test_linear_sum_assignment(emu):
    SC.check_lsap(DEV)


def test_alternative_matchers(emu):
    """SURVEY.md 8f-3 through the emulator on the Shelf records."""
    SC.check_alt_matchers(DEV, limit=6)


def test_als_tile_builds_agree(emu):
    """Both tile builds of k_als through the emulator (forced): the reference's X_bin and stopping iteration."""
    try:      # (the dispatcher's choice for these sizes - the 64 x 96 build - is what every other test runs)
        emu.mvmc_als_force_variant(1)
        assert SC.check_als(DEV, "shelf", [9], N=64, rmax=16) == 1
    finally:
        emu.mvmc_als_force_variant(-1)


def test_edge_cases_against_the_oracle(emu):
    """Empty / single-view / ragged / all-filtered frames and the return to the no-track path after every track died."""
    from pipeline_checks import check_edge_cases
    print(check_edge_cases(DEV, full=False))     # (the GPU tier also runs the all-filtered frame and the second rebirth)


def test_birth_from_many_poses(emu):
    """The slow path behind MVMC_MAX_SEL (24 and 30 poses per group) through the emulator, with a small evaluation budget."""
    print(SC.check_birth_from_many_poses(DEV, n_frames=3, max_nfev=12))


def test_assign_many_pose_groups(emu):
    SC.check_assign_many_pose_groups(DEV)
