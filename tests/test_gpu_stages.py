"""GPU tier: every C-ABI stage of libmvmc.so against the reference goldens / the oracle, full sets.
Bars (BASELINE.json north_star): association bit-exact; triangulation <= 1 mm; IK 1e-3 rad where the reference
itself is reproducible (well-posed subsets), inside the reference's own 1-ulp noise envelope elsewhere."""
import numpy as np
import pytest

import stage_checks as SC
from helpers import WARM, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SYNTH = [("synth_c4p3", 4, 8), ("synth_c8p6", 8, 12), ("synth_c8p12", 12, 16)]


def _frames(name):
    _, g = golden(name)
    return list(range(int(g["first_frame"]), int(g["last_frame"]) + 1))


def test_fundamental(cuda):
    SC.check_fundamental(DEV)


def test_affinity_shelf_all_frames(cuda):
    worst = SC.check_affinity(DEV, "shelf", _frames("shelf"))
    print("max |dst - reference| px:", worst)


@pytest.mark.parametrize("name,Pmax,Tmax", SYNTH)
def test_affinity_synthetic(cuda, name, Pmax, Tmax):
    SC.check_affinity(DEV, name, _frames(name), Pmax=Pmax, Tmax=Tmax)


def test_als_shelf_all_frames(cuda):
    """X_bin bit-identical on all 300 Shelf frames, and so is the stopping iteration."""
    fr = _frames("shelf")
    assert SC.check_als(DEV, "shelf", fr, N=64, rmax=16) == len(fr)


@pytest.mark.parametrize("name,Pmax,Tmax", SYNTH)
def test_als_synthetic(cuda, name, Pmax, Tmax):
    fr = _frames(name)
    N = -(-(Tmax + 8 * Pmax) // 32) * 32
    assert SC.check_als(DEV, name, fr, N=N, rmax=2 * max(Pmax, Tmax)) == len(fr)


@pytest.mark.parametrize("name,Pmax,Tmax", WARM)
def test_affinity_and_als_warm_goldens(cuda, name, Pmax, Tmax):
    """Stage level, on the reference's own tracked frames at the BASELINE shapes (8 x 32: n up to ~230, r = 64 - the tile
    grid and blocked inverse the benchmark runs): dst <= 1e-7 px; X_bin and the stopping iteration bit-exact."""
    fr = _frames(name)
    worst = SC.check_affinity(DEV, name, fr, Pmax=Pmax, Tmax=Tmax)
    N = -(-(Tmax + 8 * Pmax) // 32) * 32
    assert SC.check_als(DEV, name, fr, N=N, rmax=2 * max(Pmax, Tmax)) == len(fr)
    SC.check_assign(DEV, name, fr, Pmax=Pmax, Tmax=Tmax, max_new=Pmax)
    print(f"PARITY {name} stages vs reference: {len(fr)} frames, max |dst - reference| {worst:.2e} px, X_bin + iterations identical")


def test_assign_shelf_all_frames(cuda):
    SC.check_assign(DEV, "shelf", _frames("shelf"))


@pytest.mark.parametrize("name,Pmax,Tmax", SYNTH)
def test_assign_synthetic(cuda, name, Pmax, Tmax):
    SC.check_assign(DEV, name, _frames(name), Pmax=Pmax, Tmax=Tmax)


@pytest.mark.parametrize("name", ["shelf"] + [s[0] for s in SYNTH])
def test_triangulate(cuda, name):
    worst = SC.check_triangulate(DEV, name)
    print("max |linear kps3d - reference| m:", worst)
    assert worst <= 1e-6


def test_triangulate_noise_free_recovers_points(cuda):
    """Property at full size: 4096 people x 18 joints seen by 8 cameras, exact projections -> exact 3D (<= 1e-8 m)."""
    import torch
    from multiview_motion_capture_b200 import stages as S, synthetic as syn
    rng = np.random.default_rng(1)
    K, RT = syn.make_cameras(rng, 8)
    P = np.einsum("vij,vjk->vik", K, RT)
    M = 4096
    X = rng.uniform(-3, 3, (M, 18, 3))
    X[..., 2] = rng.uniform(0, 2, (M, 18))
    uvw = np.einsum("vij,mkj->mvki", P, np.concatenate([X, np.ones((M, 18, 1))], -1))
    obs = np.concatenate([uvw[..., :2] / uvw[..., 2:3], np.full((M, 8, 18, 1), 0.9)], -1)
    obs16 = np.zeros((M, 16, 18, 3))
    obs16[:, :8] = obs
    P16 = np.zeros((M, 16, 3, 4))
    P16[:, :8] = P
    out = S.triangulate(SC.T(obs16, DEV), SC.T(P16, DEV), SC.T(np.full(M, 8), DEV, torch.int32), 0.01, 0).cpu().numpy()
    assert np.abs(out[..., :3] - X).max() <= 1e-8
    assert np.allclose(out[..., 3], 0.9)


def test_fk(cuda):
    SC.check_fk(DEV, M=4096)


def test_ik_well_posed_matches_scipy_trf(cuda):
    """Literal north_star tolerance (1e-3 rad / 1 mm) where the reference's own solver is reproducible."""
    probs = SC.ik_problems("shelf", [5, 60, 150, 250]) + SC.ik_problems("synth_c8p6", [3])
    da, dj = SC.check_ik_well_posed(DEV, probs[:12])
    print(f"well-posed IK: max |d angle| = {da:.2e} rad, max |d joint| = {dj:.2e} m")


def test_ik_teacher_forced_inside_reference_noise_envelope(cuda):
    """Every tracking-mode solve the reference made on Shelf (warm start = the reference's previous parameters).
    The reference's own answer moves by millimetres / tenths of a radian when its inputs change by 1 ulp (SURVEY.md
    8c': the trust-region step is stretched along directions whose singular values are finite-difference noise), so
    the literal 1e-3 rad bar is below the reference's reproducibility. The bar here is calibrated in the test itself:
    the oracle (bit-identical to the reference) is re-run on a subsample with the projection matrices perturbed by
    1 ulp, and the CUDA-vs-reference differences must sit inside 1.5x that envelope; leaf DOFs stay put, the cost never
    rises above the warm start, and every solve spends its 5 evaluations like the reference's."""
    import mvmc_oracle as o
    probs = [p for p in SC.ik_problems("shelf", _frames("shelf")) if not p["birth"]]
    assert len(probs) > 700
    x, joints, info, cost = SC.run_ik(DEV, probs)
    dj = np.array([np.abs(joints[m] - p["j_ref"]).max() for m, p in enumerate(probs)])
    leaf = np.zeros(68, bool)
    leaf[:57] = ~SC.LEAF_PARAM_MASK[:57].astype(bool)
    dleaf = np.array([np.abs(x[m] - p["x_ref"])[leaf].max() for m, p in enumerate(probs)])
    dang = np.array([np.abs(x[m] - p["x_ref"])[3:57][~leaf[3:57]].max() for m, p in enumerate(probs)])
    skel = o.load_skeleton()

    def cost_of(p, xx):
        obs = np.array([o.add_mid_spine(k) for k in p["kps"]])[:, o.IK_OBS_IDX, :]
        return 0.5 * np.sum(o._reproj_residual(skel, obs, list(p["P"]), xx[:3], xx[3:57].reshape(18, 3), xx[57:]) ** 2)

    rel = []
    for m, p in enumerate(probs):
        c0, cr = cost_of(p, p["x0"]), cost_of(p, p["x_ref"])
        assert cost[m, 1] <= c0 * (1 + 1e-12), (m, cost[m], c0)
        rel.append(abs(cost[m, 1] - cr) / cr)
    rel = np.array(rel)
    # the reference's own envelope on every 5th solve
    rng = np.random.default_rng(0)
    sub = list(range(0, len(probs), 5))
    ej, erel, eang = [], [], []
    for m in sub:
        p = probs[m]
        P2 = [P * (1 + 2e-16 * rng.standard_normal(P.shape)) for P in p["P"]]
        prm, jb = o.solve_ik(skel, o.PoseParam.unpack(p["x0"]), list(p["kps"]), P2)
        ej.append(np.abs(jb - p["j_ref"]).max())
        xb = prm.pack()
        eang.append(np.abs(xb - p["x_ref"])[3:57][~leaf[3:57]].max())
        erel.append(abs(cost_of(p, xb) - cost_of(p, p["x_ref"])) / cost_of(p, p["x_ref"]))
    ej, erel, eang = np.array(ej), np.array(erel), np.array(eang)
    q = lambda a: (np.median(a), np.percentile(a, 90))
    print(f"IK vs reference over {len(probs)} solves: joints median {q(dj)[0]*1e3:.2f} mm p90 {q(dj)[1]*1e3:.2f} mm "
          f"(reference vs itself + 1 ulp: {q(ej)[0]*1e3:.2f} / {q(ej)[1]*1e3:.2f} mm); non-leaf angles median {q(dang)[0]:.3f} rad "
          f"(reference: {q(eang)[0]:.3f}); leaf angles max {dleaf.max():.1e}; rel cost median {q(rel)[0]:.1e} p90 {q(rel)[1]:.1e} "
          f"(reference: {q(erel)[0]:.1e} / {q(erel)[1]:.1e}); within 1e-3 rad: {np.mean(dang <= 1e-3)*100:.1f} % "
          f"(reference: {np.mean(eang <= 1e-3)*100:.1f} %), within 1 mm: {np.mean(dj <= 1e-3)*100:.1f} % "
          f"(reference: {np.mean(ej <= 1e-3)*100:.1f} %)")
    assert (info[:, :, 0] == 5).all()            # every tracking solve uses its 5 evaluations, as in the reference
    assert q(dj)[0] <= 1.5 * q(ej)[0] and q(dj)[1] <= 1.5 * q(ej)[1]
    assert q(dang)[0] <= 1.5 * q(eang)[0]
    assert q(rel)[0] <= 2.0 * q(erel)[0] + 1e-6 and q(rel)[1] <= 2.0 * q(erel)[1] + 1e-6
    assert dleaf.max() <= 1e-6                   # structurally unobservable DOFs stay put (reference: <= 6.4e-8)


def test_ik_3d_target_variants(cuda):
    w = SC.check_ik_targets(DEV)
    print(f"PARITY 3D-target IK (solve_pose / solve_pose_bone_lens) vs reference: births max joint diff {w['birth_joints']:.2e} m, "
          f"updates max joint diff {w['upd_joints']*1e3:.2f} mm, max relative final-cost diff {w['upd_cost']:.2e}")


def test_linear_sum_assignment(cuda):
    for seed in range(3):
        SC.check_lsap(DEV, seed)


def test_alternative_matchers(cuda):
    w = SC.check_alt_matchers(DEV)
    print(f"PARITY alternative matchers vs reference (Hungarian grouping across views at two thresholds, 3D ray association): "
          f"groups / matches identical on 9 frames incl. 8x16 and 8x32, max |ray cost diff| {w:.2e} m")


def test_ik_bitwise_equal_to_the_cpu_restatement(cuda):
    """I4, SURVEY.md 8c' protocol 3(i): CUDA == same-operation-order CPU restatement, bit for bit, on the real path."""
    tot = 0
    for name, frames in (("shelf", list(range(1, 40)) + [120, 121, 250]), ("synth_c8p6", [1, 2, 3, 4]), ("warm_c8p16", [3, 4])):
        n, nb, worst = SC.check_ik_bitwise_vs_cpu_restatement(DEV, name, frames)
        tot += n
        print(f"PARITY IK real path, CUDA vs CPU restatement (same sources, same operation order): {name}: {n} solves ({nb} births) "
              f"bit-identical (max |dx| {worst})")
    assert tot >= 140


def test_als_tile_builds_agree(cuda):
    """The two tile builds of k_als (64x96 / 8 warps and 48x48 / 4 warps; the dispatcher picks by problem and batch size):
    X_bin and stopping iteration of the reference with either, on Shelf, 8 x 12 and 8 x 16 frames."""
    from multiview_motion_capture_b200 import _lib
    lib = _lib.get_lib()
    try:
        for v in (1, 0):
            lib.mvmc_als_force_variant(v)
            assert SC.check_als(DEV, "shelf", list(range(1, 301, 13)), N=64, rmax=16) == 24
            for name, Pmax, Tmax in (("warm_c8p12", 12, 16), ("warm_c8p16", 16, 24)):
                fr = _frames(name)
                N = -(-(Tmax + 8 * Pmax) // 32) * 32
                assert SC.check_als(DEV, name, fr, N=N, rmax=2 * max(Pmax, Tmax)) == len(fr)
    finally:
        lib.mvmc_als_force_variant(-1)


def test_birth_from_many_poses(cuda):
    r = SC.check_birth_from_many_poses(DEV, n_frames=3, max_nfev=50)
    print("PARITY births from more than MVMC_MAX_SEL poses vs oracle (poses, max joint diff m, rel. cost diff, (nfev, njev, status) x2, oracle):", r)
    SC.check_assign_many_pose_groups(DEV)
