"""Device-agnostic parity checks of every C-ABI stage against the oracle / the reference goldens.

The same functions run in two tiers:
  * `-m "not gpu"`: through tests/emu (the real kernel sources compiled for the CPU emulator), tiny subsets;
  * `-m gpu`: through libmvmc.so on the B200, full sets.
Tolerances (BASELINE.json north_star): association bit-exact, triangulation <= 1 mm, IK see test_*_ik."""
import numpy as np
import torch

import mvmc_oracle as o
from helpers import GoldenTable, fkey, golden, golden_matches, pad_poses, view_lists

from multiview_motion_capture_b200 import stages as S
from multiview_motion_capture_b200._lib import MAX_SEL

f64, i32 = torch.float64, torch.int32


def T(x, dev, dt=f64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dt).to(dev).contiguous()


# ---------------------------------------------------------------------------------------------------
def check_fundamental(dev):
    for name in ("shelf", "synth_c8p6"):
        inp, _ = golden(name)
        K, RT = inp["K"], inp["RT"]
        Ps = np.array(o.projections(K, RT))
        C = len(Ps)
        F = S.fundamental(T(Ps[None], dev)).cpu().numpy()[0]
        F32 = S.fundamental_krt(T(K[None], dev), T(RT[None], dev)).cpu().numpy()[0]
        ref32 = o.pairwise_f_mats_krt(K, RT)
        scale = np.abs(F).max()  # F[i,i] is pure rounding noise of LAPACK's det in the reference
        for i in range(C):
            for j in range(C):
                ref = o.fundamental_from_projections(Ps[i], Ps[j])
                assert np.abs(F[i, j] - ref).max() <= 1e-11 * scale, (name, i, j)
                if i != j:
                    assert np.abs(F32[i, j] - ref32[i, j]).max() <= 2e-6 * np.abs(ref32[i, j]).max(), (name, i, j)


# ---------------------------------------------------------------------------------------------------
def _forced_inputs(name, f, Pmax, Tmax, tab):
    inp, g = golden(name)
    kps = o.body25_to_coco(inp["kps25"])
    tj = tab.joints(f)
    n = len(tj)
    assert n <= Tmax
    trk = np.zeros((1, Tmax, 18, 3))
    for i, j in enumerate(tj):
        trk[0, i] = j.reshape(18, 3)
    return pad_poses(kps[f], Pmax)[None], inp["n_pose"][f][None].astype(np.int32), np.array([n], np.int32), trk


def check_affinity(dev, name, frames, Pmax=8, Tmax=24):
    """prepare + affinity with the reference's own track poses: kept poses and index layout identical,
    dst <= 1e-7 px, sim <= 1e-9; the float32 no-track path bit-exact."""
    inp, g = golden(name)
    Ps = np.array(o.projections(inp["K"], inp["RT"]))
    C = len(Ps)
    P = T(Ps[None], dev)
    F = S.fundamental(P)
    F32 = S.fundamental_krt(T(inp["K"][None], dev), T(inp["RT"][None], dev))
    tab = GoldenTable(g)
    worst = 0.0
    for f in frames:
        k = fkey(f)
        kps, n_pose, n_trk, trk = _forced_inputs(name, f, Pmax, Tmax, tab)
        prep = S.prepare(T(kps, dev), T(n_pose, dev, i32), T(n_trk, dev, i32), Tmax)
        keep = prep["keep"].cpu().numpy()[0][:, :g[k + "kept"].shape[1]]
        assert np.array_equal(keep, g[k + "kept"]), (name, f)
        dg = prep["dim_groups"].cpu().numpy()[0]
        gd = g[k + "dim_groups"]
        if n_trk[0] > 0:
            assert np.array_equal(dg, gd), (name, f, dg, gd)
        else:  # the no-track path has no (empty) track group in the reference's dim_groups
            assert np.array_equal(dg[1:], gd), (name, f, dg, gd)
        dst, sim = S.affinity(T(kps, dev), P, F, F32, T(trk, dev), T(n_trk, dev, i32), prep)
        n = int(dg[-1])
        dst = dst.cpu().numpy()[0][:n, :n]
        sim = sim.cpu().numpy()[0][:n, :n]
        assert dst.shape == g[k + "dst"].shape
        dd = np.abs(dst - g[k + "dst"]).max()
        ds = np.abs(sim - g[k + "sim"].astype(np.float64)).max()
        worst = max(worst, dd)
        assert dd <= 1e-7 and ds <= 1e-9, (name, f, dd, ds)
        if n_trk[0] == 0:
            # the reference's no-track path is float32 NumPy (mean / std by pairwise summation, NumPy's own float32 exp):
            # restated on the device bit for bit (csrc/affinity.cu np32)
            assert g[k + "sim"].dtype == np.float32
            assert np.array_equal(dst.astype(np.float32), g[k + "dst"]), (name, f, "float32 dst")
            assert np.array_equal(sim.astype(np.float32), g[k + "sim"]), (name, f, "float32 sim")
    return worst


# ---------------------------------------------------------------------------------------------------
def _als_problem(g, f, N, Gp1):
    k = fkey(f)
    sim = np.asarray(g[k + "sim"])
    n = sim.shape[0]
    W = np.zeros((N, N))
    W[:n, :n] = sim.astype(np.float64)
    dg = np.asarray(g[k + "dim_groups"], dtype=np.int32)
    is32 = sim.dtype == np.float32
    if len(dg) < Gp1:  # no-track frame: prepend the empty track group
        dg = np.concatenate([[0], dg]).astype(np.int32)
    assert len(dg) == Gp1
    return W, dg, n, is32


def check_als(dev, name, frames, N=64, rmax=48, exact_iters=True):
    """match_als on the reference's own similarity matrices: X_bin bit-identical (and the stopping iteration)."""
    inp, g = golden(name)
    C = len(inp["K"])
    probs = [_als_problem(g, f, N, C + 2) for f in frames]
    sim = T(np.stack([p[0] for p in probs]), dev)
    dg = T(np.stack([p[1] for p in probs]), dev, i32)
    f32 = T(np.array([int(p[3]) for p in probs]), dev, i32)
    xbin, n_iter = S.match_als(sim, dg, rmax, f32_first_iter=f32)
    n_iter = n_iter.cpu().numpy()
    same_it = 0
    for b, f in enumerate(frames):
        n = probs[b][2]
        xb = S.unpack_xbin(xbin[b], n)
        assert np.array_equal(xb, g[fkey(f) + "xbin"].astype(bool)), (name, f, "X_bin")
        same_it += int(n_iter[b] == int(g[fkey(f) + "als_iters"]))
        if exact_iters:
            assert n_iter[b] == int(g[fkey(f) + "als_iters"]), (name, f, n_iter[b], int(g[fkey(f) + "als_iters"]))
    return same_it


# ---------------------------------------------------------------------------------------------------
def _decode_assign(out, b, T_):
    nsel = out["trk_nsel"][b]
    trk = {}
    for t in range(T_):
        if nsel[t] >= 0:
            trk[t] = [tuple(x) for x in out["trk_sel"][b, t, :nsel[t]].tolist()]
    new = [[tuple(x) for x in out["new_sel"][b, k, :out["new_nsel"][b, k]].tolist()] for k in range(out["new_n"][b])]
    return trk, new


def check_assign(dev, name, frames, Pmax=8, Tmax=24, max_new=16):
    """closure quirk + parse + decode from the reference's own X_bin: matches identical to the reference."""
    inp, g = golden(name)
    kpsall = o.body25_to_coco(inp["kps25"])
    C = len(inp["K"])
    N = Tmax + C * Pmax
    NW = (N + 31) // 32
    B = len(frames)
    xb = np.zeros((B, N, NW * 32), dtype=np.uint8)
    kps = np.zeros((B, C, Pmax, 17, 3))
    n_pose = np.zeros((B, C), np.int32)
    n_trk = np.zeros(B, np.int32)
    for b, f in enumerate(frames):
        k = fkey(f)
        x = g[k + "xbin"]
        xb[b, :x.shape[0], :x.shape[0]] = x
        kps[b] = pad_poses(kpsall[f], Pmax)
        n_pose[b] = inp["n_pose"][f]
        n_trk[b] = len(g[k + "alive_before"])
    words = (xb.reshape(B, N, NW, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(-1).astype(np.uint32)
    prep = S.prepare(T(kps, dev), T(n_pose, dev, i32), T(n_trk, dev, i32), Tmax)
    out = S.assign(T(words.view(np.int32), dev, i32), prep, T(n_trk, dev, i32), C, max_new)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert (out["err"] == 0).all()
    for b, f in enumerate(frames):
        trk, new = _decode_assign(out, b, int(n_trk[b]))
        tm, ng = golden_matches(g, f)
        assert trk == tm, (name, f, trk, tm)
        assert new == [s_ for s_ in ng if len(s_) >= 2], (name, f, new, ng)   # only groups that get born are stored
        assert out["counts"][b, 0] == int(g[fkey(f) + "printed"]), (name, f)
        assert out["counts"][b, 1] == sum(len(s_) < 2 for s_ in ng), (name, f)
        assert out["counts"][b, 2] == 0


# ---------------------------------------------------------------------------------------------------
def check_triangulate(dev, name, limit=None):
    """DLT (+ the 2-nfev refine) on every triangulation call the reference made: <= 1e-6 m (spec: 1 mm)."""
    _, g = golden(name)
    n = int(g["tri_count"])
    idx = list(range(n))[:limit]
    V = MAX_SEL
    obs = np.zeros((len(idx), V, 18, 3))
    Ps = np.zeros((len(idx), V, 3, 4))
    nv = np.zeros(len(idx), np.int32)
    for q, i in enumerate(idx):
        P, pts = g[f"tri{i}_P"], g[f"tri{i}_pts"]
        nv[q] = len(P)
        obs[q, :len(P)] = pts
        Ps[q, :len(P)] = P
    lin = S.triangulate(T(obs, dev), T(Ps, dev), T(nv, dev, i32), 0.01, 0).cpu().numpy()
    out = S.triangulate(T(obs, dev), T(Ps, dev), T(nv, dev, i32), 0.01, 2).cpu().numpy()
    worst = 0.0
    for q, i in enumerate(idx):
        d0 = np.abs(lin[q] - g[f"tri{i}_linear"]).max()
        d1 = np.abs(out[q] - g[f"tri{i}_out"]).max()
        # The refine is a single trust-region trial step. When the reference's step is a rank-deficient pseudo-inverse
        # direction stretched to the radius (2-view births: metres long, SciPy ends its alpha iteration at ~0), J^T J
        # resolves that direction to ~0.5 % of the step; otherwise the step is rejected or tiny and 1e-6 m holds.
        step = np.abs(g[f"tri{i}_out"][:, :3] - g[f"tri{i}_linear"][:, :3]).max()
        worst = max(worst, d0)
        assert d0 <= 1e-6, (name, i, d0)
        assert d1 <= 1e-6 + 1e-2 * step, (name, i, d1, step)
    return worst


# ---------------------------------------------------------------------------------------------------
def check_fk(dev, M=64, seed=0):
    """FK kernel vs the oracle: <= 1e-12 m; generic chain kernel vs chain_fk on the CMU-31 topology."""
    rng = np.random.default_rng(seed)
    skel = o.load_skeleton()
    x = np.zeros((M, 68))
    x[:, :3] = rng.normal(0, 2, (M, 3))
    x[:, 3:57] = rng.uniform(-1.5, 1.5, (M, 54))
    x[:, 57:] = skel.side_bone_lens * rng.uniform(0.8, 1.2, (M, 11))
    J = S.fk(T(x, dev)).cpu().numpy()
    for m in range(M):
        ref, _ = o.forward_kinematics(skel, x[m, :3], x[m, 3:57].reshape(18, 3), x[m, 57:])
        assert np.abs(J[m] - ref).max() <= 1e-12, m
    # CMU skeleton (skeleton_CMU.yml topology: 31 joints)
    parents = np.array([-1, 0, 1, 2, 3, 4, 0, 6, 7, 8, 9, 0, 11, 12, 13, 14, 15, 13, 17, 18, 19, 20, 21, 20, 13, 24, 25, 26,
                        27, 28, 27], np.int32)
    Jn = len(parents)
    offs = rng.normal(0, 0.2, (Jn, 3))
    from scipy.spatial.transform import Rotation
    rot = Rotation.random(M * Jn, random_state=1).as_matrix().reshape(M, Jn, 3, 3)
    root = rng.normal(0, 1, (M, 3))
    got = S.fk_chain(T(rot, dev), T(offs, dev), T(parents, dev, i32), T(root, dev)).cpu().numpy()
    for m in range(0, M, 7):
        ref = o.chain_fk(offs, parents, rot[m], root[m])
        assert np.abs(got[m] - ref).max() <= 1e-12


# ---------------------------------------------------------------------------------------------------
def ik_problems(name, frames):
    """Teacher-forced IK updates: for every track the reference updated at frame f, its previous parameters (warm start),
    the 2D poses / projections the reference used, and the reference's answer."""
    inp, g = golden(name)
    kps = o.body25_to_coco(inp["kps25"])
    Ps = np.array(o.projections(inp["K"], inp["RT"]))
    tab = GoldenTable(g)
    out = []
    for f in frames:
        k = fkey(f)
        ids_before = tab.seek(f)
        tm, ng = golden_matches(g, f)
        upd = g[k + "upd_ids"].tolist()
        for t_idx, sel in tm.items():
            if len(sel) < 2:
                continue
            tid = ids_before[t_idx]
            u = upd.index(tid)
            out.append(dict(frame=f, birth=False, x0=tab.table[tid]["param"].copy(), sel=sel,
                            kps=np.array([kps[f, v, p] for v, p in sel]), P=np.array([Ps[v] for v, _ in sel]),
                            x_ref=np.concatenate([g[k + "upd_root"][u], g[k + "upd_euler"][u].reshape(-1), g[k + "upd_blens"][u]]),
                            j_ref=g[k + "upd_joints"][u]))
        n_new = 0
        born = [t for t in upd if t not in ids_before]
        for sel in ng:
            if len(sel) < 2:
                continue
            tid = born[n_new]
            n_new += 1
            u = upd.index(tid)
            out.append(dict(frame=f, birth=True, x0=np.zeros(68), sel=sel,
                            kps=np.array([kps[f, v, p] for v, p in sel]), P=np.array([Ps[v] for v, _ in sel]),
                            x_ref=np.concatenate([g[k + "upd_root"][u], g[k + "upd_euler"][u].reshape(-1), g[k + "upd_blens"][u]]),
                            j_ref=g[k + "upd_joints"][u]))
    return out


def run_ik(dev, probs, free_mask=None, max_nfev=None):
    M, V = len(probs), MAX_SEL
    kps = np.zeros((M, V, 17, 3))
    Ps = np.zeros((M, V, 3, 4))
    nv = np.zeros(M, np.int32)
    x0 = np.zeros((M, 68))
    birth = np.zeros(M, np.uint8)
    nfev = np.zeros(M, np.int32)
    for m, p in enumerate(probs):
        v = len(p["sel"])
        kps[m, :v], Ps[m, :v], nv[m], x0[m], birth[m] = p["kps"], p["P"], v, p["x0"], int(p["birth"])
        nfev[m] = (50 if p["birth"] else 5) if max_nfev is None else max_nfev
    fm = T(free_mask, dev, torch.uint8) if free_mask is not None else None
    x, joints, info, cost = S.ik_solve(T(kps, dev), T(Ps, dev), T(nv, dev, i32), T(x0, dev), T(birth, dev, torch.uint8),
                                       T(nfev, dev, i32), fm)
    return x.cpu().numpy(), joints.cpu().numpy(), info.cpu().numpy(), cost.cpu().numpy()


LEAF_PARAM_MASK = np.ones(68, np.uint8)
for _j in (3, 6, 11, 14, 16, 17):  # leaf joints' Euler angles are structurally unobservable
    LEAF_PARAM_MASK[3 + 3 * _j: 6 + 3 * _j] = 0


def oracle_ik(p, max_nfev, free=None):
    """The oracle's solve for problem p with an optional free-parameter subset (same two-stage structure)."""
    skel = o.load_skeleton()
    obs = np.array([o.add_mid_spine(k) for k in p["kps"]])[:, o.IK_OBS_IDX, :]
    Ps = list(p["P"])
    x = p["x0"].copy()
    res = []
    for stage, npar in ((0, 57), (1, 68)):
        act = np.array([i for i in range(npar) if free is None or free[i]])
        lens_fixed = x[57:].copy()

        def fun(z, act=act, npar=npar, lens_fixed=lens_fixed, base=x.copy()):
            y = base.copy()
            y[act] = z
            return o._reproj_residual(skel, obs, Ps, y[:3], y[3:57].reshape(18, 3), y[57:] if npar == 68 else lens_fixed)

        r = o.trf_least_squares(fun, x[act].copy(), max_nfev)
        x[act] = r.x
        res.append(r)
    joints, _ = o.forward_kinematics(skel, x[:3], x[3:57].reshape(18, 3), x[57:])
    return x, joints, res


def well_posed_mask(p, rel=2e-3):
    """Free-parameter mask [68] keeping only parameters the observations of problem p determine well: QR with column
    pivoting on the oracle's forward-difference Jacobian at x0, columns with |R_ii| > rel |R_00|. On this subset the
    trust-region trajectory is stable (no bifurcation, SURVEY.md §8c'), so the kernel can be held to the literal
    north_star tolerance (1e-3 rad, 1 mm) against the oracle's restated SciPy TRF."""
    import scipy.linalg as sl
    skel = o.load_skeleton()
    obs = np.array([o.add_mid_spine(k) for k in p["kps"]])[:, o.IK_OBS_IDX, :]
    fun = lambda x: o._reproj_residual(skel, obs, list(p["P"]), x[:3], x[3:57].reshape(18, 3), x[57:])
    x0 = p["x0"]
    J = o.fd_jacobian(fun, x0, fun(x0))
    _, R, piv = sl.qr(J, pivoting=True, mode="economic")
    d = np.abs(np.diag(R))
    mask = np.zeros(68, np.uint8)
    mask[piv[d > rel * d[0]]] = 1
    return mask


def check_ik_well_posed(dev, probs, nfevs=(5, 30)):
    """Kernel TRF == oracle TRF (== SciPy, test_oracle_golden) on well-posed subsets: identical (nfev, njev, status),
    angles <= 1e-3 rad (measured ~1e-6), joints <= 1e-5 m."""
    worst = (0.0, 0.0)
    for p in probs:
        mask = well_posed_mask(p)
        for nf in nfevs:
            x, joints, info, cost = run_ik(dev, [p], free_mask=mask, max_nfev=nf)
            xr, jr, res = oracle_ik(p, nf, free=mask)
            got = [tuple(info[0, s, :3].tolist()) for s in range(2)]
            ref = [(r.nfev, r.njev, r.status) for r in res]
            if nf <= 8:
                assert got == ref, (p["frame"], nf, got, ref)
            else:  # long convergent runs: the ftol test may fire one evaluation earlier or later
                assert all(a[2] == b[2] and abs(a[0] - b[0]) <= 2 for a, b in zip(got, ref)), (p["frame"], nf, got, ref)
            da, dj = np.abs(x[0] - xr).max(), np.abs(joints[0] - jr).max()
            assert da <= 1e-3 and dj <= 1e-5, (p["frame"], nf, da, dj)
            assert abs(cost[0, 1] - res[1].cost) <= 1e-6 * max(1.0, res[1].cost)
            worst = (max(worst[0], da), max(worst[1], dj))
    return worst


# ---------------------------------------------------------------------------------------------------
def check_ik_targets(dev, limit=None):
    """solve_pose / solve_pose_bone_lens (3D-target IK) on the records the REAL reference produced
    (tests/golden/ik3d_ref.npz, oracle/make_golden.py ik3d): the kernel's per-stage and chained results against the
    reference. These problems are better posed than the reprojection ones (48 residuals straight on joint positions):
    births converge (status 2) and agree to micro-metres; the 5-evaluation updates take the same number of evaluations and
    stay within the trust-region noise of a rank-deficient step (48 rows, 57 / 68 columns)."""
    import os
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, "ik3d_ref.npz"))
    n = int(g["count"])
    idx = list(range(n))[:limit]
    tgt = np.stack([g[f"r{i}_obs3d"][g[f"r{i}_obs_idx"]] for i in idx])
    x0 = np.stack([g[f"r{i}_x0"] for i in idx])
    x1r = np.stack([g[f"r{i}_x1"] for i in idx])
    cap = np.array([int(g[f"r{i}_nfev_cap"]) for i in idx], np.int32)
    for i in idx:
        assert g[f"r{i}_obs_idx"].tolist() == o.IK_OBS_IDX.tolist() and g[f"r{i}_skel_idx"].tolist() == o.IK_SKEL_IDX.tolist()
    both = [a.cpu().numpy() for a in S.ik_solve_targets(T(tgt, dev), T(x0, dev), T(cap, dev, i32), 3)]
    st1 = [a.cpu().numpy() for a in S.ik_solve_targets(T(tgt, dev), T(x0, dev), T(cap, dev, i32), 1)]
    st2 = [a.cpu().numpy() for a in S.ik_solve_targets(T(tgt, dev), T(x1r, dev), T(cap, dev, i32), 2)]   # from the reference's stage 1
    worst = dict(birth_joints=0.0, upd_joints=0.0, upd_cost=0.0)
    for q, i in enumerate(idx):
        meta, cost = g[f"r{i}_meta"], g[f"r{i}_cost"]
        birth = bool(g[f"r{i}_birth"])
        # stage 1 alone, and stage 2 alone from the reference's stage-1 result: evaluation counts and status of the reference
        # (a 50-evaluation birth converges - status 2 - after a few evaluations more or less than SciPy; what it converges
        #  to is compared below. A 5-evaluation update spends its whole budget: nfev and status identical.)
        for got, ref_m in ((st1[2][q, 0], meta[0]), (st2[2][q, 1], meta[1])):
            if birth:     # converged like the reference (SciPy status 1..4: gtol / ftol / xtol; which test fires first can differ)
                assert got[2] > 0 and ref_m[2] > 0, (i, got, ref_m)
            else:
                assert got[2] == ref_m[2] and got[0] == ref_m[0], (i, got, ref_m)
        assert st1[2][q, 1, 0] == 0 and st2[2][q, 0, 0] == 0                   # the other stage did not run
        assert np.array_equal(st1[0][q, 57:], x0[q, 57:])                        # solve_pose leaves the bone lengths alone
        assert abs(st1[3][q, 0] - cost[0]) <= (1e-6 if birth else 0.5) * cost[0] + 1e-12, (i, st1[3][q, 0], cost[0])
        dj = np.abs(both[1][q] - g[f"r{i}_joints"]).max()
        if birth:
            worst["birth_joints"] = max(worst["birth_joints"], dj)
            assert dj <= 1e-5, (i, dj)
        else:
            worst["upd_joints"] = max(worst["upd_joints"], dj)
            worst["upd_cost"] = max(worst["upd_cost"], abs(both[3][q, 1] - cost[1]) / cost[1])
            # (rank-deficient 5-evaluation steps: centimetre-level trust-region noise, as on the reprojection path; the second
            #  stage never raises the first stage's cost and ends in the neighbourhood of the reference's)
            assert dj <= 2e-2 and both[3][q, 1] <= both[3][q, 0] * (1 + 1e-9) and both[3][q, 1] <= 2.0 * cost[0], (i, dj, both[3][q], cost)
    return worst


# ---------------------------------------------------------------------------------------------------
def _alt_records():
    import os
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, "altmatch_ref.npz"))
    return g, [dict(scene=str(g[f"r{i}_scene"]), frame=int(g[f"r{i}_frame"]), i=i) for i in range(int(g["count"]))]


def check_lsap(dev, seed=0):
    """mvmc_linear_sum_assignment against scipy.optimize.linear_sum_assignment: random rectangular problems (assignment and
    cost identical), ties (cost identical, assignment valid), NaN (status -1, SciPy raises)."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(seed)
    shapes = [(1, 1), (3, 5), (5, 3), (8, 8), (17, 32), (40, 9), (64, 64), (64, 200), (230, 30)]
    for R, Cc in shapes:
        B = 4
        cost = rng.uniform(0, 100, size=(B, R, Cc))
        cost[1] = np.round(cost[1] / 25)            # many exact ties
        nr = np.array([R, R, max(1, R - 1), R], np.int32)
        nc = np.array([Cc, Cc, Cc, max(1, Cc - 2)], np.int32)
        col, status = S.linear_sum_assignment(T(cost, dev), T(nr, dev, i32), T(nc, dev, i32))
        col, status = col.cpu().numpy(), status.cpu().numpy()
        assert (status == 0).all(), (R, Cc, status)
        for b in range(B):
            sub = cost[b, :nr[b], :nc[b]]
            rows, cols = linear_sum_assignment(sub)
            got_rows = np.nonzero(col[b, :nr[b]] >= 0)[0]
            got_cols = col[b, got_rows]
            assert len(got_rows) == min(nr[b], nc[b]) and len(set(got_cols.tolist())) == len(got_cols), (R, Cc, b)
            assert abs(sub[got_rows, got_cols].sum() - sub[rows, cols].sum()) <= 1e-9 * max(1.0, abs(sub[rows, cols].sum())), (R, Cc, b)
            if b != 1:
                assert np.array_equal(got_rows, rows) and np.array_equal(got_cols, cols), (R, Cc, b)
            assert (col[b, nr[b]:] == -1).all()
    bad = rng.uniform(size=(1, 4, 6))
    bad[0, 2, 3] = np.nan
    _, status = S.linear_sum_assignment(T(bad, dev))
    assert int(status.cpu().numpy()[0]) == -1


def check_alt_matchers(dev, limit=None):
    """SURVEY.md 8f-3 against what the REAL reference computed (tests/golden/altmatch_ref.npz): the view-by-view Hungarian
    grouping (match_objects_across_views) at two thresholds - identical groups in identical order, or the same failure where
    SciPy raises on a NaN cost - and the 3D ray association (tracklet_to_poses_association): costs <= 1e-9 m, identical matches."""
    g, recs = _alt_records()
    assert g["common_coco"].tolist() == o.RAY_COCO.tolist() and g["common_b18"].tolist() == o.RAY_B18.tolist()
    worst = 0.0
    for r in recs[:limit]:
        inp, gg = golden(r["scene"])
        f, i = r["frame"], r["i"]
        kps = o.body25_to_coco(inp["kps25"])
        C, Pm = kps.shape[1:3]
        Pmax = max(Pm, 1)
        tab = GoldenTable(gg)
        tj = tab.joints(f)
        Tn = len(tj)
        Tmax = max(Tn, 1)
        kp, n_pose, n_trk, trk = _forced_inputs(r["scene"], f, Pmax, Tmax, tab)
        Ps = np.array(o.projections(inp["K"], inp["RT"]))
        prep0 = S.prepare(T(kp, dev), T(n_pose, dev, i32), T(np.zeros(1, np.int32), dev, i32), 0)     # poses only: T = 0
        D = S.distances(T(kp, dev), T(Ps[None], dev), S.fundamental(T(Ps[None], dev)), T(np.zeros((1, 1, 18, 3)), dev),
                        T(np.zeros(1, np.int32), dev, i32), prep0)
        iv, ip = prep0["idx_view"].cpu().numpy()[0], prep0["idx_pose"].cpu().numpy()[0]
        for ti in (0, 1):
            gof, ng, status = S.match_views_hungarian(D, prep0["dim_groups"], float(g[f"r{i}_thr{ti}"]))
            gof, ng, status = gof.cpu().numpy()[0], int(ng.cpu().numpy()[0]), int(status.cpu().numpy()[0])
            if int(g[f"r{i}_raises{ti}"]):
                assert status == -1, (r, ti)
                continue
            assert status == 0, (r, ti)
            ref_rows = g[f"r{i}_groups{ti}"]
            n = int(prep0["dim_groups"].cpu().numpy()[0][-1])
            got = sorted((int(gof[q]), int(iv[q]), int(ip[q])) for q in range(n))
            assert ng == len(set(ref_rows[:, 0].tolist())), (r, ti)
            assert got == sorted(map(tuple, ref_rows.tolist())), (r, ti)
        # 3D ray association
        if Tn == 0:
            assert len(g[f"r{i}_ray_matches"]) == 0
            continue
        Kr_inv = np.stack([inp["RT"][v][:3, :3].T @ np.linalg.inv(inp["K"][v]) for v in range(C)])
        cam_loc = np.stack([-inp["RT"][v][:3, :3].T @ inp["RT"][v][:3, 3] for v in range(C)])
        prep = S.prepare(T(kp, dev), T(n_pose, dev, i32), T(n_trk, dev, i32), Tmax)
        match, cost, status = S.tracklet_pose_association(T(trk, dev), T(n_trk, dev, i32), T(kp, dev), prep["keep"], T(Kr_inv[None], dev),
                                                          T(cam_loc[None], dev), 0.1)
        match, cost = match.cpu().numpy()[0], cost.cpu().numpy()[0]
        assert (status.cpu().numpy() == 0).all()
        for v, t, p, c in g[f"r{i}_ray_costs"]:
            worst = max(worst, abs(cost[int(v), int(t), int(p)] - c))
        got = sorted((v, t, int(match[v, t])) for v in range(C) for t in range(Tn) if match[v, t] >= 0)
        assert got == sorted(map(tuple, g[f"r{i}_ray_matches"].tolist())), (r, got[:5])
    assert worst <= 1e-9, worst
    return worst


# ---------------------------------------------------------------------------------------------------
_emu_binding = None


def run_ik_cpu_restatement(probs, max_nfev=None):
    """The SAME kernel sources (csrc/ik.cu, trf_warp.cuh, det_math.cuh) compiled for the CPU (tests/emu/libmvmc_emu.so, g++
    -ffp-contract=off) and executed with the same operation order as on the GPU: the deterministic CPU restatement of
    FK + forward-difference Jacobian + TRF that SURVEY.md 8c' (protocol item 3) asks for. Called through its own ctypes
    binding with host arrays, next to whatever library the test tier has bound."""
    global _emu_binding
    import ctypes
    from helpers import EMU_LIB
    from multiview_motion_capture_b200 import _lib
    if _emu_binding is None:
        _emu_binding = _lib._bind(ctypes.CDLL(EMU_LIB))
    lib = _emu_binding
    M, V = len(probs), MAX_SEL
    kps = np.zeros((M, V, 17, 3))
    Ps = np.zeros((M, V, 3, 4))
    nv = np.zeros(M, np.int32)
    x0 = np.zeros((M, 68))
    birth = np.zeros(M, np.uint8)
    nfev = np.zeros(M, np.int32)
    for m, p in enumerate(probs):
        v = len(p["sel"])
        kps[m, :v], Ps[m, :v], nv[m], x0[m], birth[m] = p["kps"], p["P"], v, p["x0"], int(p["birth"])
        nfev[m] = (50 if p["birth"] else 5) if max_nfev is None else max_nfev
    ws = np.zeros(64)
    x, joints, info, cost = np.zeros((M, 68)), np.zeros((M, 18, 3)), np.zeros((M, 2, 4), np.int32), np.zeros((M, 2))
    ptr = lambda a: a.ctypes.data
    _lib.check(lib.mvmc_ik_solve(ptr(kps), ptr(Ps), ptr(nv), ptr(x0), ptr(birth), ptr(nfev), None, M, V, ptr(ws), ptr(x), ptr(joints),
                                 ptr(info), ptr(cost), None), "emulator mvmc_ik_solve")
    return x, joints, info, cost


def check_ik_bitwise_vs_cpu_restatement(dev, name, frames):
    """SURVEY.md 8c' protocol item 3(i): on the REAL (rank-deficient, chaotic) path - every update and birth the reference
    solved on the given frames - the CUDA kernel against the CPU restatement with the same operation order. The kernel has
    no compiler-chosen FMA contraction and its own sincos, so the bar is not 1e-3 rad but bitwise equality of parameters,
    joints, costs and (nfev, njev, status)."""
    probs = ik_problems(name, frames)
    assert len(probs) > 0
    got = run_ik(dev, probs)
    ref = run_ik_cpu_restatement(probs)
    n_birth = sum(int(p["birth"]) for p in probs)
    worst = float(np.abs(got[0] - ref[0]).max())
    for a, b, what in zip(got, ref, ("parameters", "joints", "info", "cost")):
        assert np.array_equal(a, b), (name, what, float(np.abs(np.asarray(a, dtype=np.float64) - b).max()))
    return len(probs), n_birth, worst


# ---------------------------------------------------------------------------------------------------
def check_birth_from_many_poses(dev, n_frames=3, max_nfev=50):
    """Births from more poses than MVMC_MAX_SEL (the reference builds a no-track group's track from ALL its poses,
    src/motion_capture.py:618-624, 942-958): k_ik_birth_big against the oracle's PoseSolver restatement on the same pose
    lists. Groups: one person of a golden scene seen in 8 views at `n_frames` consecutive frames = 8 * n_frames poses treated
    as that many "views" (consistent enough to converge), and the same with a second person's poses mixed in (the
    reference's crowded-scene case: an ill-posed fit, compared by cost)."""
    inp, g = golden("synth_c8p6")
    kps_all = o.body25_to_coco(inp["kps25"])
    C = kps_all.shape[1]
    gt = inp["gt_person"]
    frames = list(range(2, 2 + n_frames))
    Pmax = 2 * n_frames

    def poses_of(person, slot0):
        out, sel = {}, []
        for q, f in enumerate(frames):
            for v in range(C):
                hit = np.nonzero(gt[f, v] == person)[0]
                if len(hit):
                    out[(v, slot0 + q)] = kps_all[f, v, hit[0]]
                    sel.append((v, slot0 + q))
        return out, sel
    a, sel_a = poses_of(0, 0)
    b_, sel_b = poses_of(1, n_frames)
    kp = np.zeros((1, C, Pmax, 17, 3))
    for (v, p), k in {**a, **b_}.items():
        kp[0, v, p] = k
    Ps = np.array(o.projections(inp["K"], inp["RT"]))
    groups = [[sel_a, sorted(sel_a + sel_b[:6])]]
    assert len(sel_a) > MAX_SEL
    x, joints, info, cost = [t.cpu().numpy() for t in S.ik_birth_big(T(kp, dev), T(Ps[None], dev), groups, max_nfev)]
    skel = o.load_skeleton()
    res = []
    for gi, sel in enumerate(groups[0]):
        cam_kps = [kp[0, v, p] for v, p in sel]
        log = []
        prm, jr = o.solve_ik(skel, None, cam_kps, [Ps[v] for v, _ in sel], collect=log)
        tri = [e for e in log if e[0] == "tri"][0][1]
        r2 = [e for e in log if e[0] == "ik2"][0][2]
        dj = float(np.abs(joints[0, gi] - jr).max())
        rel = float(abs(cost[0, gi, 1] - r2.cost) / max(r2.cost, 1e-300))
        res.append((len(sel), dj, rel, info[0, gi, :, :3].tolist(), (r2.nfev, r2.status)))
    # the clean group converges to the oracle's pose; the mixed one to a comparable cost
    assert res[0][1] <= 5e-3 and res[0][2] <= 1e-2, res
    assert res[1][2] <= 0.25, res
    return res


# ---------------------------------------------------------------------------------------------------
def check_assign_many_pose_groups(dev):
    """A no-track X_bin with one group of 24 poses (3 views x 8), one of 17 and a pair: the overflow lists of
    mvmc_assign_groups hold every pose of the two large groups in the reference's member order (ascending global index),
    new_nsel carries their true sizes, nothing is counted as truncated; the oracle's decode gives the same groups."""
    C, Pmax, Tmax = 4, 8, 4
    kps = np.zeros((1, C, Pmax, 17, 3))
    kps[..., 0] = np.linspace(10, 500, 17)[None, None, None, :]
    kps[..., 1] = np.linspace(20, 400, 17)[None, None, None, :]
    kps[..., 2] = 0.9
    n_pose = np.full((1, C), Pmax, np.int32)
    n_trk = np.zeros(1, np.int32)
    prep = S.prepare(T(kps, dev), T(n_pose, dev, i32), T(n_trk, dev, i32), Tmax)
    n = C * Pmax
    iv, ip = prep["idx_view"].cpu().numpy()[0], prep["idx_pose"].cpu().numpy()[0]
    x = np.eye(n, dtype=bool)
    g1 = [q for q in range(n) if iv[q] in (0, 1, 2)]                 # 24 poses
    g2 = [q for q in range(n) if iv[q] == 3][:2]                      # a pair ... plus
    for grp in (g1, g2):
        for a in grp:
            for b in grp:
                x[a, b] = True
    words = S.pack_xbin(x, Tmax + n)[None].to(dev)
    out = {k: v.cpu().numpy() for k, v in S.assign(words.contiguous(), prep, T(n_trk, dev, i32), C, 16).items()}
    assert out["err"][0] == 0 and out["counts"][0, 2] == 0
    assert out["big_n"][0] == 1 and out["big_nsel"][0, 0] == 24 and out["big_slot"][0, 0] == 0
    assert out["new_n"][0] == 2 and out["new_nsel"][0, 0] == 24 and out["new_nsel"][0, 1] == 2
    want = [(int(iv[q]), int(ip[q])) for q in g1]
    assert [tuple(r) for r in out["big_sel"][0, 0, :24].tolist()] == want
    assert [tuple(r) for r in out["new_sel"][0, 0].tolist()] == want[:MAX_SEL]
    # the oracle's closure + parse + decode on the same X_bin
    groups = o.parse_groups(o.transform_closure(x), np.cumsum([0] + [Pmax] * C))
    dec = o.decode_groups(groups, 0, iv[:n], ip[:n], False)
    assert [s_ for s_ in dec.new_groups if len(s_) >= 2] == [want, [(int(iv[q]), int(ip[q])) for q in g2]]
