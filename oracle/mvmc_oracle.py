"""TEST INFRASTRUCTURE — CPU restatement (NumPy, float64) of the reference's capture hot path.

This is the parity oracle for the CUDA path. It is NOT product code: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it. The product (``multiview_motion_capture_b200``) never does.

Pinning: the reference ships no test or golden vector for this path. The oracle is
pinned instead against outputs of the reference itself run in the build container through
``oracle/ref_shim.py`` (``oracle/make_golden.py`` → ``tests/golden/shelf_ref.npz`` and
``synth_*_ref.npz``); ``tests/test_oracle_golden.py`` replays those.

Each function cites the reference lines it restates (paths relative to /root/reference/).
Third-party arithmetic that is not in /root/reference:
  * scipy.optimize.least_squares (requirements.txt pins scipy==1.3.2; container has 1.18.1):
    restated in :func:`trf_least_squares` from the published algorithm (Branch, Coleman & Li 1999;
    Moré 1977 for the trust-region sub-problem) as implemented by SciPy's ``trf_no_bounds``.
  * numpy/LAPACK ``svd``, ``det``, ``inv`` and ``RandomState(0).rand``: called directly.
  * cv2.computeCorrespondEpilines: restated in closed form (normalise F·x so that a²+b²=1).

Operation order follows the reference wherever rounding can leak into a discrete decision
(ALS stopping iteration, TRF accept/reject), so that on the golden inputs the oracle reproduces
the reference's numbers bit for bit on the same machine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import svd as _svd

EPS = np.finfo(np.float64).eps

# ----------------------------------------------------------------------------------------------
# joint tables  (src/pose_def.py:72-98 COCO order, :112-139 BODY_25 order, :210-230 BASIC_18)
# ----------------------------------------------------------------------------------------------
N_COCO = 17
N_B18 = 18
# COCO slot <- BODY_25 slot (src/pose_def.py:262-270)
BODY25_TO_COCO = np.array([0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11])
# BASIC_18 parents (src/pose_def.py:186-233)
B18_PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 8, 12, 13, 8, 15, 15])
# joints shared by BASIC_18 (3D track pose) and COCO (2D pose), in BASIC_18 order
# (src/pose_def.py:288-298 get_common_kps_idxs(BASIC_18, COCO))
COMMON_B18 = np.array([1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 17])
COMMON_COCO = np.array([11, 13, 15, 12, 14, 16, 5, 7, 9, 6, 8, 10, 0, 3, 4])
# IK residual joints: BASIC_18 ∩ (COCO + synthetic Spine at slot 17)
# (src/inverse_kinematics.py:366-378 with get_common_kps_idxs_1)
IK_SKEL_IDX = np.array([1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17])
IK_OBS_IDX = np.array([11, 13, 15, 12, 14, 16, 17, 5, 7, 9, 6, 8, 10, 0, 3, 4])
COCO_L_SHOULDER, COCO_R_SHOULDER, COCO_L_HIP, COCO_R_HIP = 5, 6, 11, 12

# reference skeleton (src/inverse_kinematics.py:120-173)
B18_OFFSETS = np.array([
    [0, 0, 0], [0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5], [-0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5],
    [0, 0, 0.3], [0, 0, 0.3], [0.2, 0, 0], [0.3, 0, 0], [0.3, 0, 0], [-0.2, 0, 0], [-0.3, 0, 0],
    [-0.3, 0, 0], [0, -0.02, 0.15], [0.07, 0.02, 0.1], [-0.07, 0.02, 0.1]], dtype=np.float64)
# side(+mid) bone-length vector -> per-joint length (11 -> 18)
B18_SIDE_TO_FULL = np.array([7, 0, 1, 2, 0, 1, 2, 8, 9, 3, 4, 5, 3, 4, 5, 10, 6, 6])
# which joint supplies each of the 11 side lengths: L_Hip,L_Knee,L_Ankle,L_Shoulder,L_Elbow,L_Wrist,L_Ear,
# Mid_Hip,Spine,Neck,Nose
B18_SIDE_SRC = np.array([1, 2, 3, 9, 10, 11, 16, 0, 7, 8, 15])


@dataclass
class Skeleton:
    """Same content as src/inverse_kinematics.py:94-117, arrays only."""
    bone_dirs: np.ndarray          # (J,3); row 0 is the raw root offset (not normalised)
    side_bone_lens: np.ndarray     # (11,)
    side_to_full: np.ndarray       # (J,)
    parents: np.ndarray            # (J,)

    @property
    def n_joints(self):
        return len(self.parents)


def load_skeleton() -> Skeleton:
    """src/inverse_kinematics.py:120-173 (offsets → unit directions + lengths; row 0 untouched)."""
    lens = np.linalg.norm(B18_OFFSETS, axis=-1)
    dirs = B18_OFFSETS.copy()
    dirs[1:, :] = dirs[1:, :] / lens[1:][:, np.newaxis]
    return Skeleton(bone_dirs=dirs, side_bone_lens=lens[B18_SIDE_SRC].copy(),
                    side_to_full=B18_SIDE_TO_FULL.copy(), parents=B18_PARENTS.copy())


# ----------------------------------------------------------------------------------------------
# A0  filter_bad_pose   (src/motion_capture.py:1023-1043)
# ----------------------------------------------------------------------------------------------
def pose_is_bad(kps: np.ndarray, min_score=0.01, n_min_valid=4, min_bb=5) -> bool:
    """kps (17,3) [x,y,score]."""
    valid = kps[:, 2] > min_score
    if np.sum(valid) < n_min_valid:
        return True
    pts = kps[valid, :2]
    size = np.max(pts, axis=0) - np.min(pts, axis=0)
    return bool(np.any(size < min_bb))


def body25_to_coco(kps25: np.ndarray) -> np.ndarray:
    """src/pose_def.py:262-270 — pure gather on the joint axis (second to last)."""
    return kps25[..., BODY25_TO_COCO, :]


# ----------------------------------------------------------------------------------------------
# A1  fundamental matrix from two projection matrices   (src/mv_math_util.py:57-77)
# ----------------------------------------------------------------------------------------------
def fundamental_from_projections(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    rows1 = [(1, 2), (2, 0), (0, 1)]
    f = np.zeros((3, 3), dtype=p1.dtype)
    for i in range(3):
        for j in range(3):
            stack = np.vstack([p1[rows1[j][0]], p1[rows1[j][1]], p2[rows1[i][0]], p2[rows1[i][1]]])
            f[i, j] = np.linalg.det(stack)
    return f


def _epilines(f: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """cv2.computeCorrespondEpilines(pts, 1, f) in closed form: rows of f applied to (x,y,1), scaled so
    a²+b²=1 (scale left at 1 if a=b=0). For whichImage=2 pass f.T."""
    x, y = pts[:, 0], pts[:, 1]
    a = f[0, 0] * x + f[0, 1] * y + f[0, 2]
    b = f[1, 0] * x + f[1, 1] * y + f[1, 2]
    c = f[2, 0] * x + f[2, 1] * y + f[2, 2]
    nu = a * a + b * b
    with np.errstate(divide="ignore", invalid="ignore"):
        nu = np.where(nu != 0, 1.0 / np.sqrt(nu), 1.0)
    return np.stack([a * nu, b * nu, c * nu], axis=1)


# ----------------------------------------------------------------------------------------------
# A2  symmetric epipolar distance between two 2D poses   (src/mv_math_util.py:80-115)
# ----------------------------------------------------------------------------------------------
def epipolar_error(f_mat: np.ndarray, kps1: np.ndarray, kps2: np.ndarray, min_score=0.1) -> float:
    """kps* (17,3). f_mat = fundamental_from_projections(P1, P2)."""
    l12 = _epilines(f_mat, kps1[:, :2])
    l21 = _epilines(f_mat.T, kps2[:, :2])
    valid = (kps1[:, 2] * kps2[:, 2]) > min_score
    if not np.any(valid):
        return np.nan
    total = 0
    cnt = 0
    for i in np.nonzero(valid)[0]:
        a, b, c = l12[i]
        d1 = abs(a * kps2[i, 0] + b * kps2[i, 1] + c) / np.sqrt(a ** 2 + b ** 2)
        a, b, c = l21[i]
        d2 = abs(a * kps1[i, 0] + b * kps1[i, 1] + c) / np.sqrt(a ** 2 + b ** 2)
        total = total + 0.5 * (d1 + d2)
        cnt += 1
    return total / cnt


# ----------------------------------------------------------------------------------------------
# A3  reprojection error between a 3D track pose and a 2D pose   (src/motion_capture.py:403-414)
# ----------------------------------------------------------------------------------------------
def reprojection_error(joints3d: np.ndarray, kps2d: np.ndarray, P: np.ndarray, min_score=0.1) -> float:
    """joints3d (18,3) BASIC_18 (scores are all 1, src/inverse_kinematics.py:430-431); kps2d (17,3)."""
    p3 = joints3d[COMMON_B18]
    p2 = kps2d[COMMON_COCO]
    homo = np.concatenate([p3[:, :3], np.ones((len(p3), 1))], axis=1)
    proj = P @ homo.T
    proj = (proj[:2] / (1e-5 + proj[2])).T
    mask = (p2[:, 2] * 1.0) > min_score
    if not mask.any():
        return np.nan
    e = np.linalg.norm(proj[mask, :2] - p2[mask, :2], axis=-1)
    return np.mean(e)


# ----------------------------------------------------------------------------------------------
# A4  distance + similarity matrix with alive tracks   (src/motion_capture.py:643-756)
# ----------------------------------------------------------------------------------------------
def build_dst_sim(track_joints: Sequence[np.ndarray], view_kps: Sequence[np.ndarray],
                  Ps: Sequence[np.ndarray]):
    """track_joints: T × (18,3); view_kps: per view (P_v,17,3) kept poses in id order; Ps: per view (3,4).
    Returns dst (n,n), sim (n,n), dim_groups (C+2,), idx_view (n,), idx_local (n,)."""
    T = len(track_joints)
    C = len(view_kps)
    sizes = [0, T] + [len(v) for v in view_kps]
    dim_groups = np.cumsum(sizes)
    n = int(dim_groups[-1])
    idx_view = np.full(n, -1, dtype=np.int64)
    idx_local = np.zeros(n, dtype=np.int64)
    idx_local[:T] = np.arange(T)
    for v in range(C):
        idx_view[dim_groups[v + 1]:dim_groups[v + 2]] = v
        idx_local[dim_groups[v + 1]:dim_groups[v + 2]] = np.arange(len(view_kps[v]))
    fcache: Dict[Tuple[int, int], np.ndarray] = {}
    dst = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        vi = idx_view[i]
        for j in range(n):
            if i == j:
                continue
            vj = idx_view[j]
            if vi >= 0 and vi == vj:
                dst[i, j] = np.nan
            elif vi >= 0 and vj >= 0:
                if (vi, vj) not in fcache:
                    fcache[(vi, vj)] = fundamental_from_projections(Ps[vi], Ps[vj])
                dst[i, j] = epipolar_error(fcache[(vi, vj)], view_kps[vi][idx_local[i]],
                                           view_kps[vj][idx_local[j]], 0.1)
            elif vi >= 0 and vj < 0:
                dst[i, j] = reprojection_error(track_joints[j], view_kps[vi][idx_local[i]], Ps[vi], 0.1)
            elif vi < 0 and vj >= 0:
                dst[i, j] = reprojection_error(track_joints[i], view_kps[vj][idx_local[j]], Ps[vj], 0.1)
            else:
                dst[i, j] = np.nan
    with np.errstate(all="ignore"):
        mx = np.nanmax(dst)
        dst[np.isnan(dst)] = mx + 1.0
        s = (dst - 15) / 30
        s = 1 / (1 + np.exp(5 * s))
    s[s < 1e-3] = 0
    s[s > 1.0] = 1.0
    return dst, s, dim_groups, idx_view, idx_local


# ----------------------------------------------------------------------------------------------
# A7  affinity when there is no alive track (float32 path)
#     src/motion_capture.py:597-631, src/mv_math_util.py:267-351
# ----------------------------------------------------------------------------------------------
def _skew(x):
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]], dtype=np.float64)


def pairwise_f_mats_krt(Ks: Sequence[np.ndarray], Rts: Sequence[np.ndarray]) -> np.ndarray:
    """src/mv_math_util.py:267-285 — F[i,j] from (K,R,t), computed in float64 and stored float32."""
    C = len(Ks)
    F = np.zeros((C, C, 3, 3), dtype=np.float32)
    for i in range(C):
        K0, R0, T0 = Ks[i], Rts[i][:, :3], Rts[i][:, 3]
        for j in range(C):
            K1, R1, T1 = Ks[j], Rts[j][:, :3], Rts[j][:, 3]
            f = np.linalg.inv(K0).T @ (R0 @ R1.T) @ K1.T @ _skew(K1 @ R1 @ R0.T @ (T0 - R0 @ R1.T @ T1))
            F[i, j] += f.astype(np.float64)
            if F[i, j].sum() == 0:
                F[i, j] += 1e-12
    return F


def _projected_distance(pts0: np.ndarray, pts1: np.ndarray, F: np.ndarray) -> np.ndarray:
    """src/mv_math_util.py:288-317: un-normalised |l·x| with l = normalise(Fᵀ x0), mean over 17 joints.
    pts0 (N0,17,2), pts1 (N1,17,2) float64, F (3,3) float32 → (N0,N1) float64."""
    lines = _epilines(F.astype(np.float64).T, pts0.reshape(-1, 2)).reshape(-1, 1, 17, 3)
    p1 = np.ones([1, pts1.shape[0], 17, 3])
    p1[0, :, :, :2] = pts1
    dist = np.abs(np.sum(lines * p1, axis=3))
    return np.mean(dist, axis=2)


def build_dst_sim_no_tracks(view_kps: Sequence[np.ndarray], Ks, Rts):
    C = len(view_kps)
    dim_groups = np.cumsum([0] + [len(v) for v in view_kps])
    M = int(dim_groups[-1])
    Fs = pairwise_f_mats_krt(Ks, Rts)
    D = np.ones((M, M), dtype=np.float32) * 50
    np.fill_diagonal(D, 0)
    for h in range(C):
        for k in range(h + 1, C):
            if len(view_kps[h]) == 0 or len(view_kps[k]) == 0:
                continue
            p0 = view_kps[h][:, :, :2]
            p1 = view_kps[k][:, :, :2]
            mean_dst = 0.5 * (_projected_distance(p0, p1, Fs[h, k]) + _projected_distance(p1, p0, Fs[k, h]).T)
            D[dim_groups[h]:dim_groups[h + 1], dim_groups[k]:dim_groups[k + 1]] = mean_dst
            D[dim_groups[k]:dim_groups[k + 1], dim_groups[h]:dim_groups[h + 1]] = \
                D[dim_groups[h]:dim_groups[h + 1], dim_groups[k]:dim_groups[k + 1]].T
    with np.errstate(all="ignore"):
        aff = -(D - D.mean()) / D.std()
        aff = 1 / (1 + np.exp(-5 * aff))
    return D, aff, dim_groups


# ----------------------------------------------------------------------------------------------
# NumPy's float32 kernels restated (third party, not under /root/reference): what `D.mean()`, `D.std()` and `np.exp`
# compute on the float32 distance matrix of the no-track path (src/mv_math_util.py:348-350). The CUDA side
# (csrc/affinity.cu, namespace np32) follows these; tests pin them bit for bit against the container's NumPy.
# ----------------------------------------------------------------------------------------------
_f32 = np.float32


def np32_pairwise_sum(a: np.ndarray):
    """numpy/_core/src/umath/loops_utils.h.src pairwise_sum for a contiguous float32 vector."""
    n = len(a)
    if n < 8:
        r = _f32(0.0)
        for v in a:
            r = _f32(r + v)
        return r
    if n <= 128:
        r = [_f32(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = _f32(r[j] + a[i + j])
            i += 8
        res = _f32(_f32(_f32(r[0] + r[1]) + _f32(r[2] + r[3])) + _f32(_f32(r[4] + r[5]) + _f32(r[6] + r[7])))
        while i < n:
            res = _f32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return _f32(np32_pairwise_sum(a[:n2]) + np32_pairwise_sum(a[n2:]))


def np32_mean_std(D: np.ndarray):
    """(D.mean(), D.std()) of a float32 matrix, as numpy/_core/_methods.py _mean/_var compute them."""
    flat = np.ascontiguousarray(D, dtype=_f32).reshape(-1)
    cnt = _f32(len(flat))
    mean = _f32(_f32(_f32(0) + np32_pairwise_sum(flat)) / cnt)
    x = (flat - mean).astype(_f32)
    x = (x * x).astype(_f32)
    return mean, np.sqrt(_f32(_f32(_f32(0) + np32_pairwise_sum(x)) / cnt))


def np32_exp(x: np.ndarray) -> np.ndarray:
    """NumPy's float32 exp (loops_exponent_log.dispatch.c.src, AVX2 / AVX512F kernel): Cody-Waite reduction and a (5,2)
    rational approximation evaluated with FMAs. Valid for |x| < 88."""
    def fma(a, b, c):   # a*b is exact in float64; the sum is rounded once more (double rounding is negligible for a test aid)
        return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(_f32)
    x = np.asarray(x, dtype=_f32)
    magic = _f32(12582912.0)
    k = (x * _f32(1.442695040888963407359924681001892137)).astype(_f32)
    k = ((k + magic).astype(_f32) - magic).astype(_f32)
    r = fma(k, _f32(-6.93145752e-1), x)
    r = fma(k, _f32(-1.42860677e-6), r)
    num = fma(_f32(5.082762527590693718096e-04), r, _f32(6.757896990527504603057e-03))
    for c in (5.114512081637298353406e-02, 2.473615434895520810817e-01, 7.257664613233124478488e-01, 9.999999999980870924916e-01):
        num = fma(num, r, _f32(c))
    den = fma(_f32(2.159509375685829852307e-02), r, _f32(-2.742335390411667452936e-01))
    den = fma(den, r, _f32(1.0))
    return np.ldexp((num / den).astype(_f32), k.astype(np.int32)).astype(_f32)


# ----------------------------------------------------------------------------------------------
# A5  ALS / ADMM low-rank multi-way matcher   (src/mv_association.py:222-318)
# ----------------------------------------------------------------------------------------------
def match_als(W: np.ndarray, dim_groups, alpha=50, beta=0.1, tol=1e-4, max_iter=1000):
    sizes = np.diff(dim_groups)
    n = W.shape[0]
    rank = min(n, int(max(sizes)) * 2)
    W = 0.5 * (W + W.T)
    X = W.copy()
    Z = W.copy()
    Y = np.zeros_like(W)
    mu = 64
    A = np.random.RandomState(0).rand(n, rank)
    eye = np.eye(rank)
    n_iter = 0
    ar = np.arange(n)
    for it in range(max_iter):
        n_iter = it + 1
        X0 = X
        X = Z - (Y - W + beta) / mu
        B = (np.linalg.inv(A.T @ A + alpha / mu * eye) @ (A.T @ X)).T
        A = (np.linalg.inv(B.T @ B + alpha / mu * eye) @ (B.T @ X.T)).T
        X = A @ B.T
        Z = X + Y / mu
        for g in range(len(dim_groups) - 1):
            Z[dim_groups[g]:dim_groups[g + 1], dim_groups[g]:dim_groups[g + 1]] = 0
        Z[ar, ar] = 1
        Z[Z < 0] = 0
        Z[Z > 1] = 1
        Y = Y + mu * (X - Z)
        p_res = np.linalg.norm(X - Z) / n
        d_res = mu * np.linalg.norm(X - X0) / n
        if p_res < tol and d_res < tol:
            break
        if p_res > 10 * d_res:
            mu = 2 * mu
        elif d_res > 10 * p_res:
            mu = mu / 2
    X = 0.5 * (X + X.T)
    return X > 0.5, n_iter


# ----------------------------------------------------------------------------------------------
# A6  closure quirk + leader assignment + parse   (src/mv_association.py:99-121,
#     src/motion_capture.py:417-446)
# ----------------------------------------------------------------------------------------------
def transform_closure(x_bin: np.ndarray) -> np.ndarray:
    """The reference's triple loop overwrites `temp` for every k, so only k = N-1 survives:
    temp = X | (X[:,N-1] ⊗ X[N-1,:]). Then rows become 'leaders' in index order."""
    n = x_bin.shape[0]
    if n == 0:
        return np.zeros_like(x_bin)
    temp = x_bin | np.outer(x_bin[:, n - 1], x_bin[n - 1, :])
    vis = np.zeros(n, dtype=bool)
    out = np.zeros_like(x_bin)
    for i in range(n):
        if vis[i]:
            continue
        members = temp[i]
        vis |= members
        out[members, i] = True
    return out


def parse_groups(match_mat: np.ndarray, dim_groups) -> List[List[Tuple[int, int, int]]]:
    """Groups = kept leader columns (≥2 members); every row joins the FIRST kept column it belongs to.
    Returns per group a list of (group_idx, local_idx, global_idx)."""
    n = match_mat.shape[0]
    cols = np.nonzero(match_mat.sum(axis=0) > 1.9)[0]
    sub = match_mat[:, cols]
    lists: List[List[int]] = [[] for _ in cols]
    for row in range(n):
        if sub[row].any():
            lists[int(np.argmax(sub[row]))].append(row)
    dg = np.asarray(dim_groups)
    out = []
    for members in lists:
        cur = []
        for idx in members:
            g = int(np.nonzero(dg <= idx)[0][-1])
            cur.append((g, idx - int(dg[g]), idx))
        if cur:
            out.append(cur)
    return out


@dataclass
class Association:
    track_matches: Dict[int, List[Tuple[int, int]]] = field(default_factory=dict)  # t_idx -> [(view, pose_id)]
    new_groups: List[List[Tuple[int, int]]] = field(default_factory=list)          # [(view, pose_id)]
    dst: Optional[np.ndarray] = None
    sim: Optional[np.ndarray] = None
    x_bin: Optional[np.ndarray] = None
    n_iter: int = 0
    dim_groups: Optional[np.ndarray] = None
    n_dup_view: int = 0


def decode_groups(groups, T: int, idx_view, idx_pose_id, has_tracks: bool) -> Association:
    """src/motion_capture.py:762-808 (with tracks) / :618-624 (without: no one-pose-per-view hack)."""
    out = Association()
    for g in groups:
        if not has_tracks:
            out.new_groups.append([(int(idx_view[gi]), int(idx_pose_id[gi])) for _, _, gi in g])
            continue
        t_idx = -1
        for _, _, gi in g:
            if gi < T:
                t_idx = gi
                break
        views: List[int] = []
        sel: List[Tuple[int, int]] = []
        for _, _, gi in g:
            if gi >= T:
                v = int(idx_view[gi])
                if v in views:
                    out.n_dup_view += 1
                    continue
                views.append(v)
                sel.append((v, int(idx_pose_id[gi])))
        if sel:
            if t_idx >= 0:
                out.track_matches[t_idx] = sel
            else:
                out.new_groups.append(sel)
    return out


def associate(track_joints, view_kps, view_pose_ids, Ps, Ks, Rts) -> Association:
    """src/motion_capture.py:829-835 associate_tracking."""
    T = len(track_joints)
    if T > 0:
        dst, sim, dg, idx_view, idx_local = build_dst_sim(track_joints, view_kps, Ps)
        pose_id = np.array([-1] * T + [pid for ids in view_pose_ids for pid in ids], dtype=np.int64)
    else:
        dst, sim, dg = build_dst_sim_no_tracks(view_kps, Ks, Rts)
        idx_view = np.concatenate([np.full(len(v), vi) for vi, v in enumerate(view_kps)]).astype(np.int64) \
            if len(view_kps) else np.zeros(0, dtype=np.int64)
        pose_id = np.array([pid for ids in view_pose_ids for pid in ids], dtype=np.int64)
    if sim.shape[0] == 0:
        return Association(dst=dst, sim=sim, dim_groups=dg)
    x_bin, n_iter = match_als(sim, dg)
    mm = transform_closure(x_bin)
    groups = parse_groups(mm, dg)
    a = decode_groups(groups, T, idx_view, pose_id, T > 0)
    a.dst, a.sim, a.x_bin, a.n_iter, a.dim_groups = dst, sim, x_bin, n_iter, np.asarray(dg)
    return a


# ----------------------------------------------------------------------------------------------
# I4  scipy.optimize.least_squares(method='trf', jac='2-point', unbounded, tr_solver='exact')
# ----------------------------------------------------------------------------------------------
def fd_jacobian(fun: Callable, x0: np.ndarray, f0: np.ndarray) -> np.ndarray:
    """Forward differences with SciPy's default step h_i = sqrt(eps)·sgn(x_i)·max(1,|x_i|), sgn(0)=+1,
    and dx_i = (x_i + h_i) − x_i."""
    n = x0.size
    sign = (x0 >= 0).astype(np.float64) * 2 - 1
    h = (EPS ** 0.5) * sign * np.maximum(1.0, np.abs(x0))
    jt = np.empty((n, f0.size), dtype=np.float64)
    for i in range(n):
        x1 = np.copy(x0)
        x1[i] = x0[i] + h[i]
        dx = (x0[i] + h[i]) - x0[i]
        jt[i] = (fun(x1) - f0) / dx
    return jt.T


def solve_tr_subproblem(n, m, uf, s, V, delta, alpha0, rtol=0.01, max_iter=10):
    """Moré's root-find on the LM parameter using one SVD (SciPy `solve_lsq_trust_region`)."""
    suf = s * uf
    if m >= n:
        full_rank = s[-1] > EPS * m * s[0]
    else:
        full_rank = False
    if full_rank:
        p = -V.dot(uf / s)
        if np.linalg.norm(p) <= delta:
            return p, 0.0, 0

    def phi_dphi(al):
        den = s ** 2 + al
        pn = np.linalg.norm(suf / den)
        return pn - delta, -np.sum(suf ** 2 / den ** 3) / pn

    a_hi = np.linalg.norm(suf) / delta
    if full_rank:
        ph, dph = phi_dphi(0.0)
        a_lo = -ph / dph
    else:
        a_lo = 0.0
    if alpha0 is None or (not full_rank and alpha0 == 0):
        alpha = max(0.001 * a_hi, (a_lo * a_hi) ** 0.5)
    else:
        alpha = alpha0
    it = -1
    for it in range(max_iter):
        if alpha < a_lo or alpha > a_hi:
            alpha = max(0.001 * a_hi, (a_lo * a_hi) ** 0.5)
        ph, dph = phi_dphi(alpha)
        if ph < 0:
            a_hi = alpha
        ratio = ph / dph
        a_lo = max(a_lo, alpha - ratio)
        alpha -= (ph + delta) * ratio / delta
        if np.abs(ph) < rtol * delta:
            break
    p = -V.dot(suf / (s ** 2 + alpha))
    p *= delta / np.linalg.norm(p)
    return p, alpha, it + 1


@dataclass
class LsqResult:
    x: np.ndarray
    cost: float
    fun: np.ndarray
    nfev: int
    njev: int
    status: int
    optimality: float
    trace: list = field(default_factory=list)


def trf_least_squares(fun: Callable, x0: np.ndarray, max_nfev: int, ftol=1e-8, xtol=1e-8, gtol=1e-8,
                      trace=False) -> LsqResult:
    x = np.array(x0, dtype=np.float64).copy()
    f = fun(x)
    nfev = 1
    J = fd_jacobian(fun, x, f)
    njev = 1
    m, n = J.shape
    cost = 0.5 * np.dot(f, f)
    g = J.T.dot(f)
    delta = np.linalg.norm(x)
    if delta == 0:
        delta = 1.0
    alpha = 0.0
    status = None
    tr = []
    while True:
        g_norm = np.linalg.norm(g, ord=np.inf)
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        U, s, Vt = _svd(J, full_matrices=False)
        V = Vt.T
        uf = U.T.dot(f)
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            p, alpha, n_it = solve_tr_subproblem(n, m, uf, s, V, delta, alpha)
            Jp = J.dot(p)
            predicted = -(0.5 * np.dot(Jp, Jp) + np.dot(p, g))
            x_new = x + p
            f_new = fun(x_new)
            nfev += 1
            p_norm = np.linalg.norm(p)
            if not np.all(np.isfinite(f_new)):
                delta = 0.25 * p_norm
                continue
            cost_new = 0.5 * np.dot(f_new, f_new)
            actual = cost - cost_new
            if predicted > 0:
                ratio = actual / predicted
            elif predicted == actual == 0:
                ratio = 1
            else:
                ratio = 0
            if ratio < 0.25:
                delta_new = 0.25 * p_norm
            elif ratio > 0.75 and p_norm > 0.95 * delta:
                delta_new = delta * 2.0
            else:
                delta_new = delta
            if trace:
                tr.append(dict(alpha=alpha, delta=delta, p_norm=p_norm, ratio=ratio, actual=actual,
                               predicted=predicted, n_it=n_it))
            f_ok = actual < ftol * cost and ratio > 0.25
            x_ok = p_norm < xtol * (xtol + np.linalg.norm(x))
            if f_ok and x_ok:
                status = 4
            elif f_ok:
                status = 2
            elif x_ok:
                status = 3
            if status is not None:
                break
            alpha *= delta / delta_new
            delta = delta_new
        if actual > 0:
            x = x_new
            f = f_new
            cost = cost_new
            J = fd_jacobian(fun, x, f)
            njev += 1
            g = J.T.dot(f)
    if status is None:
        status = 0
    return LsqResult(x=x, cost=cost, fun=f, nfev=nfev, njev=njev, status=status,
                     optimality=float(np.linalg.norm(g, ord=np.inf)), trace=tr)


# ----------------------------------------------------------------------------------------------
# B1/B2  DLT triangulation   (src/mv_math_util.py:152-240)
# ----------------------------------------------------------------------------------------------
def triangulate_point(Ps: Sequence[np.ndarray], pts: np.ndarray) -> np.ndarray:
    V = len(Ps)
    a = np.zeros((2 * V, 4))
    for j in range(V):
        a[2 * j] = pts[j][0] * Ps[j][2, :] - Ps[j][0, :]
        a[2 * j + 1] = pts[j][1] * Ps[j][2, :] - Ps[j][1, :]
    _, _, vh = np.linalg.svd(a, full_matrices=False)
    h = vh[3, :]
    return (h.T[:-1] / h.T[-1]).T


def triangulate_groups(Ps: Sequence[np.ndarray], groups: Sequence[np.ndarray], min_score: float,
                       post_optimize=False, max_nfev=2) -> np.ndarray:
    """groups: per view (K,3). Returns (K,4) [x,y,z,score]."""
    n_kps = len(groups[0])
    out = []
    for k in range(n_kps):
        sel = [v for v, g in enumerate(groups) if g[k, 2] >= min_score]
        if len(sel) < 2:
            sel = list(range(len(groups)))
        pts = np.array([groups[v][k, :] for v in sel])
        score = np.mean(pts[:, 2])
        p3 = triangulate_point([Ps[v] for v in sel], pts[:, :2])
        out.append((p3[0], p3[1], p3[2], score))
    out = np.array(out)
    if post_optimize:
        n_cams = len(Ps)

        def residual(x):
            loc = x.reshape((-1, 3))
            homo = np.concatenate([loc, np.ones((loc.shape[0], 1))], axis=-1).T
            res = []
            for v in range(n_cams):
                pr = Ps[v] @ homo
                pr = (pr[:2] / (pr[2] + 1e-6)).T
                d = np.linalg.norm(pr - groups[v][:, :2], axis=-1)
                res.append(d * groups[v][:, -1])
            return np.array(res).flatten()

        r = trf_least_squares(residual, out[:, :3].flatten().copy(), max_nfev)
        out[:, :3] = r.x.reshape((-1, 3))
    return out


# ----------------------------------------------------------------------------------------------
# I1  forward kinematics   (src/inverse_kinematics.py:176-199 + src/Quaternions.py:97-115,335-366,442-462)
# ----------------------------------------------------------------------------------------------
def _quat_axis_angle(angles: np.ndarray, axis: np.ndarray) -> np.ndarray:
    ax = axis / (np.sqrt(np.sum(axis ** 2, axis=-1)) + 1e-10)[..., np.newaxis]
    s = np.sin(angles / 2.0)[..., np.newaxis]
    c = np.cos(angles / 2.0)[..., np.newaxis]
    return np.concatenate([c, ax * s], axis=-1)


def _quat_mul(sq: np.ndarray, oq: np.ndarray) -> np.ndarray:
    """Quaternions.__mul__: note the (unusual) operand naming — result = sq ⊗ oq."""
    q0, q1, q2, q3 = sq[..., 0], sq[..., 1], sq[..., 2], sq[..., 3]
    r0, r1, r2, r3 = oq[..., 0], oq[..., 1], oq[..., 2], oq[..., 3]
    out = np.empty(sq.shape)
    out[..., 0] = r0 * q0 - r1 * q1 - r2 * q2 - r3 * q3
    out[..., 1] = r0 * q1 + r1 * q0 - r2 * q3 + r3 * q2
    out[..., 2] = r0 * q2 + r1 * q3 + r2 * q0 - r3 * q1
    out[..., 3] = r0 * q3 - r1 * q2 + r2 * q1 + r3 * q0
    return out


_AX = (np.array([1, 0, 0]), np.array([0, 1, 0]), np.array([0, 0, 1]))


def euler_to_rotmats(euler: np.ndarray) -> np.ndarray:
    """(J,3) → (J,3,3); local R = Rx(a)·Ry(b)·Rz(c) through half-angle quaternions."""
    qx = _quat_axis_angle(euler[..., 0], _AX[0])
    qy = _quat_axis_angle(euler[..., 1], _AX[1])
    qz = _quat_axis_angle(euler[..., 2], _AX[2])
    q = _quat_mul(qx, _quat_mul(qy, qz))
    qw, qx_, qy_, qz_ = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x2 = qx_ + qx_
    y2 = qy_ + qy_
    z2 = qz_ + qz_
    xx = qx_ * x2
    yy = qy_ * y2
    wx = qw * x2
    xy = qx_ * y2
    yz = qy_ * z2
    wy = qw * y2
    xz = qx_ * z2
    zz = qz_ * z2
    wz = qw * z2
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = 1.0 - (yy + zz)
    m[..., 0, 1] = xy - wz
    m[..., 0, 2] = xz + wy
    m[..., 1, 0] = xy + wz
    m[..., 1, 1] = 1.0 - (xx + zz)
    m[..., 1, 2] = yz - wx
    m[..., 2, 0] = xz - wy
    m[..., 2, 1] = yz + wx
    m[..., 2, 2] = 1.0 - (xx + yy)
    return m


def forward_kinematics(skel: Skeleton, root, euler: np.ndarray, side_lens: np.ndarray):
    """Returns (J,3) joint positions and (J,4,4) global transforms."""
    J = skel.n_joints
    rot = euler_to_rotmats(euler)
    full = np.array([side_lens[i] for i in skel.side_to_full])
    offs = skel.bone_dirs * full[:, np.newaxis]
    loc = np.zeros((J, 4, 4))
    loc[:, 3, 3] = 1.0
    loc[:, :3, :3] = rot
    loc[1:, :3, 3] = offs[1:]
    if root is not None:
        loc[0, :3, 3] = root
    glob = loc.copy()
    for j in range(1, J):
        glob[j, :, :] = glob[skel.parents[j], :, :] @ loc[j, :, :]
    pos = glob[:, :, 3]
    pos = pos[:, :3] / pos[:, 3, np.newaxis]
    return pos, glob


# ----------------------------------------------------------------------------------------------
# I0  synthetic mid-spine observation   (src/inverse_kinematics.py:339-348,370-378)
# ----------------------------------------------------------------------------------------------
def add_mid_spine(kps: np.ndarray) -> np.ndarray:
    """(17,3) COCO → (18,3): slot 17 = mean(mid-shoulder, mid-hip), score = product of the 4 scores."""
    ms = 0.5 * (kps[COCO_L_SHOULDER, :] + kps[COCO_R_SHOULDER, :])
    mh = 0.5 * (kps[COCO_L_HIP, :] + kps[COCO_R_HIP, :])
    sp = 0.5 * (ms + mh)
    score = kps[COCO_L_SHOULDER, -1] * kps[COCO_R_SHOULDER, -1]
    score *= kps[COCO_L_HIP, -1] * kps[COCO_R_HIP, -1]
    return np.concatenate([kps, np.array([sp[0], sp[1], score]).reshape((-1, 3))], axis=0)


# ----------------------------------------------------------------------------------------------
# I2/I3/I5  reprojection IK   (src/inverse_kinematics.py:202-277,380-433)
# ----------------------------------------------------------------------------------------------
@dataclass
class PoseParam:
    root: np.ndarray        # (3,)
    euler: np.ndarray       # (18,3)
    bone_lens: np.ndarray   # (11,)

    def pack(self) -> np.ndarray:
        return np.concatenate([self.root.flatten(), self.euler.flatten(), self.bone_lens.flatten()])

    @staticmethod
    def unpack(x: np.ndarray) -> "PoseParam":
        return PoseParam(x[:3].copy(), x[3:57].reshape(18, 3).copy(), x[57:68].copy())


def _reproj_residual(skel: Skeleton, obs: np.ndarray, Ps, root, euler, lens) -> np.ndarray:
    """obs (V,16,3) already gathered at IK_OBS_IDX."""
    pos, _ = forward_kinematics(skel, root, euler, lens)
    pos = pos[IK_SKEL_IDX, :]
    homo = np.concatenate([pos, np.ones((len(pos), 1), dtype=pos.dtype)], axis=-1).T
    proj = []
    for v in range(len(Ps)):
        kp = Ps[v] @ homo
        proj.append((kp[:2] / (1e-5 + kp[2])).T)
    d = np.array(proj) - obs[:, :, :2]
    d = d * obs[:, :, -1:]
    return d.flatten()


def solve_ik(skel: Skeleton, init: Optional[PoseParam], cam_kps: Sequence[np.ndarray], Ps: Sequence[np.ndarray],
             lsq=None, collect: Optional[list] = None):
    """PoseSolver.solve (src/inverse_kinematics.py:380-433). cam_kps: V × (17,3) COCO; returns
    (PoseParam, joints (18,3)). `lsq(fun, x0, max_nfev)` defaults to :func:`trf_least_squares`."""
    lsq = lsq or trf_least_squares
    obs18 = [add_mid_spine(k) for k in cam_kps]
    if init is None:
        p3 = triangulate_groups(Ps, obs18, 0.01, True) if lsq is trf_least_squares else \
            _triangulate_groups_with(lsq, Ps, obs18, 0.01)
        if collect is not None:
            collect.append(("tri", p3.copy()))
        root = 0.5 * (p3[COCO_L_HIP, :3] + p3[COCO_R_HIP, :3])
        init = PoseParam(root, np.zeros((skel.n_joints, 3), dtype=root.dtype), skel.side_bone_lens.copy())
        max_nfev = 50
    else:
        max_nfev = 5
    obs = np.array(obs18)[:, IK_OBS_IDX, :]
    nj = skel.n_joints

    def res1(x):
        return _reproj_residual(skel, obs, Ps, x[:3], x[3:].reshape((-1, 3)), init.bone_lens)

    x0 = np.concatenate([init.root.flatten(), init.euler.flatten()])
    r1 = lsq(res1, x0, max_nfev)
    if collect is not None:
        collect.append(("ik1", x0.copy(), r1))
    p1 = PoseParam(r1.x[:3], r1.x[3:].reshape((-1, 3)), init.bone_lens)

    def res2(x):
        return _reproj_residual(skel, obs, Ps, x[:3], x[3:3 + nj * 3].reshape((-1, 3)), x[3 + nj * 3:])

    x0b = np.concatenate([p1.root.flatten(), p1.euler.flatten(), p1.bone_lens.flatten()])
    r2 = lsq(res2, x0b, max_nfev)
    if collect is not None:
        collect.append(("ik2", x0b.copy(), r2))
    p2 = PoseParam(r2.x[:3], r2.x[3:3 + nj * 3].reshape((-1, 3)), r2.x[3 + nj * 3:])
    joints, _ = forward_kinematics(skel, p2.root, p2.euler, p2.bone_lens)
    return p2, joints


def _joints3d_residual(skel: Skeleton, target: np.ndarray, root, euler, lens) -> np.ndarray:
    """target (16,4) = obs_pose_3d[obs_kps_idxs] (x, y, z, score)."""
    pos, _ = forward_kinematics(skel, root, euler, lens)
    d = pos[IK_SKEL_IDX, :] - target[:, :3]
    d = d * target[:, -1:]
    return d.flatten()


def solve_pose_3d(skel: Skeleton, obs_pose_3d: np.ndarray, init: PoseParam, max_nfev=5, lsq=None):
    """solve_pose (src/inverse_kinematics.py:280-306): root + angles against 3D targets, bone lengths fixed."""
    lsq = lsq or trf_least_squares
    target = obs_pose_3d[IK_OBS_IDX, :]
    x0 = np.concatenate([init.root.flatten(), init.euler.flatten()])
    r = lsq(lambda x: _joints3d_residual(skel, target, x[:3], x[3:].reshape((-1, 3)), init.bone_lens), x0, max_nfev)
    return PoseParam(r.x[:3], r.x[3:].reshape((-1, 3)), init.bone_lens), r


def solve_pose_bone_lens_3d(skel: Skeleton, obs_pose_3d: np.ndarray, init: PoseParam, max_nfev=5, lsq=None):
    """solve_pose_bone_lens (src/inverse_kinematics.py:309-336): + the 11 side bone lengths."""
    lsq = lsq or trf_least_squares
    target = obs_pose_3d[IK_OBS_IDX, :]
    nj = skel.n_joints
    x0 = np.concatenate([init.root.flatten(), init.euler.flatten(), init.bone_lens.flatten()])
    r = lsq(lambda x: _joints3d_residual(skel, target, x[:3], x[3:3 + nj * 3].reshape((-1, 3)), x[3 + nj * 3:]), x0, max_nfev)
    return PoseParam(r.x[:3], r.x[3:3 + nj * 3].reshape((-1, 3)), r.x[3 + nj * 3:]), r


def solve_ik_3d(skel: Skeleton, init: Optional[PoseParam], cam_kps: Sequence[np.ndarray], Ps: Sequence[np.ndarray], lsq=None):
    """PoseSolver.solve with use_only_reproj = False (src/inverse_kinematics.py:385-415): triangulate (with the 2-nfev
    refine), then solve_pose and solve_pose_bone_lens against the triangulated points. Returns (PoseParam, joints, log)."""
    obs18 = [add_mid_spine(k) for k in cam_kps]
    if init is None:
        p3 = triangulate_groups(Ps, obs18, 0.01, True)
        root = 0.5 * (p3[COCO_L_HIP, :3] + p3[COCO_R_HIP, :3])
        init = PoseParam(root, np.zeros((skel.n_joints, 3), dtype=root.dtype), skel.side_bone_lens.copy())
        max_nfev = 50
    else:
        max_nfev = 5
    obs_pose_3d = triangulate_groups(Ps, obs18, 0.01, True)
    p1, r1 = solve_pose_3d(skel, obs_pose_3d, init, max_nfev, lsq)
    p2, r2 = solve_pose_bone_lens_3d(skel, obs_pose_3d, p1, max_nfev, lsq)
    joints, _ = forward_kinematics(skel, p2.root, p2.euler, p2.bone_lens)
    return p2, joints, dict(obs_pose_3d=obs_pose_3d, r1=r1, r2=r2, init=init)


def _triangulate_groups_with(lsq, Ps, groups, min_score):
    out = triangulate_groups(Ps, groups, min_score, False)
    n_cams = len(Ps)

    def residual(x):
        loc = x.reshape((-1, 3))
        homo = np.concatenate([loc, np.ones((loc.shape[0], 1))], axis=-1).T
        res = []
        for v in range(n_cams):
            pr = Ps[v] @ homo
            pr = (pr[:2] / (pr[2] + 1e-6)).T
            res.append(np.linalg.norm(pr - groups[v][:, :2], axis=-1) * groups[v][:, -1])
        return np.array(res).flatten()

    r = lsq(residual, out[:, :3].flatten().copy(), 2)
    out[:, :3] = r.x.reshape((-1, 3))
    return out


# ----------------------------------------------------------------------------------------------
# I6  generic chain FK with the quirks of src/kinematics.py:18-31
# ----------------------------------------------------------------------------------------------
def chain_fk_cmu(offsets: np.ndarray, parents: np.ndarray, rotmats: np.ndarray) -> np.ndarray:
    """Quirks kept: joint 0's parent index -1 wraps to the LAST joint's (still local) transform, joint 0
    gets no translation, and the output is xyz / z."""
    J = len(offsets)
    loc = np.zeros((J, 4, 4))
    loc[:, 3, 3] = 1.0
    loc[:, :3, :3] = rotmats
    loc[1:, :3, 3] = offsets[1:]
    glob = loc.copy()
    for j in range(J):
        glob[j] = glob[parents[j]] @ loc[j]
    pos = glob[:, :3, 3]
    with np.errstate(all="ignore"):
        return pos[:, :3] / pos[:, 2:3]


def chain_fk(offsets: np.ndarray, parents: np.ndarray, rotmats: np.ndarray, root=None) -> np.ndarray:
    """The same chain without the quirks (what the CUDA FK kernel computes for a generic skeleton)."""
    J = len(offsets)
    R = np.zeros((J, 3, 3))
    t = np.zeros((J, 3))
    for j in range(J):
        if parents[j] < 0:
            R[j] = rotmats[j]
            t[j] = offsets[j] if root is None else root
        else:
            p = parents[j]
            R[j] = R[p] @ rotmats[j]
            t[j] = R[p] @ offsets[j] + t[p]
    return t


# ----------------------------------------------------------------------------------------------
# SURVEY.md 8f-3  alternative matchers (dead code in the reference, kept for A/B):
#   match_objects_across_views  src/motion_capture.py:166-241 with PoseAssociation :44-101
#   tracklet_to_poses_association  src/motion_capture.py:844-871 with src/mv_math_util.py:11-32
# scipy.optimize.linear_sum_assignment is third party (called directly, as the reference does).
# ----------------------------------------------------------------------------------------------
# COCO / BASIC_18 indices of the 15 common joints in COCO order (get_common_kps_idxs(COCO, BASIC_18), src/pose_def.py:288-298)
RAY_COCO = np.array([0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
RAY_B18 = np.array([15, 16, 17, 9, 12, 10, 13, 11, 14, 1, 4, 2, 5, 3, 6])


def epipolar_matrix(view_kps: Sequence[np.ndarray], Ps: Sequence[np.ndarray]):
    """(n, n) calc_epipolar_error of every ordered cross-view pose pair (NaN: same view, or no commonly valid joint), and the
    views' offsets [0, P_0, P_0 + P_1, ...]."""
    off = np.cumsum([0] + [len(v) for v in view_kps])
    n = int(off[-1])
    D = np.full((n, n), np.nan)
    np.fill_diagonal(D, 0.0)
    for vi in range(len(view_kps)):
        for vj in range(len(view_kps)):
            if vi == vj:
                continue
            Fm = fundamental_from_projections(Ps[vi], Ps[vj])
            for a in range(len(view_kps[vi])):
                for b in range(len(view_kps[vj])):
                    D[off[vi] + a, off[vj] + b] = epipolar_error(Fm, view_kps[vi][a], view_kps[vj][b], 0.1)
    return D, off


def match_views_hungarian(dst: np.ndarray, view_offsets, threshold: float):
    """dst: (n, n) matrix whose 2D-2D entries are calc_epipolar_error(row pose, column pose) (build_dst_sim's `dst` before
    the NaN fill, or the pair errors themselves); view_offsets: cumulative offsets of the views' poses in dst
    [o_0, o_1, ..., o_C]. Returns the groups as lists of global indices, in the reference's creation / merge order."""
    from scipy.optimize import linear_sum_assignment
    counts = np.diff(view_offsets)
    init = int(np.argmax(counts))
    groups = [[g] for g in range(view_offsets[init], view_offsets[init + 1])]
    for v in range(len(counts)):
        if v == init or counts[v] < 1:
            continue
        poses = list(range(view_offsets[v], view_offsets[v + 1]))
        cost = np.zeros((len(groups), len(poses)))
        mask = np.zeros_like(cost).astype(np.int32)
        for pi, g in enumerate(poses):
            for hi, h in enumerate(groups):
                total, too_wrong = 0, False
                for q in h:
                    total += dst[q, g]
                    if total > threshold:
                        too_wrong = True
                cost[hi, pi] = total / len(h)
                mask[hi, pi] = int(too_wrong)
        rows, cols = linear_sum_assignment(cost)
        matched = set()
        for hi, pi in zip(rows, cols):
            matched.add(pi)
            if mask[hi, pi] == 1:
                groups.append([poses[pi]])
            else:
                groups[hi].append(poses[pi])
        for pi, g in enumerate(poses):
            if pi not in matched:
                groups.append([g])
    return groups


def tracklet_pose_costs(track_joints: Sequence[np.ndarray], kps_view: np.ndarray, Kr_inv: np.ndarray, cam_loc: np.ndarray):
    """cost[t, p] = tracklet_to_pose_2d_cost (src/motion_capture.py:845-850): mean over the 15 common joints of the
    distance between the track's 3D joint and the camera ray through the pose's 2D joint."""
    cost = np.zeros((len(track_joints), len(kps_view)))
    for t, j3 in enumerate(track_joints):
        for p, k2 in enumerate(kps_view):
            pts = np.concatenate([k2[RAY_COCO, :2], np.ones((15, 1))], axis=1)
            rays = (Kr_inv @ pts.T).T
            rays = rays / np.linalg.norm(rays, axis=-1, keepdims=True)
            d = [np.linalg.norm(np.cross(j3[RAY_B18[i], :3] - cam_loc, rays[i, :3])) for i in range(15)]
            cost[t, p] = np.mean(d)
    return cost


def tracklet_pose_association(track_joints, kps_view, pose_ids, Kr_inv, cam_loc, max_dst=0.1):
    from scipy.optimize import linear_sum_assignment
    if not len(track_joints) or not len(kps_view):
        return [], np.zeros((len(track_joints), len(kps_view)))
    cost = tracklet_pose_costs(track_joints, kps_view, Kr_inv, cam_loc)
    rows, cols = linear_sum_assignment(cost)
    return [(int(t), int(pose_ids[p])) for t, p in zip(rows, cols) if not cost[t, p] > max_dst], cost


# ----------------------------------------------------------------------------------------------
# L1  tracker lifecycle   (src/motion_capture.py:288-400, 873-963, run loop :1046-1129)
# ----------------------------------------------------------------------------------------------
TENTATIVE, CONFIRMED, DEAD = 1, 2, 3


@dataclass
class Track:
    track_id: int
    frame_idxs: List[int]
    params: List[PoseParam]
    joints: List[np.ndarray]
    views: List[List[Tuple[int, int]]]
    state: int = TENTATIVE
    hits: int = 1
    time_since_update: int = 0
    max_age: int = 0
    n_inits: int = 3

    def __len__(self):
        return len(self.frame_idxs)


class Tracker:
    """MvTracker on packed arrays. `frames` per call: kps (C,Pmax,17,3), n_pose (C,)."""

    def __init__(self, Ps, Ks, Rts, skel: Optional[Skeleton] = None, lsq=None):
        self.Ps, self.Ks, self.Rts = list(Ps), list(Ks), list(Rts)
        self.skel = skel or load_skeleton()
        self.tracks: List[Track] = []
        self.dead: List[Track] = []
        self.next_id = 0
        self.lsq = lsq
        self.last_assoc: Optional[Association] = None
        self.solve_log: list = []

    def step(self, frm_idx: int, kps: np.ndarray, n_pose: np.ndarray, forced_track_joints=None):
        C = len(self.Ps)
        view_ids, view_kps = [], []
        for v in range(C):
            ids = [p for p in range(int(n_pose[v])) if not pose_is_bad(kps[v, p])]
            view_ids.append(ids)
            view_kps.append(kps[v, ids] if ids else np.zeros((0, N_COCO, 3)))
        for t in self.tracks:
            t.time_since_update += 1
        alive = [t for t in self.tracks if t.state != DEAD]
        tj = [t.joints[-1] for t in alive] if forced_track_joints is None else list(forced_track_joints)
        a = associate(tj, view_kps, view_ids, self.Ps, self.Ks, self.Rts)
        self.last_assoc = a
        self.solve_log = []
        for ti, t in enumerate(alive):
            if ti in a.track_matches:
                sel = a.track_matches[ti]
                if len(sel) >= 2:
                    cam_kps = [kps[v, p] for v, p in sel]
                    prm, joints = solve_ik(self.skel, t.params[-1], cam_kps, [self.Ps[v] for v, _ in sel],
                                           self.lsq, self.solve_log)
                    t.frame_idxs.append(frm_idx)
                    t.params.append(prm)
                    t.joints.append(joints)
                    t.views.append(sel)
                    t.time_since_update = 0
                    t.hits += 1
                    if t.state == TENTATIVE and t.hits >= t.n_inits:
                        t.state = CONFIRMED
            else:
                if t.state == TENTATIVE:
                    t.state = DEAD
                elif t.time_since_update > t.max_age:
                    t.state = DEAD
        for sel in a.new_groups:
            if len(sel) >= 2:
                cam_kps = [kps[v, p] for v, p in sel]
                prm, joints = solve_ik(self.skel, None, cam_kps, [self.Ps[v] for v, _ in sel], self.lsq,
                                       self.solve_log)
                self.tracks.append(Track(self.next_id, [frm_idx], [prm], [joints], [sel]))
                self.next_id += 1
        self.dead.extend([t for t in self.tracks if t.state == DEAD])
        self.tracks = [t for t in self.tracks if t.state != DEAD]
        return a

    def finish(self) -> List[Track]:
        return sorted(self.tracks + self.dead, key=lambda t: -len(t))


def projections(Ks, Rts):
    return [K @ Rt for K, Rt in zip(Ks, Rts)]


def run_clip(kps_coco: np.ndarray, n_pose: np.ndarray, Ks, Rts, first_frame=1, last_frame=None, lsq=None):
    """run_main's loop (src/motion_capture.py:1062-1116): frame 0 is skipped."""
    F = kps_coco.shape[0]
    last = F - 1 if last_frame is None else last_frame
    trk = Tracker(projections(Ks, Rts), Ks, Rts, lsq=lsq)
    for f in range(first_frame, last + 1):
        trk.step(f, kps_coco[f], n_pose[f])
    return trk
