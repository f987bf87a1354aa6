"""Stage-level wrappers over the C-ABI (one per reference seam). Tensors are torch tensors
on the CUDA device (or CPU tensors when the test tier has bound the kernel emulator)."""
import torch

from . import _lib
from ._lib import MAX_SEL, N_B18, N_COCO, N_PARAM, check, ptr

f64 = torch.float64
i32 = torch.int32


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else None


def _c(t, dtype):
    assert t.dtype == dtype, (t.dtype, dtype)
    return t.contiguous()


def rand_stream(n, device):
    """First n doubles of numpy.random.RandomState(0).rand (mv_association.py:271)."""
    import numpy as np
    out = np.empty(n, dtype=np.float64)
    check(_lib.get_lib().mvmc_rand_stream_host(out.ctypes.data, n), "mvmc_rand_stream_host")
    return torch.from_numpy(out).to(device)


def fundamental(P):
    """mv_math_util.py:57-77 for all ordered camera pairs. P [B,C,3,4] -> F [B,C,C,3,3]."""
    P = _c(P, f64)
    B, C = P.shape[:2]
    F = torch.empty((B, C, C, 3, 3), dtype=f64, device=P.device)
    check(_lib.get_lib().mvmc_fundamental(ptr(P), ptr(F), B, C, _stream(P)), "mvmc_fundamental")
    return F


def fundamental_krt(K, Rt):
    """mv_math_util.py:267-285. K [B,C,3,3], Rt [B,C,3,4] -> F32 [B,C,C,3,3] float32."""
    K, Rt = _c(K, f64), _c(Rt, f64)
    B, C = K.shape[:2]
    F = torch.empty((B, C, C, 3, 3), dtype=torch.float32, device=K.device)
    check(_lib.get_lib().mvmc_fundamental_krt(ptr(K), ptr(Rt), ptr(F), B, C, _stream(K)), "mvmc_fundamental_krt")
    return F


def prepare(kps, n_pose, n_trk, Tmax):
    """filter_bad_pose + index layout. kps [B,C,Pmax,17,3]; returns dict(keep, dim_groups, idx_view, idx_pose)."""
    kps, n_pose, n_trk = _c(kps, f64), _c(n_pose, i32), _c(n_trk, i32)
    B, C, Pmax = kps.shape[:3]
    N = Tmax + C * Pmax
    dev = kps.device
    keep = torch.zeros((B, C, Pmax), dtype=torch.uint8, device=dev)
    dg = torch.zeros((B, C + 2), dtype=i32, device=dev)
    iv = torch.zeros((B, N), dtype=i32, device=dev)
    ip = torch.zeros((B, N), dtype=i32, device=dev)
    check(_lib.get_lib().mvmc_prepare(ptr(kps), ptr(n_pose), ptr(n_trk), B, C, Pmax, Tmax, ptr(keep), ptr(dg), ptr(iv),
                                      ptr(ip), _stream(kps)), "mvmc_prepare")
    return dict(keep=keep, dim_groups=dg, idx_view=iv, idx_pose=ip, N=N, Tmax=Tmax)


def affinity(kps, P, F, F32, trk_joints, n_trk, prep):
    """dst / sim matrices [B,N,N] (leading n x n block valid)."""
    kps, P, F, trk_joints, n_trk = _c(kps, f64), _c(P, f64), _c(F, f64), _c(trk_joints, f64), _c(n_trk, i32)
    F32 = _c(F32, torch.float32)
    B, C, Pmax = kps.shape[:3]
    Tmax, N = prep["Tmax"], prep["N"]
    assert trk_joints.shape == (B, Tmax, N_B18, 3)
    dst = torch.zeros((B, N, N), dtype=f64, device=kps.device)
    sim = torch.zeros((B, N, N), dtype=f64, device=kps.device)
    check(_lib.get_lib().mvmc_affinity(ptr(kps), ptr(P), ptr(F), ptr(F32), ptr(trk_joints), ptr(n_trk),
                                       ptr(prep["dim_groups"]), ptr(prep["idx_view"]), ptr(prep["idx_pose"]), B, C, Pmax,
                                       Tmax, ptr(dst), ptr(sim), _stream(kps)), "mvmc_affinity")
    return dst, sim


def match_als(sim, dim_groups, rmax, f32_first_iter=None, rand=None):
    """mv_association.py:222-318. sim [B,N,N], dim_groups [B,G+1] -> xbin [B,N,NW] uint32 (as int32 bits), n_iter [B]."""
    sim, dim_groups = _c(sim, f64), _c(dim_groups, i32)
    B, N, _ = sim.shape
    G = dim_groups.shape[1] - 1
    dev = sim.device
    lib = _lib.get_lib()
    if rand is None:
        rand = rand_stream(N * rmax, dev)
    ws = torch.empty(lib.mvmc_match_als_workspace_bytes(B, N, rmax) // 8, dtype=f64, device=dev)
    NW = (N + 31) // 32
    xbin = torch.zeros((B, N, NW), dtype=i32, device=dev)
    n_iter = torch.zeros((B,), dtype=i32, device=dev)
    f32p = ptr(_c(f32_first_iter, i32)) if f32_first_iter is not None else None
    check(lib.mvmc_match_als(ptr(sim), ptr(dim_groups), G, f32p, ptr(rand), B, N, rmax, ptr(ws), ptr(xbin), ptr(n_iter),
                             _stream(sim)), "mvmc_match_als")
    return xbin, n_iter


def unpack_xbin(xbin, n):
    """[N,NW] int32 bit rows -> (n,n) bool numpy."""
    import numpy as np
    w = xbin.cpu().numpy().view(np.uint32)
    bits = ((w[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(w.shape[0], -1)
    return bits[:n, :n].astype(bool)


def transform_closure(xbin, n):
    """mv_association.py:99-121 on the device. xbin [B,N,NW] int32 bit rows, n [B] int32 -> match_mat [B,N,N] uint8."""
    xbin, n = _c(xbin, i32), _c(n, i32)
    B, N, _ = xbin.shape
    out = torch.zeros((B, N, N), dtype=torch.uint8, device=xbin.device)
    check(_lib.get_lib().mvmc_transform_closure(ptr(xbin), ptr(n), B, N, ptr(out), _stream(xbin)), "mvmc_transform_closure")
    return out


def pack_xbin(x_bin, N=None):
    """(n,n) bool numpy -> [N,NW] int32 bit rows (inverse of unpack_xbin)."""
    import numpy as np
    n = x_bin.shape[0]
    N = N or n
    NW = (N + 31) // 32
    bits = np.zeros((N, NW * 32), dtype=np.uint32)
    bits[:n, :n] = x_bin
    words = (bits.reshape(N, NW, 32) << np.arange(32, dtype=np.uint32)).sum(-1).astype(np.uint32)
    return torch.from_numpy(words.view(np.int32).copy())


def assign(xbin, prep, n_trk, C, max_new):
    """closure + parse + decode. Returns dict of int tensors (see include/mvmc.h mvmc_assign)."""
    xbin, n_trk = _c(xbin, i32), _c(n_trk, i32)
    B, N, _ = xbin.shape
    Tmax = prep["Tmax"]
    dev = xbin.device
    z = lambda *s: torch.zeros(s, dtype=i32, device=dev)
    from ._lib import MAX_BIG, MAX_GROUP
    out = dict(trk_nsel=z(B, max(Tmax, 1)), trk_sel=z(B, max(Tmax, 1), MAX_SEL, 2), new_n=z(B), new_nsel=z(B, max_new),
               new_sel=z(B, max_new, MAX_SEL, 2), counts=z(B, 4), err=z(B), new_seq=z(B, max_new), singles=z(B, max_new, 3),
               big_n=z(B), big_nsel=z(B, MAX_BIG), big_sel=z(B, MAX_BIG, MAX_GROUP, 2), big_slot=z(B, MAX_BIG))
    check(_lib.get_lib().mvmc_assign_groups(ptr(xbin), ptr(prep["dim_groups"]), ptr(prep["idx_view"]), ptr(prep["idx_pose"]),
                                            ptr(n_trk), B, C, N, Tmax, max_new, ptr(out["trk_nsel"]), ptr(out["trk_sel"]),
                                            ptr(out["new_n"]), ptr(out["new_nsel"]), ptr(out["new_sel"]), ptr(out["counts"]),
                                            ptr(out["err"]), ptr(out["new_seq"]), ptr(out["singles"]), ptr(out["big_n"]),
                                            ptr(out["big_nsel"]), ptr(out["big_sel"]), ptr(out["big_slot"]), _stream(xbin)),
          "mvmc_assign_groups")
    return out


def triangulate(obs, Psel, n_views, min_score, refine_nfev=0):
    """mv_math_util.py:152-240. obs [M,V,K,3], Psel [M,V,3,4], n_views [M] -> [M,K,4]."""
    obs, Psel, n_views = _c(obs, f64), _c(Psel, f64), _c(n_views, i32)
    if obs.is_cuda and _lib.torch_ops() is not None:
        return _lib.torch_ops().triangulate(obs, Psel, n_views, float(min_score), int(refine_nfev))
    M, V, K = obs.shape[:3]
    out = torch.zeros((M, K, 4), dtype=f64, device=obs.device)
    check(_lib.get_lib().mvmc_triangulate(ptr(obs), ptr(Psel), ptr(n_views), M, V, K, float(min_score), int(refine_nfev),
                                          ptr(out), _stream(obs)), "mvmc_triangulate")
    return out


def fk(params):
    """inverse_kinematics.py:176-199. params [M,68] -> joints [M,18,3]."""
    params = _c(params, f64)
    if params.is_cuda and _lib.torch_ops() is not None:
        return _lib.torch_ops().fk(params)
    M = params.shape[0]
    out = torch.empty((M, N_B18, 3), dtype=f64, device=params.device)
    check(_lib.get_lib().mvmc_fk(ptr(params), M, ptr(out), _stream(params)), "mvmc_fk")
    return out


def fk_chain(rot, offsets, parents, root=None):
    """Generic chain FK. rot [M,J,3,3], offsets [J,3], parents [J] int32, root [M,3] or None -> [M,J,3]."""
    rot, offsets, parents = _c(rot, f64), _c(offsets, f64), _c(parents, i32)
    M, J = rot.shape[:2]
    out = torch.empty((M, J, 3), dtype=f64, device=rot.device)
    rp = ptr(_c(root, f64)) if root is not None else None
    check(_lib.get_lib().mvmc_fk_chain(ptr(rot), ptr(offsets), ptr(parents), rp, M, J, ptr(out), _stream(rot)),
          "mvmc_fk_chain")
    return out


def ik_solve(kps2d, Psel, n_views, x0, birth=None, max_nfev=None, free_mask=None):
    """PoseSolver.solve batched. kps2d [M,V,17,3] (V == 8), Psel [M,V,3,4], n_views [M], x0 [M,68].
    Returns x_out [M,68], joints [M,18,3], info [M,2,4] (nfev, njev, status, n_free), cost [M,2]."""
    kps2d, Psel, n_views, x0 = _c(kps2d, f64), _c(Psel, f64), _c(n_views, i32), _c(x0, f64)
    M, V = kps2d.shape[:2]
    dev = kps2d.device
    lib = _lib.get_lib()
    if max_nfev is None:
        max_nfev = torch.full((M,), 5, dtype=i32, device=dev)
    max_nfev = _c(max_nfev, i32)
    if kps2d.is_cuda and _lib.torch_ops() is not None:
        return _lib.torch_ops().ik_solve(kps2d, Psel, n_views, x0, _c(birth, torch.uint8) if birth is not None else None, max_nfev,
                                         _c(free_mask, torch.uint8) if free_mask is not None else None)
    bp = ptr(_c(birth, torch.uint8)) if birth is not None else None
    fp = ptr(_c(free_mask, torch.uint8)) if free_mask is not None else None
    ws = torch.empty(lib.mvmc_ik_workspace_bytes(M, V) // 8, dtype=f64, device=dev)
    x_out = torch.zeros((M, N_PARAM), dtype=f64, device=dev)
    joints = torch.zeros((M, N_B18, 3), dtype=f64, device=dev)
    info = torch.zeros((M, 2, 4), dtype=i32, device=dev)
    cost = torch.zeros((M, 2), dtype=f64, device=dev)
    check(lib.mvmc_ik_solve(ptr(kps2d), ptr(Psel), ptr(n_views), ptr(x0), bp, ptr(max_nfev), fp, M, V, ptr(ws), ptr(x_out),
                            ptr(joints), ptr(info), ptr(cost), _stream(kps2d)), "mvmc_ik_solve")
    return x_out, joints, info, cost


def ik_solve_targets(target, x0, max_nfev, stages=3):
    """inverse_kinematics.py:280-336 (solve_pose, solve_pose_bone_lens) batched: target [M,16,4] (x, y, z, score at the 16 IK
    joints), x0 [M,68], max_nfev [M]; stages bit 0 = root + angles, bit 1 = + bone lengths. Returns like ik_solve."""
    target, x0, max_nfev = _c(target, f64), _c(x0, f64), _c(max_nfev, i32)
    M = target.shape[0]
    dev = target.device
    x_out = torch.zeros((M, N_PARAM), dtype=f64, device=dev)
    joints = torch.zeros((M, N_B18, 3), dtype=f64, device=dev)
    info = torch.zeros((M, 2, 4), dtype=i32, device=dev)
    cost = torch.zeros((M, 2), dtype=f64, device=dev)
    check(_lib.get_lib().mvmc_ik_solve_targets(ptr(target), ptr(x0), ptr(max_nfev), int(stages), M, ptr(x_out), ptr(joints), ptr(info),
                                               ptr(cost), _stream(target)), "mvmc_ik_solve_targets")
    return x_out, joints, info, cost


def distances(kps, P, F, trk_joints, n_trk, prep):
    """Float64 distance matrix only (mvmc_distances): [B,N,N]."""
    kps, P, F, trk_joints, n_trk = _c(kps, f64), _c(P, f64), _c(F, f64), _c(trk_joints, f64), _c(n_trk, i32)
    B, C, Pmax = kps.shape[:3]
    Tmax, N = prep["Tmax"], prep["N"]
    dst = torch.zeros((B, N, N), dtype=f64, device=kps.device)
    check(_lib.get_lib().mvmc_distances(ptr(kps), ptr(P), ptr(F), ptr(trk_joints), ptr(n_trk), ptr(prep["dim_groups"]),
                                        ptr(prep["idx_view"]), ptr(prep["idx_pose"]), B, C, Pmax, Tmax, ptr(dst), _stream(kps)),
          "mvmc_distances")
    return dst


def linear_sum_assignment(cost, n_rows=None, n_cols=None):
    """scipy.optimize.linear_sum_assignment batched on the device: cost [B,R,C] -> col_of_row [B,R] (-1 unassigned), status [B]."""
    cost = _c(cost, f64)
    B, R, Cc = cost.shape
    dev = cost.device
    n_rows = torch.full((B,), R, dtype=i32, device=dev) if n_rows is None else _c(n_rows, i32)
    n_cols = torch.full((B,), Cc, dtype=i32, device=dev) if n_cols is None else _c(n_cols, i32)
    col = torch.full((B, R), -1, dtype=i32, device=dev)
    status = torch.zeros((B,), dtype=i32, device=dev)
    check(_lib.get_lib().mvmc_linear_sum_assignment(ptr(cost), ptr(n_rows), ptr(n_cols), B, R, Cc, ptr(col), ptr(status), _stream(cost)),
          "mvmc_linear_sum_assignment")
    return col, status


def match_views_hungarian(dst, dim_groups, threshold):
    """motion_capture.py:166-241 on the device: dst [B,N,N] (distances()), dim_groups [B,C+2] -> group_of [B,N], n_groups [B], status [B]."""
    dst, dim_groups = _c(dst, f64), _c(dim_groups, i32)
    B, N, _ = dst.shape
    C = dim_groups.shape[1] - 2
    dev = dst.device
    lib = _lib.get_lib()
    ws = torch.empty(lib.mvmc_match_views_workspace_bytes(B) // 8, dtype=f64, device=dev)
    gof = torch.full((B, N), -1, dtype=i32, device=dev)
    ng = torch.zeros((B,), dtype=i32, device=dev)
    status = torch.zeros((B,), dtype=i32, device=dev)
    check(lib.mvmc_match_views_hungarian(ptr(dst), ptr(dim_groups), B, C, N, float(threshold), ptr(ws), ptr(gof), ptr(ng), ptr(status),
                                         _stream(dst)), "mvmc_match_views_hungarian")
    return gof, ng, status


def tracklet_pose_association(trk_joints, n_trk, kps, keep, Kr_inv, cam_loc, max_dst=0.1):
    """motion_capture.py:844-871 on the device -> match [B,C,Tmax] (pose id or -1), cost [B,C,Tmax,Pmax], status [B*C]."""
    trk_joints, n_trk, kps, Kr_inv, cam_loc = _c(trk_joints, f64), _c(n_trk, i32), _c(kps, f64), _c(Kr_inv, f64), _c(cam_loc, f64)
    keep = _c(keep, torch.uint8)
    B, C, Pmax = kps.shape[:3]
    Tmax = trk_joints.shape[1]
    dev = kps.device
    match = torch.full((B, C, Tmax), -1, dtype=i32, device=dev)
    cost = torch.zeros((B, C, Tmax, Pmax), dtype=f64, device=dev)
    status = torch.zeros((B * C,), dtype=i32, device=dev)
    check(_lib.get_lib().mvmc_tracklet_pose_association(ptr(trk_joints), ptr(n_trk), ptr(kps), ptr(keep), ptr(Kr_inv), ptr(cam_loc), B, C,
                                                        Pmax, Tmax, float(max_dst), ptr(match), ptr(cost), ptr(status), _stream(kps)),
          "mvmc_tracklet_pose_association")
    return match, cost, status


def ik_birth_big(kps, P, groups, max_nfev=50):
    """Births from groups of any size (mvmc_ik_birth_big; the slow path behind MVMC_MAX_SEL). kps [B,C,Pmax,17,3], P [B,C,3,4],
    groups: per clip a list of groups, each a list of (view, pose id). Returns x [B,G,68], joints [B,G,18,3], info [B,G,2,4],
    cost [B,G,2] with G = the largest number of groups of a clip."""
    import numpy as np
    from ._lib import MAX_GROUP
    kps, P = _c(kps, f64), _c(P, f64)
    B, C, Pmax = kps.shape[:3]
    G = max(1, max(len(g) for g in groups))
    big_n = np.array([len(g) for g in groups], dtype=np.int32)
    big_nsel = np.zeros((B, G), dtype=np.int32)
    big_sel = np.zeros((B, G, MAX_GROUP, 2), dtype=np.int32)
    big_slot = np.tile(np.arange(G, dtype=np.int32), (B, 1))
    for b, gs in enumerate(groups):
        for g, sel in enumerate(gs):
            assert 2 <= len(sel) <= MAX_GROUP
            big_nsel[b, g] = len(sel)
            big_sel[b, g, :len(sel)] = np.asarray(sel, dtype=np.int32)
    dev = kps.device
    t = lambda a: torch.from_numpy(a).to(dev)
    lib = _lib.get_lib()
    ws = torch.zeros(lib.mvmc_ik_birth_big_workspace_bytes() // 8 + 1, dtype=f64, device=dev)
    x = torch.zeros((B, G, N_PARAM), dtype=f64, device=dev)
    joints = torch.zeros((B, G, N_B18, 3), dtype=f64, device=dev)
    info = torch.zeros((B, G, 2, 4), dtype=i32, device=dev)
    cost = torch.zeros((B, G, 2), dtype=f64, device=dev)
    bn, bs, bl, bo = t(big_n), t(big_nsel), t(big_sel), t(big_slot)
    check(lib.mvmc_ik_birth_big(ptr(kps), ptr(P), ptr(bn), ptr(bs), ptr(bl), ptr(bo), B, C, Pmax, G, G, 0, int(max_nfev), ptr(ws), ptr(x),
                                ptr(joints), ptr(info), ptr(cost), _stream(kps)), "mvmc_ik_birth_big")
    return x, joints, info, cost
