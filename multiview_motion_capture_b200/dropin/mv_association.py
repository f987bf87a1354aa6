"""`mv_association` under the reference's module name: same-signature seams of the matchers the capture path can use
(src/mv_association.py), running on the CUDA kernels of libmvmc.so. No CPU arithmetic: without the library and a GPU they
raise.

    match_als(W, dimGroup)      -> (match_mat, X_bin)      src/mv_association.py:222-318  (mvmc_match_als + mvmc_transform_closure)
    transform_closure(x_bin)    -> match_mat               src/mv_association.py:99-121   (mvmc_transform_closure)
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def _dev():
    from multiview_motion_capture_b200 import _lib
    return _lib.default_device()


def transform_closure(x_bin):
    """Binary relation matrix -> the reference's `match_result_mat` (its closure only runs through the last index;
    SURVEY.md 7). Same dtype and shape as `x_bin`."""
    import torch
    from multiview_motion_capture_b200 import stages
    x = np.asarray(x_bin)
    n = x.shape[0]
    if n == 0:
        return np.zeros_like(x)
    dev = _dev()
    words = stages.pack_xbin(x.astype(bool)).to(dev)[None]
    out = stages.transform_closure(words, torch.tensor([n], dtype=torch.int32, device=dev))
    return out[0, :n, :n].cpu().numpy().astype(x.dtype)


def match_als(W: np.ndarray, dimGroup, **kwargs):
    """ADMM / alternating-least-squares multi-way matcher. W: (n, n) similarity (float64, or the float32 matrix of the
    no-track path, whose first iteration the reference runs in float32); dimGroup: cumulative group offsets
    [0, n_0, n_0 + n_1, ...]. Returns (match_mat, X_bin) like the reference. The reference's keyword options
    (alpha=50, beta=0.1, tol=1e-4, maxIter=1000 ...) are the kernel's constants; passing other values raises."""
    import torch
    from multiview_motion_capture_b200 import stages
    from multiview_motion_capture_b200._lib import MAX_VIEWS
    defaults = dict(alpha=50, beta=0.1, tol=1e-4, maxIter=1000, max_iter=1000, pSelect=1, p_select=1, verbose=False, eigenvalues=False)
    for k, v in kwargs.items():
        if k not in defaults or defaults[k] != v:
            raise ValueError(f"match_als option {k}={v!r} is not supported by the CUDA kernel (reference defaults only)")
    W = np.asarray(W)
    n = W.shape[0]
    dg = np.asarray(dimGroup, dtype=np.int32).reshape(-1)
    if n == 0:
        return np.zeros((0, 0), dtype=bool), np.zeros((0, 0), dtype=bool)
    if len(dg) - 1 > MAX_VIEWS + 1:
        raise ValueError(f"at most {MAX_VIEWS + 1} groups")
    assert dg[0] == 0 and dg[-1] == n, "dimGroup must run from 0 to n"
    dev = _dev()
    N = -(-n // 32) * 32
    sim = np.zeros((1, N, N))
    sim[0, :n, :n] = W.astype(np.float64)
    sizes = np.diff(dg)
    rmax = -(-int(min(n, 2 * sizes.max())) // 16) * 16
    f32 = torch.tensor([int(W.dtype == np.float32)], dtype=torch.int32, device=dev)
    xbin, _ = stages.match_als(torch.from_numpy(sim).to(dev), torch.from_numpy(dg[None].copy()).to(dev), rmax, f32_first_iter=f32)
    x_bin = stages.unpack_xbin(xbin[0], n)
    mm = stages.transform_closure(xbin, torch.tensor([n], dtype=torch.int32, device=dev))[0, :n, :n].cpu().numpy().astype(bool)
    return mm, x_bin
