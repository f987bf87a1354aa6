"""`mv_math_util.triangulate_point_groups_from_multiple_views_linear` under the reference's name and signature
(src/mv_math_util.py:152-212), running on the CUDA kernel (mvmc_triangulate)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def triangulate_point_groups_from_multiple_views_linear(proj_matricies, points_grps, min_score, post_optimize=False,
                                                        n_max_iter=2):
    """proj_matricies: V x (3, 4); points_grps: V x (K, 3) [x, y, score], K <= 18 -> (K, 4) [x, y, z, mean score]."""
    import torch
    from inverse_kinematics import _device
    from multiview_motion_capture_b200 import stages
    from multiview_motion_capture_b200._lib import MAX_SEL
    V, K = len(points_grps), len(points_grps[0])
    obs = np.zeros((1, MAX_SEL, K, 3))
    P = np.zeros((1, MAX_SEL, 3, 4))
    obs[0, :V] = np.stack([np.asarray(g, dtype=np.float64) for g in points_grps])
    P[0, :V] = np.asarray(proj_matricies, dtype=np.float64)
    dev = _device()
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    out = stages.triangulate(t(obs), t(P), t([V], torch.int32), float(min_score), int(n_max_iter) if post_optimize else 0)
    return out.cpu().numpy()[0]
