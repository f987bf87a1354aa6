"""`inverse_kinematics` under the reference's module name: the classes pickled inside tracklets.pkl
(`inverse_kinematics.Skeleton`, `inverse_kinematics.PoseShapeParam`; src/inverse_kinematics.py:86-117) and
same-signature seams `load_skeleton`, `foward_kinematics`, `PoseSolver` (src/inverse_kinematics.py:120-199,351-433)
that run on the CUDA kernels (mvmc_fk, mvmc_ik_solve). No CPU arithmetic: without libmvmc.so and a GPU they raise."""
import os
import sys
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pose_def import BASIC_18_PARENTS, KpsFormat, Pose, get_pose_bones_index  # noqa: E402

# BASIC_18 rest offsets (metres), in joint order; side(+mid) length slot of every joint; joint that supplies each of the
# 11 side lengths (src/inverse_kinematics.py:120-173)
_OFFSETS = np.array([
    [0, 0, 0], [0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5], [-0.15, 0, 0], [0, 0, -0.5], [0, 0, -0.5],
    [0, 0, 0.3], [0, 0, 0.3], [0.2, 0, 0], [0.3, 0, 0], [0.3, 0, 0], [-0.2, 0, 0], [-0.3, 0, 0],
    [-0.3, 0, 0], [0, -0.02, 0.15], [0.07, 0.02, 0.1], [-0.07, 0.02, 0.1]], dtype=np.float64)
_SIDE_TO_FULL = [7, 0, 1, 2, 0, 1, 2, 8, 9, 3, 4, 5, 3, 4, 5, 10, 6, 6]
_SIDE_SRC = [1, 2, 3, 9, 10, 11, 16, 0, 7, 8, 15]


@dataclass
class PoseShapeParam:
    root: np.ndarray          # (3,)
    euler_angles: np.ndarray  # (18, 3)
    bone_lens: np.ndarray     # (11,)


@dataclass
class Skeleton:
    ref_joint_euler_angles: np.ndarray
    ref_bone_dirs: np.ndarray
    ref_side_bone_lens: np.ndarray
    ref_side_to_full_bone_lens_map: List[int]
    n_joints: int
    joint_parents: np.ndarray
    kps_format: KpsFormat

    @property
    def bone_idxs(self):
        return get_pose_bones_index(self.kps_format)

    def to_full_bone_lens(self, side_blens):
        assert len(side_blens) == len(self.ref_side_bone_lens)
        return np.array([side_blens[i] for i in self.ref_side_to_full_bone_lens_map])


def load_skeleton() -> Skeleton:
    lens = np.linalg.norm(_OFFSETS, axis=-1)
    dirs = _OFFSETS.copy()
    dirs[1:] = dirs[1:] / lens[1:, None]
    return Skeleton(ref_joint_euler_angles=np.zeros((18, 3)), ref_bone_dirs=dirs, ref_side_bone_lens=lens[_SIDE_SRC].copy(),
                    ref_side_to_full_bone_lens_map=list(_SIDE_TO_FULL), n_joints=18,
                    joint_parents=np.array(BASIC_18_PARENTS), kps_format=KpsFormat.BASIC_18)


def _device():
    from multiview_motion_capture_b200 import _lib
    return _lib.default_device()


def _pack(param: PoseShapeParam) -> np.ndarray:
    return np.concatenate([np.asarray(param.root).reshape(3), np.asarray(param.euler_angles).reshape(54),
                           np.asarray(param.bone_lens).reshape(11)])


def _unpack(x: np.ndarray) -> PoseShapeParam:
    return PoseShapeParam(root=x[:3].copy(), euler_angles=x[3:57].reshape(18, 3).copy(), bone_lens=x[57:68].copy())


def foward_kinematics(skel: Skeleton, param: PoseShapeParam):
    """(g_pos (18, 3), g_transforms) like the reference (src/inverse_kinematics.py:176-199): every reference caller unpacks
    two values. The positions come from the CUDA kernel (mvmc_fk); the second value is None - nothing on the capture path
    reads the 4x4 chain (the callers write `locs, _ = foward_kinematics(...)`)."""
    import torch
    from multiview_motion_capture_b200 import stages
    x = torch.as_tensor(_pack(param)[None], dtype=torch.float64, device=_device())
    return stages.fk(x).cpu().numpy()[0], None


_IK_OBS_IDX = [11, 13, 15, 12, 14, 16, 17, 5, 7, 9, 6, 8, 10, 0, 3, 4]     # 18-point observation index of the 16 IK joints
_IK_SKEL_IDX = [1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17]


def _solve_targets(obs_pose_3d, obs_kps_idxs, skel_kps_idxs, init_param, n_max_iter, stages):
    import torch
    from multiview_motion_capture_b200 import stages as S
    if list(obs_kps_idxs) != _IK_OBS_IDX or list(skel_kps_idxs) != _IK_SKEL_IDX:
        raise ValueError("the kernel solves the reference's 16 common joints (get_common_kps_idxs_1 of BASIC_18 and COCO + Spine)")
    dev = _device()
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    target = np.asarray(obs_pose_3d, dtype=np.float64)[_IK_OBS_IDX][None]
    x, joints, info, cost = S.ik_solve_targets(t(target), t(_pack(init_param)[None]), t([int(n_max_iter)], torch.int32), stages)
    return _unpack(x.cpu().numpy()[0]), info.cpu().numpy()[0], cost.cpu().numpy()[0]


def solve_pose(skel, obs_pose_3d, obs_kps_idxs, skel_kps_idxs, init_param: PoseShapeParam, n_max_iter=5) -> PoseShapeParam:
    """src/inverse_kinematics.py:280-306 on the device (mvmc_ik_solve_targets, stage 1): root + Euler angles fitted to the
    triangulated 3D joints `obs_pose_3d` (18, 4) [x, y, z, score]; bone lengths stay."""
    return _solve_targets(obs_pose_3d, obs_kps_idxs, skel_kps_idxs, init_param, n_max_iter, 1)[0]


def solve_pose_bone_lens(skel, obs_pose_3d, obs_kps_idxs, skel_kps_idxs, init_param: PoseShapeParam, n_max_iter=5) -> PoseShapeParam:
    """src/inverse_kinematics.py:309-336 on the device (stage 2): root + angles + the 11 side bone lengths."""
    return _solve_targets(obs_pose_3d, obs_kps_idxs, skel_kps_idxs, init_param, n_max_iter, 2)[0]


class PoseSolver:
    """PoseSolver(skel, init_pose, cam_poses_2d, cam_projs[, cam_calibs], obs_kps_format).solve() -> (PoseShapeParam, Pose).
    `use_only_reproj` (the reference's hard-coded toggle, src/inverse_kinematics.py:402): True = the live reprojection
    solves; False = triangulate, then solve_pose + solve_pose_bone_lens against the 3D points (:409-415).
    cam_poses_2d: V x (17, 3) COCO arrays [x, y, score]; cam_projs: V x (3, 4). init_pose=None is a track birth
    (triangulation + 50-evaluation solves), otherwise a 5-evaluation update from the previous parameters."""

    def __init__(self, skeleton: Skeleton, init_pose: Optional[PoseShapeParam], cam_poses_2d, cam_projs, cam_calibs=None,
                 obs_kps_format: KpsFormat = KpsFormat.COCO, use_only_reproj=True):
        if obs_kps_format != KpsFormat.COCO:
            raise ValueError("the capture path observes COCO-17 poses")
        self.skel, self.init_pose, self.use_only_reproj = skeleton, init_pose, use_only_reproj
        self.obs_kps_idxs, self.skel_kps_idxs = list(_IK_OBS_IDX), list(_IK_SKEL_IDX)
        self.cam_poses_2d = [np.asarray(p, dtype=np.float64) for p in cam_poses_2d]
        self.cam_projs = [np.asarray(p, dtype=np.float64) for p in cam_projs]

    def solve(self):
        import torch
        from multiview_motion_capture_b200 import stages
        from multiview_motion_capture_b200._lib import MAX_SEL
        V = len(self.cam_poses_2d)
        if not 2 <= V <= MAX_SEL:
            raise ValueError(f"a solve needs 2..{MAX_SEL} views, got {V}")
        dev = _device()
        kps = np.zeros((1, MAX_SEL, 17, 3))
        P = np.zeros((1, MAX_SEL, 3, 4))
        kps[0, :V] = np.stack(self.cam_poses_2d)
        P[0, :V] = np.stack(self.cam_projs)
        birth = self.init_pose is None
        x0 = np.zeros((1, 68)) if birth else _pack(self.init_pose)[None]
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        if not self.use_only_reproj:
            # 18-point observations (COCO + mid spine, built on the device by the birth path of mvmc_triangulate's caller: here
            # the spine row is formed by the same expression on the device through torch) -> triangulate + 2-nfev refine
            k = t(kps)
            ls, rs, lh, rh = k[:, :, 5], k[:, :, 6], k[:, :, 11], k[:, :, 12]
            spine = torch.cat([0.5 * (0.5 * (ls[..., :2] + rs[..., :2]) + 0.5 * (lh[..., :2] + rh[..., :2])),
                               ((ls[..., 2] * rs[..., 2]) * (lh[..., 2] * rh[..., 2]))[..., None]], -1)
            obs18 = torch.cat([k, spine[:, :, None]], 2).contiguous()
            p3 = stages.triangulate(obs18, t(P), t([V], torch.int32), 0.01, 2)           # [1,18,4]
            if birth:
                x0t = torch.zeros((1, 68), dtype=torch.float64, device=dev)
                x0t[0, :3] = 0.5 * (p3[0, 11, :3] + p3[0, 12, :3])
                x0t[0, 57:] = t(self.skel.ref_side_bone_lens)
            else:
                x0t = t(x0)
            target = p3[:, _IK_OBS_IDX].contiguous()
            x, joints, info, cost = stages.ik_solve_targets(target, x0t, t([50 if birth else 5], torch.int32), 3)
            self.info, self.cost, self.obs_pose_3d = info.cpu().numpy()[0], cost.cpu().numpy()[0], p3.cpu().numpy()[0]
            pose = Pose(KpsFormat.BASIC_18, joints.cpu().numpy()[0], np.ones((18, 1)), None)
            return _unpack(x.cpu().numpy()[0]), pose
        x, joints, info, cost = stages.ik_solve(t(kps), t(P), t([V], torch.int32), t(x0), t([int(birth)], torch.uint8),
                                                t([50 if birth else 5], torch.int32))
        self.info, self.cost = info.cpu().numpy()[0], cost.cpu().numpy()[0]
        pose = Pose(KpsFormat.BASIC_18, joints.cpu().numpy()[0], np.ones((18, 1)), None)
        return _unpack(x.cpu().numpy()[0]), pose
