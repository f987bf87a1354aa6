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


class PoseSolver:
    """PoseSolver(skel, init_pose, cam_poses_2d, cam_projs[, cam_calibs], obs_kps_format).solve() -> (PoseShapeParam, Pose).
    cam_poses_2d: V x (17, 3) COCO arrays [x, y, score]; cam_projs: V x (3, 4). init_pose=None is a track birth
    (triangulation + 50-evaluation solves), otherwise a 5-evaluation update from the previous parameters."""

    def __init__(self, skeleton: Skeleton, init_pose: Optional[PoseShapeParam], cam_poses_2d, cam_projs, cam_calibs=None,
                 obs_kps_format: KpsFormat = KpsFormat.COCO):
        if obs_kps_format != KpsFormat.COCO:
            raise ValueError("the capture path observes COCO-17 poses")
        self.skel, self.init_pose = skeleton, init_pose
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
        x, joints, info, cost = stages.ik_solve(t(kps), t(P), t([V], torch.int32), t(x0), t([int(birth)], torch.uint8),
                                                t([50 if birth else 5], torch.int32))
        self.info, self.cost = info.cpu().numpy()[0], cost.cpu().numpy()[0]
        pose = Pose(KpsFormat.BASIC_18, joints.cpu().numpy()[0], np.ones((18, 1)), None)
        return _unpack(x.cpu().numpy()[0]), pose
