"""`common.Calib` / `common.FrameData` under the reference's module name (src/common.py:7-25): the classes inside the
per-frame input pickles (`{data_dir}/{frame:06d}.pkl` = List[FrameData])."""
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from pose_def import Pose


@dataclass
class Calib:
    K: np.ndarray        # 3x3
    Rt: np.ndarray       # 3x4
    P: np.ndarray        # 3x4 = K @ Rt
    Kr_inv: np.ndarray   # 3x3 = R^T K^-1
    img_wh_size: Tuple[int, int]

    @property
    def cam_loc(self):
        return -self.Rt[:3, :3].T @ self.Rt[:3, 3]


@dataclass
class FrameData:
    frame_idx: int
    poses: Dict[int, Pose]
    calib: Calib
    view_id: int
