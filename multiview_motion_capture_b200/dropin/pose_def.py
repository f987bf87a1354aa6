"""Data model of the capture path under the reference's module name, so that its pickles (`pose_def.Pose`,
`pose_def.KpsFormat`) load and are written unchanged. Reference: src/pose_def.py:49-69 (KpsFormat, Pose), :262-270
(BODY_25 -> COCO gather), :186-233 (BASIC_18 joint order / parents). Only what the run-mode path and the
tracklets.pkl consumer need; no arithmetic lives here."""
from dataclasses import dataclass
from enum import Enum
from typing import List, Optional, Tuple

import numpy as np


class KpsFormat(Enum):
    COCO = 0
    OPENPOSE_25 = 1
    SMPLX_22 = 2
    BASIC_18 = 3


@dataclass
class Pose:
    pose_type: KpsFormat
    keypoints: np.ndarray
    keypoints_score: Optional[np.ndarray]
    box: Optional[np.ndarray]

    def to_kps_array(self):
        return np.concatenate([self.keypoints, self.keypoints_score.reshape((-1, 1))], axis=1)


COCO_JOINTS = ["Nose", "L_Eye", "R_Eye", "L_Ear", "R_Ear", "L_Shoulder", "R_Shoulder", "L_Elbow", "R_Elbow", "L_Wrist",
               "R_Wrist", "L_Hip", "R_Hip", "L_Knee", "R_Knee", "L_Ankle", "R_Ankle"]
BASIC_18_JOINTS = ["Mid_Hip", "L_Hip", "L_Knee", "L_Ankle", "R_Hip", "R_Knee", "R_Ankle", "Spine", "Neck", "L_Shoulder",
                   "L_Elbow", "L_Wrist", "R_Shoulder", "R_Elbow", "R_Wrist", "Nose", "L_Ear", "R_Ear"]
BASIC_18_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 8, 12, 13, 8, 15, 15]
# COCO slot <- OpenPose BODY_25 slot
BODY25_TO_COCO = [0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11]
_COCO_BONES = [("Nose", "L_Eye"), ("L_Eye", "L_Ear"), ("Nose", "R_Eye"), ("R_Eye", "R_Ear"), ("L_Shoulder", "R_Shoulder"),
               ("L_Shoulder", "L_Elbow"), ("L_Elbow", "L_Wrist"), ("R_Shoulder", "R_Elbow"), ("R_Elbow", "R_Wrist"),
               ("L_Shoulder", "L_Hip"), ("L_Hip", "L_Knee"), ("L_Knee", "L_Ankle"), ("R_Shoulder", "R_Hip"),
               ("R_Hip", "R_Knee"), ("R_Knee", "R_Ankle")]


def conversion_openpose_25_to_coco(kps25: np.ndarray) -> np.ndarray:
    """(25, 3) OpenPose BODY_25 -> (17, 3) COCO: a pure gather on the joint axis."""
    return np.asarray(kps25)[BODY25_TO_COCO, :]


def get_joint_names(fmt: KpsFormat) -> List[str]:
    if fmt == KpsFormat.COCO:
        return COCO_JOINTS
    if fmt == KpsFormat.BASIC_18:
        return BASIC_18_JOINTS
    raise ValueError(f"joint table of {fmt} is not part of the capture path")


def get_pose_bones_index(fmt: KpsFormat) -> List[Tuple[int, int]]:
    """Bone list as joint-index pairs (what the tracklet visualiser draws)."""
    if fmt == KpsFormat.COCO:
        return [(COCO_JOINTS.index(a), COCO_JOINTS.index(b)) for a, b in _COCO_BONES]
    if fmt == KpsFormat.BASIC_18:
        return [(j, p) for j, p in enumerate(BASIC_18_PARENTS) if p >= 0]
    raise ValueError(f"bone table of {fmt} is not part of the capture path")
