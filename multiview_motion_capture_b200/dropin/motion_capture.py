"""Drop-in for the reference's run-mode entry point (src/motion_capture.py:1221-1255):

    python motion_capture.py --mode prepare --opn_kps_dir D --calib_dir D --out_data_dir D
    python motion_capture.py --mode run --data_dir D --output_dir D [--video_dir D] [--max_frames 300]

Same flags, same per-frame input pickles (List[FrameData]) and the same `tracklets.pkl` layout
({"tracklets": List[MvTracklet]} sorted by -len; src/motion_capture.py:1120-1129), but every stage of the per-frame hot
path (pose filter, affinities, ALS matcher, assignment, triangulation, IK, track lifecycle) runs in the CUDA library
(libmvmc.so) through `multiview_motion_capture_b200.clips.ClipBatch`. The Python objects below only hold results.
`--video_dir` is accepted and ignored: the reference reads the videos only to draw debug images (:1070-1075).

Several clips can be tracked at once (the GPU path is batched over independent clips): pass `--data_dir` a directory
whose sub-directories each hold one clip's pickles; outputs go to `{output_dir}/{clip}/tracklets.pkl`.
"""
import argparse
import json
import os
import pickle
import sys
from enum import Enum
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
for _p in (_HERE, _ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from common import Calib, FrameData  # noqa: E402
from inverse_kinematics import PoseShapeParam, PoseSolver, Skeleton, load_skeleton  # noqa: E402,F401
from pose_def import KpsFormat, Pose, conversion_openpose_25_to_coco  # noqa: E402


class TrackState(Enum):
    Tentative = 1
    Confirmed = 2
    Dead = 3


class SpatialMatch:
    def __init__(self, view_idxs, pose_ids):
        self.view_idxs: List[int] = list(view_idxs)
        self.pose_ids: List[int] = list(pose_ids)

    def __len__(self):
        return len(self.view_idxs)


class SpatialTimeMatch:
    """What associate_tracking returns in the reference (src/motion_capture.py:449-478)."""

    def __init__(self):
        self.spatial_time_matches: Dict[int, SpatialMatch] = {}
        self.spatial_matches: List[SpatialMatch] = []
        self.dst_mat = self.sim_mat = self.match_mat = None


class MvTracklet:
    """Result holder with the reference's attribute layout (src/motion_capture.py:312-340). It never solves anything:
    births and updates are computed on the device and appended here by MvTracker."""

    def __init__(self, frm_idx, cam_poses_2d, cam_projs, cam_calibs, skel, pparam, pose, n_inits=3, max_age=0):
        self.frame_idxs: List[int] = [frm_idx]
        self.cam_poses_2d: List[List[Tuple[int, Pose]]] = [cam_poses_2d]
        self.cam_projs: List[List[np.ndarray]] = [cam_projs]
        self.cam_calibs: List[List[Calib]] = [cam_calibs]
        self.skel = skel
        self.poses: List[Tuple[int, PoseShapeParam, Pose]] = [(frm_idx, pparam, pose)]
        self.time_since_update = 0
        self.hits = 1
        self.state = TrackState.Tentative
        self.max_age = max_age
        self.n_inits = n_inits

    @property
    def last_pose_3d(self):
        return self.poses[-1][-1]

    def __len__(self):
        return len(self.frame_idxs)

    def is_tentative(self):
        return self.state == TrackState.Tentative

    def is_confirmed(self):
        return self.state == TrackState.Confirmed

    def is_dead(self):
        return self.state == TrackState.Dead


def _param_of(rec_track) -> PoseShapeParam:
    x = rec_track["param"]
    return PoseShapeParam(root=x[:3].copy(), euler_angles=x[3:57].reshape(18, 3).copy(), bone_lens=x[57:68].copy())


class MvTracker:
    """MvTracker (src/motion_capture.py:838-963) for a batch of independent clips. `update_4d` keeps the reference's
    single-clip signature; `update_4d_batch` advances every clip of the batch by one frame in one device step."""

    def __init__(self, skeleton: Skeleton, n_clips=1, n_views=None, max_poses=None, max_tracks=None, device=None):
        self.skel = skeleton
        self.n_clips = n_clips
        self._cfg = dict(n_views=n_views, max_poses=max_poses, max_tracks=max_tracks, device=device)
        self._cb = None
        self._calib_set = False
        self._tlets: List[Dict[int, MvTracklet]] = [dict() for _ in range(n_clips)]     # track_id -> tracklet
        self._order: List[List[int]] = [[] for _ in range(n_clips)]
        self._dead: List[List[MvTracklet]] = [[] for _ in range(n_clips)]
        self.n_dup_view = 0

    # reference attribute names (clip 0)
    @property
    def tracklets(self) -> List[MvTracklet]:
        return self.clip_tracklets(0)

    @property
    def dead_tracklets(self) -> List[MvTracklet]:
        return self._dead[0]

    def clip_tracklets(self, b):
        return [self._tlets[b][i] for i in self._order[b]]

    def clip_dead_tracklets(self, b):
        return self._dead[b]

    def _ensure(self, frames_per_clip):
        if self._cb is not None:
            return
        from multiview_motion_capture_b200.clips import ClipBatch
        C = self._cfg["n_views"] or len(frames_per_clip[0])
        pm = max([max(f.poses.keys(), default=-1) + 1 for fr in frames_per_clip for f in fr] + [1])
        Pmax = self._cfg["max_poses"] or min(32, max(8, 2 * pm))
        self._cb = ClipBatch(self.n_clips, C, Pmax, max_tracks=self._cfg["max_tracks"], device=self._cfg["device"])

    def update_4d(self, frm_idx: int, frames: List[FrameData], debug_view_imgs=None):
        assert self.n_clips == 1
        self.update_4d_batch(frm_idx, [frames])

    def update_4d_batch(self, frm_idx: int, frames_per_clip: List[List[FrameData]]):
        assert len(frames_per_clip) == self.n_clips
        self._ensure(frames_per_clip)
        cb = self._cb
        if not self._calib_set:
            K = np.stack([np.stack([f.calib.K for f in fr]) for fr in frames_per_clip])
            Rt = np.stack([np.stack([f.calib.Rt for f in fr]) for fr in frames_per_clip])
            P = np.stack([np.stack([f.calib.P for f in fr]) for fr in frames_per_clip])
            cb.set_calib(K, Rt, P)
            self._calib_set = True
        kps, n_pose = cb.kps_host, cb.n_pose_host
        kps[:] = 0.0
        n_pose[:] = 0
        for b, fr in enumerate(frames_per_clip):
            assert len(fr) == cb.C, "every frame must carry one FrameData per camera"
            for v, f in enumerate(fr):
                for pid, pose in f.poses.items():
                    if pid >= cb.Pmax:
                        raise ValueError(f"pose id {pid} exceeds max_poses={cb.Pmax}")
                    kps[b, v, pid, :, :2] = pose.keypoints
                    kps[b, v, pid, :, 2] = np.asarray(pose.keypoints_score).reshape(-1)
                    n_pose[b, v] = max(n_pose[b, v], pid + 1)
        recs = cb.step_pinned(frm_idx)
        for b, fr in enumerate(frames_per_clip):
            self._fold(b, frm_idx, fr, recs[b])

    def _fold(self, b, frm_idx, frames, rec):
        tl = self._tlets[b]
        n = int(rec["n_alive"])
        self.n_dup_view += int(rec["n_dup_view"])
        alive_ids = []
        for t in rec["tracks"][:n]:
            tid = int(t["track_id"])
            alive_ids.append(tid)
            sel = [(int(v), int(p)) for v, p in t["sel"][:t["n_sel"]]]
            if t["updated"] == 2:
                cam_poses = [(v, frames[v].poses[p]) for v, p in sel]
                pose = Pose(KpsFormat.BASIC_18, t["joints"].reshape(18, 3).copy(), np.ones((18, 1)), None)
                tl[tid] = MvTracklet(frm_idx, cam_poses, [frames[v].calib.P for v, _ in sel], [frames[v].calib for v, _ in sel],
                                     self.skel, _param_of(t), pose, n_inits=3, max_age=0)
            elif t["updated"] == 1:
                trk = tl[tid]
                trk.frame_idxs.append(frm_idx)
                trk.cam_poses_2d.append([(v, frames[v].poses[p]) for v, p in sel])
                trk.cam_projs.append([frames[v].calib.P for v, _ in sel])
                pose = Pose(KpsFormat.BASIC_18, t["joints"].reshape(18, 3).copy(), np.ones((18, 1)), None)
                trk.poses.append((frm_idx, _param_of(t), pose))
            trk = tl[tid]
            trk.time_since_update = int(t["time_since_update"])
            trk.hits = int(t["hits"])
            trk.state = TrackState(int(t["state"]))
        for tid in rec["died_ids"][:int(rec["n_died"])]:
            trk = tl.pop(int(tid))
            trk.state = TrackState.Dead
            trk.time_since_update += 1
            self._dead[b].append(trk)
        self._order[b] = alive_ids

    def finish(self, b=0) -> List[MvTracklet]:
        return sorted(self.clip_tracklets(b) + self._dead[b], key=lambda t: -len(t))


# ---- prepare mode (src/motion_capture.py:250-272, 974-1005) ----------------------------------------------------------
def load_calib(path) -> Calib:
    with open(path) as f:
        js = json.load(f)
    K = np.array(js["K"], dtype=np.float64).reshape(3, 3)
    Rt = np.array(js["RT"], dtype=np.float64).reshape(3, 4)
    return Calib(K=K, Rt=Rt, P=K @ Rt, Kr_inv=Rt[:3, :3].T @ np.linalg.inv(K), img_wh_size=js["imgSize"])


def parse_openpose_kps(path) -> Dict[int, Pose]:
    with open(path) as f:
        people = json.load(f)["people"]
    poses = {}
    for idx, person in enumerate(people):
        kps = conversion_openpose_25_to_coco(np.array(person["pose_keypoints_2d"], dtype=np.float64).reshape(-1, 3))
        poses[idx] = Pose(KpsFormat.COCO, keypoints=kps[:, :2], keypoints_score=kps[:, -1][:, np.newaxis], box=None)
    return poses


def extract_frame_data_from_openpose(in_dir: Path, calib_dir: Path, out_data_dir: Path):
    cam_dirs = sorted([d for d in Path(in_dir).iterdir() if d.is_dir()], key=lambda d: d.stem)
    calibs = [load_calib(Path(calib_dir) / f"{d.stem}.json") for d in cam_dirs]
    per_cam = [sorted(d.glob("*.json"), key=lambda p: p.stem) for d in cam_dirs]
    n_frames = min(len(p) for p in per_cam)
    os.makedirs(out_data_dir, exist_ok=True)
    for frm in range(n_frames):
        frames = [FrameData(frm, parse_openpose_kps(per_cam[v][frm]), calibs[v], view_id=v + 1) for v in range(len(cam_dirs))]
        with open(Path(out_data_dir) / f"{frm:06d}.pkl", "wb") as f:
            pickle.dump(frames, f)
    return n_frames


# ---- run mode (src/motion_capture.py:1046-1129) ----------------------------------------------------------------------
def _clip_dirs(pose_dir: Path) -> List[Path]:
    if any(Path(pose_dir).glob("*.pkl")):
        return [Path(pose_dir)]
    return sorted(d for d in Path(pose_dir).iterdir() if d.is_dir() and any(d.glob("*.pkl")))


def run_main(video_dir: Path, pose_dir: Path, out_dir: Path, max_frames: int = 300, device=None):
    clips = _clip_dirs(pose_dir)
    if not clips:
        raise SystemExit(f"no frame pickles under {pose_dir}")
    paths = [sorted(c.glob("*.pkl"), key=lambda p: int(p.stem)) for c in clips]
    tracker = MvTracker(load_skeleton(), n_clips=len(clips), device=device)
    n_frames = min(len(p) for p in paths)
    n_test = min(n_frames, max_frames)
    frm_idx = 0
    while True:                      # frame 0 is skipped, as in the reference (:1063 precedes :1077)
        frm_idx += 1
        if frm_idx >= n_frames:
            break
        frames = []
        for p in paths:
            with open(p[frm_idx], "rb") as f:
                frames.append(pickle.load(f))
        tracker.update_4d_batch(frm_idx, frames)
        if frm_idx >= n_test:
            break
    for b, c in enumerate(clips):
        dst = Path(out_dir) if len(clips) == 1 else Path(out_dir) / c.name
        os.makedirs(dst, exist_ok=True)
        with open(dst / "tracklets.pkl", "wb") as f:
            pickle.dump(file=f, obj={"tracklets": tracker.finish(b)})
    return tracker


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--mode", type=str, choices=["prepare", "run", "viz"],
                        help="run motion capture or prepare pre-generated data")
    parser.add_argument("--tlet_path", type=str, default="./tracklets.pkl")
    parser.add_argument("--video_dir", type=str, default="", help="accepted for compatibility; unused")
    parser.add_argument("--data_dir", type=str, default="", help="pre-generated data directory")
    parser.add_argument("--output_dir", type=str, default="", help="output directory")
    parser.add_argument("--opn_kps_dir", type=str, default="")
    parser.add_argument("--calib_dir", type=str, default="", help="calibration directory")
    parser.add_argument("--out_data_dir", type=str, default="", help="output data directory")
    parser.add_argument("--max_frames", type=int, default=300, help="the reference stops after 300 frames (:1059)")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    if args.mode == "run":
        run_main(Path(args.video_dir), Path(args.data_dir), Path(args.output_dir), max_frames=args.max_frames)
    elif args.mode == "prepare":
        extract_frame_data_from_openpose(Path(args.opn_kps_dir), Path(args.calib_dir), Path(args.out_data_dir))
    elif args.mode == "viz":
        raise SystemExit("viz mode (matplotlib/Qt animation of tracklets.pkl) is outside the capture hot path; "
                         "the pickle written by --mode run loads in the reference's viz_tracklets")
    else:
        raise SystemExit("--mode is required")


if __name__ == "__main__":
    main()
