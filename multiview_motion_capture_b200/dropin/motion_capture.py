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
from dataclasses import dataclass, field
from enum import Enum
from pathlib import Path
from typing import Dict, List, Optional, Tuple

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


@dataclass
class SpatialMatch:
    """src/motion_capture.py:274-285."""
    view_idxs: List[int]
    pose_ids: List[int]
    cost_matrix_idxs: List[int] = field(default_factory=list)   # (debug field of the reference; filled with global indices)

    def __len__(self):
        return len(self.view_idxs)


@dataclass
class SpatialTimeMatch:
    """What associate_tracking returns in the reference (src/motion_capture.py:449-478)."""
    spatial_time_matches: Dict[int, SpatialMatch]
    spatial_matches: List[SpatialMatch]
    tlet_matrix_idxs: Dict[int, int] = field(default_factory=dict)
    view_pose_matrix_idxs: Dict[int, List[Tuple[int, int]]] = field(default_factory=dict)
    dst_mat: np.ndarray = None
    sim_mat: np.ndarray = None
    match_mat: np.ndarray = None

    def find_spatial_match(self, view_idx, pose_id) -> Optional[SpatialMatch]:
        for s_match in list(self.spatial_matches) + list(self.spatial_time_matches.values()):
            if (view_idx, pose_id) in zip(s_match.view_idxs, s_match.pose_ids):
                return s_match
        return None

    def find_matrix_idx_from_view_pose_id(self, view_idx, pose_id):
        for pid, mat_idx in self.view_pose_matrix_idxs.get(view_idx, []):
            if pid == pose_id:
                return mat_idx
        return None


def _pack_frames(frames: List[FrameData], Pmax):
    C = len(frames)
    kps = np.zeros((1, C, Pmax, 17, 3))
    n_pose = np.zeros((1, C), dtype=np.int32)
    for v, f in enumerate(frames):
        for pid, pose in f.poses.items():
            kps[0, v, pid, :, :2] = pose.keypoints
            kps[0, v, pid, :, 2] = np.asarray(pose.keypoints_score).reshape(-1)
            n_pose[0, v] = max(n_pose[0, v], pid + 1)
    return kps, n_pose


def associate_tracking(tlets, frames: List[FrameData], min_pixel_error_hard_threshold=None) -> SpatialTimeMatch:
    """Same-signature seam of src/motion_capture.py:829-835 (match_spatial_time :634-808 when there are tracklets,
    match_spatial :597-631 when there are none), every stage on the device: pose layout (mvmc_prepare), distance and
    similarity matrices (mvmc_affinity), the ALS matcher (mvmc_match_als), closure + parse + group decoding
    (mvmc_assign_listed). `tlets`: objects with `.last_pose_3d.keypoints` (18, 3), e.g. the reference's own MvTracklets.
    Poses that fail filter_bad_pose are dropped by mvmc_prepare (the reference filters in run_main before update_4d, so on
    its inputs nothing changes). The threshold argument is unused, as in the reference."""
    import torch
    from multiview_motion_capture_b200 import _lib, stages
    dev = _lib.default_device()
    C = len(frames)
    T = len(tlets)
    pm = max([max(f.poses.keys(), default=-1) + 1 for f in frames] + [1])
    if pm > _lib.MAX_POSES or T > _lib.MAX_TRACKS or C > _lib.MAX_VIEWS:
        raise ValueError(f"associate_tracking: at most {_lib.MAX_VIEWS} views, {_lib.MAX_POSES} pose ids per view, {_lib.MAX_TRACKS} tracklets")
    Pmax, Tmax = pm, max(T, 1)
    kps, n_pose = _pack_frames(frames, Pmax)
    # a pose id missing from a frame's dict (filtered out) is an all-zero pose: mvmc_prepare drops it like filter_bad_pose
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    K = np.stack([f.calib.K for f in frames])[None]
    Rt = np.stack([f.calib.Rt for f in frames])[None]
    P = np.stack([f.calib.P for f in frames])[None]
    trk = np.zeros((1, Tmax, 18, 3))
    for i, tl in enumerate(tlets):
        trk[0, i] = np.asarray(tl.last_pose_3d.keypoints)[:, :3]
    n_trk = t([T], torch.int32)
    prep = stages.prepare(t(kps), t(n_pose, torch.int32), n_trk, Tmax)
    dst, sim = stages.affinity(t(kps), t(P), stages.fundamental(t(P)), stages.fundamental_krt(t(K), t(Rt)), t(trk), n_trk, prep)
    dg = prep["dim_groups"].cpu().numpy()[0]
    n = int(dg[-1])
    N = prep["N"]
    rmax = -(-min(N, 2 * max(Pmax, Tmax)) // 16) * 16
    xbin, _ = stages.match_als(sim, prep["dim_groups"], rmax, f32_first_iter=t([int(T == 0)], torch.int32))
    max_new = max(4, C * Pmax)
    a = {k: v.cpu().numpy()[0] for k, v in stages.assign(xbin, prep, n_trk, C, max_new).items()}
    if a["err"] != 0:
        raise _lib.MvmcError("associate_tracking: assignment capacity exceeded")
    iv, ip = prep["idx_view"].cpu().numpy()[0], prep["idx_pose"].cpu().numpy()[0]
    gidx = {(int(iv[g]), int(ip[g])): g for g in range(T, n)}
    mk = lambda pairs: SpatialMatch([int(v) for v, _ in pairs], [int(p) for _, p in pairs], [gidx[(int(v), int(p))] for v, p in pairs])
    out = SpatialTimeMatch({}, [])
    for ti in range(T):
        if a["trk_nsel"][ti] > 0:
            out.spatial_time_matches[ti] = mk(a["trk_sel"][ti, :a["trk_nsel"][ti]])
    groups = [(int(a["new_seq"][k]), a["new_sel"][k, :a["new_nsel"][k]]) for k in range(int(a["new_n"]))]
    groups += [(int(sq), np.array([[v, p]])) for sq, v, p in a["singles"][:min(int(a["counts"][1]), max_new)]]
    out.spatial_matches = [mk(pairs) for _, pairs in sorted(groups, key=lambda g: g[0])]
    for g in range(T, n):
        out.view_pose_matrix_idxs.setdefault(int(iv[g]), []).append((int(ip[g]), g))
    out.dst_mat = dst[0, :n, :n].cpu().numpy()
    out.sim_mat = sim[0, :n, :n].cpu().numpy()
    if T == 0:                           # the reference's no-track matrices are float32 (src/mv_math_util.py:322)
        out.dst_mat, out.sim_mat = out.dst_mat.astype(np.float32), out.sim_mat.astype(np.float32)
    out.match_mat = stages.unpack_xbin(xbin[0], n)
    out.n_truncated = int(a["counts"][2])
    return out


def filter_bad_pose(frame: FrameData, min_kps_score, n_min_valid_kps, min_bbox_size) -> FrameData:
    """src/motion_capture.py:1023-1043 on the device (mvmc_prepare's keep mask). Reference thresholds only."""
    import torch
    from multiview_motion_capture_b200 import _lib, stages
    if (min_kps_score, n_min_valid_kps, min_bbox_size) != (0.01, 4, 5):
        raise ValueError("filter_bad_pose: the kernel implements the reference's thresholds (0.01, 4, 5)")
    if not frame.poses:
        return frame
    dev = _lib.default_device()
    Pmax = max(frame.poses) + 1
    kps, n_pose = _pack_frames([frame], Pmax)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    keep = stages.prepare(t(kps), t(n_pose, torch.int32), t([0], torch.int32), 1)["keep"].cpu().numpy()[0, 0]
    for pid in [p for p in frame.poses if not keep[p]]:
        del frame.poses[pid]
    return frame


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


_B25_TO_COCO = [0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11]   # pose_def.py:262-270


class _LazyPoses:
    """`FrameData.poses` view over a packed COCO array [P,17,3]: Pose objects are only built for the poses a track
    actually used (the result holders need them; the arithmetic never does)."""

    def __init__(self, coco, n):
        self._c, self._n = coco, n

    def __getitem__(self, pid):
        k = self._c[pid]
        return Pose(KpsFormat.COCO, keypoints=k[:, :2].copy(), keypoints_score=k[:, 2:3].copy(), box=None)

    def keys(self):
        return range(self._n)

    def __len__(self):
        return self._n


def _param_of(rec_track) -> PoseShapeParam:
    x = rec_track["param"]
    return PoseShapeParam(root=x[:3].copy(), euler_angles=x[3:57].reshape(18, 3).copy(), bone_lens=x[57:68].copy())


class MvTracker:
    """MvTracker (src/motion_capture.py:838-963) for a batch of independent clips. `update_4d` keeps the reference's
    single-clip signature; `update_4d_batch` advances every clip of the batch by one frame in one device step."""

    def __init__(self, skeleton: Skeleton, n_clips=1, n_views=None, max_poses=None, max_tracks=None, device=None):
        """Device capacities are fixed when the first frame is processed. Defaults are the library's maxima
        (MVMC_MAX_POSES = 32 pose ids per view, MVMC_MAX_TRACKS = 64 alive tracks, 64 births per frame), so that a crowd
        that grows after frame 1 cannot overflow them (the reference has no caps); run_main passes the exact sizes from
        a pre-scan of the clip instead. Smaller explicit values save device memory."""
        self.skel = skeleton
        self.n_clips = n_clips
        self._cfg = dict(n_views=n_views, max_poses=max_poses, max_tracks=max_tracks, device=device)
        self.n_truncated = 0
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
        from multiview_motion_capture_b200._lib import MAX_POSES, MAX_TRACKS
        C = self._cfg["n_views"] or len(frames_per_clip[0])
        Pmax = self._cfg["max_poses"] or MAX_POSES
        Tmax = self._cfg["max_tracks"] or MAX_TRACKS
        self._cb = ClipBatch(self.n_clips, C, Pmax, max_tracks=Tmax, max_new=MAX_TRACKS, device=self._cfg["device"])

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

    def update_4d_body25(self, frm_idx: int, kps25, n_people, calibs_per_clip):
        """One frame of every clip straight from packed OpenPose BODY_25 arrays (kps25 [B,C,P,25,3], n_people [B,C]; what
        `--mode prepare` packs into clip.npz): BODY_25 -> COCO and the pose filter run on the device
        (mvmc_clips_step_body25_host). calibs_per_clip: B x C Calib objects."""
        from multiview_motion_capture_b200._lib import MAX_POSES, MAX_TRACKS
        from multiview_motion_capture_b200.clips import ClipBatch
        kps25 = np.asarray(kps25, dtype=np.float64)
        B, C, P = kps25.shape[:3]
        assert B == self.n_clips
        if self._cb is None:
            self._cb = ClipBatch(B, C, self._cfg["max_poses"] or MAX_POSES, max_tracks=self._cfg["max_tracks"] or MAX_TRACKS,
                                 max_new=MAX_TRACKS, device=self._cfg["device"])
        cb = self._cb
        if not self._calib_set:
            cb.set_calib(np.stack([np.stack([c.K for c in cs]) for cs in calibs_per_clip]),
                         np.stack([np.stack([c.Rt for c in cs]) for cs in calibs_per_clip]),
                         np.stack([np.stack([c.P for c in cs]) for cs in calibs_per_clip]))
            self._calib_set = True
        if P > cb.Pmax:
            raise ValueError(f"{P} pose slots exceed max_poses={cb.Pmax}")
        buf = np.zeros((B, C, cb.Pmax, 25, 3))
        buf[:, :, :P] = kps25
        recs = cb.step_body25(buf, n_people, frm_idx)
        coco = buf[:, :, :, _B25_TO_COCO, :]
        for b in range(B):
            frames = [FrameData(frm_idx, _LazyPoses(coco[b, v], int(n_people[b][v])), calibs_per_clip[b][v], view_id=v + 1)
                      for v in range(C)]
            self._fold(b, frm_idx, frames, recs[b])

    def _fold(self, b, frm_idx, frames, rec):
        tl = self._tlets[b]
        n = int(rec["n_alive"])
        self.n_dup_view += int(rec["n_dup_view"])
        if int(rec["n_truncated"]):
            # a group beyond even the overflow capacities (MVMC_MAX_BIG groups of MVMC_MAX_GROUP poses per frame) was cut to its
            # first MVMC_MAX_SEL poses; the reference would solve the birth from all of them (include/mvmc.h)
            self.n_truncated += int(rec["n_truncated"])
            import warnings
            warnings.warn(f"frame {frm_idx}, clip {b}: {int(rec['n_truncated'])} association group(s) held more than "
                          f"{len(rec['tracks']['sel'][0])} poses and were cut to that many before the birth solve "
                          f"(the reference solves from all of them)")
        alive_ids = []
        big, k_born = None, 0
        for t in rec["tracks"][:n]:
            tid = int(t["track_id"])
            alive_ids.append(tid)
            sel = [(int(v), int(p)) for v, p in t["sel"][:min(int(t["n_sel"]), len(t["sel"]))]]
            if t["updated"] == 2:
                if int(t["n_sel"]) > len(t["sel"]):
                    # born from more poses than a record lists (a crowded no-track frame: the reference solves the new track
                    # from every pose of the group): the full list comes from the library's overflow table
                    big = self._cb.read_big_groups(b) if big is None else big
                    sel = [(int(v), int(p)) for v, p in big[k_born]]
                k_born += 1
            if t["updated"] == 2:
                cam_poses = [(v, frames[v].poses[p]) for v, p in sel]
                pose = Pose(KpsFormat.BASIC_18, t["joints"].reshape(18, 3).copy(), np.ones((18, 1)), None)
                tl[tid] = MvTracklet(frm_idx, cam_poses, [frames[v].calib.P for v, _ in sel], [frames[v].calib for v, _ in sel],
                                     self.skel, _param_of(t), pose, n_inits=3, max_age=0)
                tl[tid].pose_ids_2d = [[p for _, p in sel]]     # (extra attribute: pose ids, for the .npz side format)
            elif t["updated"] == 1:
                trk = tl[tid]
                trk.frame_idxs.append(frm_idx)
                trk.pose_ids_2d.append([p for _, p in sel])
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
    from multiview_motion_capture_b200.ingest import load_calib_arrays
    K, Rt, wh = load_calib_arrays(Path(path))
    return Calib(K=K, Rt=Rt, P=K @ Rt, Kr_inv=Rt[:3, :3].T @ np.linalg.inv(K), img_wh_size=wh)


def _poses_from_body25(kps25, n) -> Dict[int, Pose]:
    poses = {}
    for idx in range(n):
        kps = conversion_openpose_25_to_coco(kps25[idx])
        poses[idx] = Pose(KpsFormat.COCO, keypoints=kps[:, :2], keypoints_score=kps[:, -1][:, np.newaxis], box=None)
    return poses


def parse_openpose_kps(path) -> Dict[int, Pose]:
    """src/motion_capture.py:974-984, the JSON scanned by the library's native parser (mvmc_parse_openpose_host)."""
    from multiview_motion_capture_b200.ingest import parse_openpose_text
    from multiview_motion_capture_b200._lib import MAX_POSES
    with open(path, "rb") as f:
        kps25, n = parse_openpose_text(f.read(), 4 * MAX_POSES)
    return _poses_from_body25(kps25, n)


def extract_frame_data_from_openpose(in_dir: Path, calib_dir: Path, out_data_dir: Path, write_pickles=True):
    """src/motion_capture.py:987-1005. Camera directories sorted by stem, files by int(stem.split('_')[1]), calibration
    matched by stem (json or pkl). Writes the reference's per-frame pickles AND the packed `clip.npz` that `--mode run`
    reads directly (no per-frame pickle round trip)."""
    from multiview_motion_capture_b200.ingest import pack_openpose_clip, save_clip_npz
    clip = pack_openpose_clip(in_dir, calib_dir)
    os.makedirs(out_data_dir, exist_ok=True)
    save_clip_npz(Path(out_data_dir) / "clip.npz", clip)
    calibs = _calibs_of(clip)
    F, C = clip["n_pose"].shape
    if write_pickles:
        for frm in range(F):
            frames = [FrameData(frm, _poses_from_body25(clip["kps25"][frm, v], int(clip["n_pose"][frm, v])), calibs[v], view_id=v + 1)
                      for v in range(C)]
            with open(Path(out_data_dir) / f"{frm:06d}.pkl", "wb") as f:
                pickle.dump(frames, f)
    return F


def _calibs_of(clip) -> List[Calib]:
    return [Calib(K=K, Rt=Rt, P=K @ Rt, Kr_inv=Rt[:3, :3].T @ np.linalg.inv(K), img_wh_size=[int(x) for x in wh])
            for K, Rt, wh in zip(clip["K"], clip["RT"], clip["img_wh"])]


# ---- run mode (src/motion_capture.py:1046-1129) ----------------------------------------------------------------------
def _clip_dirs(pose_dir: Path) -> List[Path]:
    has = lambda d: any(d.glob("*.pkl")) or (d / "clip.npz").exists()
    if has(Path(pose_dir)):
        return [Path(pose_dir)]
    return sorted(d for d in Path(pose_dir).iterdir() if d.is_dir() and has(d))


def _write_outputs(tracker, clips, out_dir):
    from multiview_motion_capture_b200.tracklets_io import save_tracklets_npz
    for b, c in enumerate(clips):
        dst = Path(out_dir) if len(clips) == 1 else Path(out_dir) / c.name
        os.makedirs(dst, exist_ok=True)
        tlets = tracker.finish(b)
        with open(dst / "tracklets.pkl", "wb") as f:
            pickle.dump(file=f, obj={"tracklets": tlets})
        save_tracklets_npz(dst / "tracklets.npz", tlets)


def run_main(video_dir: Path, pose_dir: Path, out_dir: Path, max_frames: int = 300, device=None):
    """The reference's run loop (frames 1..min(n-1, max_frames); frame 0 is skipped, :1063 precedes :1077) over one clip
    or a directory of clips. A clip directory holds the reference's per-frame pickles and/or the packed clip.npz of
    `--mode prepare` (preferred: no unpickling, BODY_25 -> COCO on the device). Device capacities come from a pre-scan of
    the inputs. Whatever was tracked is written even if a later frame fails."""
    from multiview_motion_capture_b200.ingest import load_clip_npz
    clips = _clip_dirs(pose_dir)
    if not clips:
        raise SystemExit(f"no frame pickles or clip.npz under {pose_dir}")
    packed = [load_clip_npz(c / "clip.npz") if (c / "clip.npz").exists() else None for c in clips]
    use_packed = all(p is not None for p in packed)
    if use_packed:
        n_frames = min(p["n_pose"].shape[0] for p in packed)
        pm = max(int(p["n_pose"][:n_frames].max()) for p in packed)
        calibs = [_calibs_of(p) for p in packed]
    else:
        paths = [sorted(c.glob("*.pkl"), key=lambda p: int(p.stem)) for c in clips]
        n_frames = min(len(p) for p in paths)
        pm = None     # pose ids are only known once the pickles are read: the tracker then uses the library's maximum
    tracker = MvTracker(load_skeleton(), n_clips=len(clips), device=device, max_poses=max(pm, 1) if pm is not None else None)
    n_test = min(n_frames, max_frames)
    frm_idx = 0
    try:
        while True:
            frm_idx += 1
            if frm_idx >= n_frames:
                break
            if use_packed:
                P = max(p["kps25"].shape[2] for p in packed)
                k = np.zeros((len(clips), packed[0]["kps25"].shape[1], P, 25, 3))
                for b, p in enumerate(packed):
                    k[b, :, :p["kps25"].shape[2]] = p["kps25"][frm_idx]
                tracker.update_4d_body25(frm_idx, k, np.stack([p["n_pose"][frm_idx] for p in packed]), calibs)
            else:
                frames = []
                for p in paths:
                    with open(p[frm_idx], "rb") as f:
                        frames.append(pickle.load(f))
                tracker.update_4d_batch(frm_idx, frames)
            if frm_idx >= n_test:
                break
    finally:
        if tracker._cb is not None:
            _write_outputs(tracker, clips, out_dir)
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
