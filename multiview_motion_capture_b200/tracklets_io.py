"""Output side format: `tracklets.npz`, a compact array form of the reference's `tracklets.pkl`
({"tracklets": List[MvTracklet]} sorted by -len, src/motion_capture.py:1120-1129) that needs no class on the reader's
side, plus a reader that rebuilds objects with the accessors the reference's `viz_tracklets` uses
(src/motion_capture.py:1177-1198: `tlet.poses` -> `(p[0], p[-1])`, `pose.pose_type`, `pose.keypoints[:, :3]`).

Arrays (T tracklets in file order, L = total frames over all tracklets, rows of a tracklet contiguous):
    offsets [T+1] int32 row range of each tracklet; state, hits, time_since_update, max_age, n_inits [T] int32
    frame_idx [L] int32; root [L,3]; euler [L,18,3]; bone_lens [L,11]; joints [L,18,3] float64
    n_views [L] int32; view_idx, pose_id [L,Vmax] int32 (-1 padded): the 2D poses every frame was solved from
    Optional (keep_2d=True): kps2d [L,Vmax,17,3] float64, the COCO poses themselves
"""
import numpy as np


def tracklets_to_arrays(tracklets, keep_2d=False):
    T = len(tracklets)
    lens = [len(t.frame_idxs) for t in tracklets]
    L = int(sum(lens))
    vmax = max([len(v) for t in tracklets for v in t.cam_poses_2d] + [1])
    a = dict(offsets=np.concatenate([[0], np.cumsum(lens)]).astype(np.int32),
             state=np.array([t.state.value for t in tracklets], dtype=np.int32).reshape(T),
             hits=np.array([t.hits for t in tracklets], dtype=np.int32).reshape(T),
             time_since_update=np.array([t.time_since_update for t in tracklets], dtype=np.int32).reshape(T),
             max_age=np.array([t.max_age for t in tracklets], dtype=np.int32).reshape(T),
             n_inits=np.array([t.n_inits for t in tracklets], dtype=np.int32).reshape(T),
             frame_idx=np.zeros(L, np.int32), root=np.zeros((L, 3)), euler=np.zeros((L, 18, 3)), bone_lens=np.zeros((L, 11)),
             joints=np.zeros((L, 18, 3)), n_views=np.zeros(L, np.int32), view_idx=np.full((L, vmax), -1, np.int32),
             pose_id=np.full((L, vmax), -1, np.int32))
    if keep_2d:
        a["kps2d"] = np.zeros((L, vmax, 17, 3))
    r = 0
    for t in tracklets:
        ids = getattr(t, "pose_ids_2d", None)
        for i, (frm, prm, pose) in enumerate(t.poses):
            a["frame_idx"][r] = frm
            a["root"][r], a["euler"][r], a["bone_lens"][r] = prm.root, prm.euler_angles, prm.bone_lens
            a["joints"][r] = pose.keypoints[:, :3]
            views = t.cam_poses_2d[i]
            a["n_views"][r] = len(views)
            for q, (v, p2) in enumerate(views):
                a["view_idx"][r, q] = v
                if ids is not None:
                    a["pose_id"][r, q] = ids[i][q]
                if keep_2d:
                    a["kps2d"][r, q, :, :2] = p2.keypoints
                    a["kps2d"][r, q, :, 2] = np.asarray(p2.keypoints_score).reshape(-1)
            r += 1
    return a


def save_tracklets_npz(path, tracklets, keep_2d=False):
    np.savez_compressed(path, **tracklets_to_arrays(tracklets, keep_2d))


class _Pose:
    """Minimal stand-in for pose_def.Pose (same attribute names)."""
    def __init__(self, pose_type, keypoints, keypoints_score):
        self.pose_type, self.keypoints, self.keypoints_score, self.box = pose_type, keypoints, keypoints_score, None


class _Param:
    def __init__(self, root, euler_angles, bone_lens):
        self.root, self.euler_angles, self.bone_lens = root, euler_angles, bone_lens


class NpzTracklet:
    """Read-only tracklet rebuilt from tracklets.npz: frame_idxs, poses [(frm, param, pose)], views [(view, pose id)],
    state/hits/... as integers. `pose_types` = (BASIC_18, COCO) enum members to tag the poses with (pass the reference's
    `pose_def.KpsFormat` members to feed its own viz code)."""

    def __init__(self, a, t, pose_types=("BASIC_18", "COCO")):
        lo, hi = int(a["offsets"][t]), int(a["offsets"][t + 1])
        self.frame_idxs = a["frame_idx"][lo:hi].tolist()
        self.poses = [(int(a["frame_idx"][r]), _Param(a["root"][r], a["euler"][r], a["bone_lens"][r]),
                       _Pose(pose_types[0], a["joints"][r], np.ones((18, 1)))) for r in range(lo, hi)]
        self.views = [[(int(a["view_idx"][r, q]), int(a["pose_id"][r, q])) for q in range(int(a["n_views"][r]))] for r in range(lo, hi)]
        self.cam_poses_2d = None
        if "kps2d" in a:
            self.cam_poses_2d = [[(int(a["view_idx"][r, q]), _Pose(pose_types[1], a["kps2d"][r, q, :, :2], a["kps2d"][r, q, :, 2:3]))
                                  for q in range(int(a["n_views"][r]))] for r in range(lo, hi)]
        self.state, self.hits = int(a["state"][t]), int(a["hits"][t])
        self.time_since_update, self.max_age, self.n_inits = int(a["time_since_update"][t]), int(a["max_age"][t]), int(a["n_inits"][t])

    def __len__(self):
        return len(self.frame_idxs)

    @property
    def last_pose_3d(self):
        return self.poses[-1][-1]


def load_tracklets_npz(path, pose_types=("BASIC_18", "COCO")):
    with np.load(path, allow_pickle=False) as z:
        a = {k: z[k] for k in z.files}
    return [NpzTracklet(a, t, pose_types) for t in range(len(a["offsets"]) - 1)]
