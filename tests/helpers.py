"""Shared test helpers: golden fixtures, golden track-table replay (teacher forcing), packing."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libmvmc_emu.so")

_cache = {}


WARM = [("warm_c5p4", 4, 8), ("warm_c8p12", 12, 16), ("warm_c8p16", 16, 24), ("warm_c8p32", 32, 40)]   # (name, Pmax, Tmax)


def oracle_tracker_from_golden(name):
    """An oracle.Tracker for a golden scene; a warm golden's tracker is seeded with the same tracks the reference's was."""
    import mvmc_oracle as o
    inp, g = golden(name)
    trk = o.Tracker(o.projections(inp["K"], inp["RT"]), inp["K"], inp["RT"])
    if "init_ids" in g.files:
        f0 = int(g["first_frame"]) - 1
        for i, tid in enumerate(g["init_ids"].tolist()):
            prm = o.PoseParam.unpack(g["init_param"][i].copy())
            trk.tracks.append(o.Track(tid, [f0], [prm], [g["init_joints"][i].copy()], [[]], state=o.CONFIRMED, hits=3))
        trk.next_id = len(trk.tracks)
    return trk


def golden(name):
    """(inputs, reference outputs) of a golden scene: 'shelf', 'synth_c4p3', 'synth_c8p6', 'synth_c8p12', and the
    warm-started 'warm_c5p4', 'warm_c8p12', 'warm_c8p16', 'warm_c8p32' (tracked frames at the BASELINE shapes)."""
    if name not in _cache:
        _cache[name] = (np.load(os.path.join(GOLD, f"{name}_inputs.npz")), np.load(os.path.join(GOLD, f"{name}_ref.npz")))
    return _cache[name]


def fkey(f):
    return f"f{f:04d}_"


class GoldenTable:
    """Replays the reference's alive-track table frame by frame from a *_ref.npz (teacher forcing): before
    frame f, `entries(f)` is what the reference's tracker held (ids in list order, last params/joints,
    lifecycle counters)."""

    def __init__(self, ref):
        self.g = ref
        self.table = {}
        self.frame = int(ref["first_frame"]) - 1
        self.init_ids = []
        if "init_ids" in ref.files:   # warm-started golden (oracle/make_golden.py warm): the tracker was seeded with these
            self.init_ids = ref["init_ids"].tolist()
            for i, tid in enumerate(self.init_ids):
                st = ref["init_state"][i]
                self.table[tid] = dict(param=ref["init_param"][i].copy(), joints=ref["init_joints"][i].copy(),
                                       state=int(st[0]), hits=int(st[1]), tsu=int(st[2]), len=int(st[3]))

    def advance(self, f):
        """Fold the reference's outputs of frame f into the table."""
        g, k = self.g, fkey(f)
        for tid in self.table:
            self.table[tid]["tsu"] += 1
        for i, tid in enumerate(g[k + "upd_ids"].tolist()):
            prm = np.concatenate([g[k + "upd_root"][i], g[k + "upd_euler"][i].reshape(-1), g[k + "upd_blens"][i]])
            self.table.setdefault(tid, {})
            self.table[tid].update(param=prm, joints=g[k + "upd_joints"][i].copy())
        after = g[k + "alive_after"].tolist()
        st = g[k + "alive_state"]
        for i, tid in enumerate(after):
            self.table[tid].update(state=int(st[i, 0]), hits=int(st[i, 1]), tsu=int(st[i, 2]), len=int(st[i, 3]))
        for tid in list(self.table):
            if tid not in after:
                del self.table[tid]
        self.frame = f

    def seek(self, f):
        """Table as it was BEFORE frame f."""
        assert f - 1 >= self.frame
        for q in range(self.frame + 1, f):
            self.advance(q)
        ids = self.g[fkey(f) + "alive_before"].tolist()
        assert sorted(ids) == sorted(self.table), (ids, sorted(self.table))
        return ids

    def next_id(self):
        return (max(self.table) + 1) if self.table else self._max_seen()

    def _max_seen(self):
        return 0

    def packed(self, f, B, Tmax):
        """Arrays for ClipBatch.set_tracks: the reference's table before frame f replicated over B clips."""
        ids = self.seek(f)
        n = len(ids)
        a = dict(n_trk=np.full(B, n, np.int32), ids=np.zeros((B, Tmax), np.int32), state=np.zeros((B, Tmax), np.int32),
                 hits=np.zeros((B, Tmax), np.int32), tsu=np.zeros((B, Tmax), np.int32), length=np.zeros((B, Tmax), np.int32),
                 param=np.zeros((B, Tmax, 68)), joints=np.zeros((B, Tmax, 54)))
        for i, tid in enumerate(ids):
            e = self.table[tid]
            a["ids"][:, i] = tid
            a["state"][:, i] = e["state"]
            a["hits"][:, i] = e["hits"]
            a["tsu"][:, i] = e["tsu"]
            a["length"][:, i] = e["len"]
            a["param"][:, i] = e["param"]
            a["joints"][:, i] = e["joints"].reshape(-1)
        # ids are handed out in creation order, so the next id is one past every id ever seen
        seen = list(self.init_ids) + [int(x) for q in range(int(self.g["first_frame"]), f)
                                      for x in self.g[fkey(q) + "alive_after"].tolist()]
        a["next_id"] = np.full(B, (max(seen) + 1) if seen else 0, np.int32)
        return a

    def joints(self, f):
        ids = self.seek(f)
        return [self.table[t]["joints"] for t in ids]

    def params(self, f):
        ids = self.seek(f)
        return [self.table[t]["param"] for t in ids]


def golden_matches(ref, f):
    """(track_matches {t_idx: [(view, pose)]}, new_groups [[(view, pose)]]) of the reference at frame f."""
    rows = ref[fkey(f) + "matches"]
    trk, new = {}, {}
    for kind, gi, t_idx, v, p in rows.tolist():
        if kind == 0:
            trk.setdefault(t_idx, []).append((v, p))
        else:
            new.setdefault(gi, []).append((v, p))
    return trk, [new[k] for k in sorted(new)]


def view_lists(kps_f, n_pose_f, kept=None):
    """Per-view kept pose ids and keypoints (C x (P_v,17,3)) of one frame, as the oracle wants them."""
    import mvmc_oracle as o
    ids, arr = [], []
    for v in range(kps_f.shape[0]):
        cur = [p for p in range(int(n_pose_f[v])) if not o.pose_is_bad(kps_f[v, p])]
        if kept is not None:
            assert cur == [p for p in range(kps_f.shape[1]) if kept[v, p]], (v, cur, kept[v])
        ids.append(cur)
        arr.append(kps_f[v, cur] if cur else np.zeros((0, 17, 3)))
    return ids, arr


def pad_poses(kps_f, Pmax):
    """(C,P,17,3) -> (C,Pmax,17,3) zero padded."""
    C, P = kps_f.shape[:2]
    out = np.zeros((C, Pmax, 17, 3))
    out[:, :P] = kps_f
    return out
