"""Device-agnostic replay of a golden scene through the clip-batch pipeline (mvmc_clips_*), teacher-forced (track
table taken from the reference before every frame) or free-running. Used by the GPU tier on full scenes and by the
CPU tier (kernel emulator) on a couple of frames."""
import numpy as np

import mvmc_oracle as o
from helpers import GoldenTable, fkey, golden, pad_poses


def run_golden_clip(dev, name, Pmax, Tmax, forced, B=1, max_new=8, frames=None):
    from multiview_motion_capture_b200.clips import ClipBatch
    inp, g = golden(name)
    kps = o.body25_to_coco(inp["kps25"])
    C = kps.shape[1]
    cb = ClipBatch(B, C, Pmax, max_tracks=Tmax, max_new=max_new, device=dev)
    cb.set_calib(np.repeat(inp["K"][None], B, 0), np.repeat(inp["RT"][None], B, 0))
    tab = GoldenTable(g)
    st = dict(frames=0, xbin=0, iters=0, alive=0, n_alive=0, upd=0, dj=[], replicas=0, first_mismatch=10 ** 9)
    for f in (frames or range(int(g["first_frame"]), int(g["last_frame"]) + 1)):
        k = fkey(f)
        if forced:
            cb.set_tracks(**tab.packed(f, B, Tmax))
        recs = cb.step(np.repeat(pad_poses(kps[f], Pmax)[None], B, 0), np.repeat(inp["n_pose"][f][None], B, 0), f)
        rec = recs[0].copy()
        dst, sim, xb, dg = cb.read_matrices(0)
        n_alive = int(rec["n_alive"])
        tr = rec["tracks"][:n_alive]
        upd = tr[tr["updated"] > 0]
        st["frames"] += 1
        same_x = xb.shape == g[k + "xbin"].shape and np.array_equal(xb, g[k + "xbin"].astype(bool))
        st["xbin"] += int(same_x)
        st["iters"] += int(rec["als_iters"] == int(g[k + "als_iters"]))
        st["alive"] += int(tr["track_id"].tolist() == g[k + "alive_after"].tolist())
        st["n_alive"] += int(n_alive == len(g[k + "alive_after"]))
        if not (same_x and tr["track_id"].tolist() == g[k + "alive_after"].tolist()):
            st["first_mismatch"] = min(st["first_mismatch"], f)
        same_upd = upd["track_id"].tolist() == g[k + "upd_ids"].tolist()
        st["upd"] += int(same_upd)
        # No-track frames run the reference's float32 affinity (NumPy float32 mean / std / exp): restated bit for bit on the
        # device (csrc/affinity.cu np32), so X_bin is asserted there too, even where the reference's ALS stops at its 1000-iteration cap.
        st.setdefault("unstable_frames", [])
        if len(g[k + "alive_before"]) == 0 and int(g[k + "als_iters"]) >= 1000:
            st["unstable_frames"].append((f, bool(same_x)))
        if forced:
            assert same_x, (name, f, "X_bin")
            assert rec["n_dup_view"] == int(g[k + "printed"])
            assert tr["track_id"].tolist() == g[k + "alive_after"].tolist(), (name, f, "track ids")
            state = np.stack([tr["state"], tr["hits"], tr["time_since_update"], tr["length"]], 1).reshape(-1, 4)
            assert np.array_equal(state, g[k + "alive_state"]), (name, f, "lifecycle counters")
            assert same_upd, (name, f)
            # the (view, pose) pairs each updated track was solved from
            for u, t in enumerate(upd):
                views = sorted({int(v) for v in t["sel"][:t["n_sel"], 0]})   # no-track groups may hold >1 pose per view
                assert views == np.nonzero(g[k + "upd_views"][u])[0].tolist(), (name, f, u)
        if same_upd and len(upd):
            st["dj"].extend(np.abs(upd["joints"].reshape(-1, 18, 3) - g[k + "upd_joints"]).max(axis=(1, 2)).tolist())
        if B > 1:
            st["replicas"] += int(all(recs[b].tobytes() == recs[0].tobytes() for b in range(1, B)))
    cb.close()
    return st




def check_clip_streams_equal_clip_batch(dev, name="synth_c4p3", Pmax=4, Tmax=8, max_new=4, frames=(2, 6), B=3, groups=2):
    """ClipStreams (groups of clips on their own streams, mvmc_clips_step_host_async) returns, clip for clip, the records
    ClipBatch returns: B clips with different inputs (the golden clip with its camera views rotated), the track table of
    every frame taken from the golden run (tracking frames: the matcher converges in tens of iterations)."""
    import torch
    from multiview_motion_capture_b200._lib import STEP_OUT_DTYPE
    from multiview_motion_capture_b200.clips import ClipBatch, ClipStreams
    inp, g = golden(name)
    kps = o.body25_to_coco(inp["kps25"])
    C = kps.shape[1]
    rot = [np.roll(np.arange(C), b) for b in range(B)]           # clip b sees the cameras in a rotated order
    K = np.stack([inp["K"][r] for r in rot])
    RT = np.stack([inp["RT"][r] for r in rot])
    cb = ClipBatch(B, C, Pmax, max_tracks=Tmax, max_new=max_new, device=dev)
    cs = ClipStreams(B, C, Pmax, groups=groups, max_tracks=Tmax, max_new=max_new, device=dev)
    cb.set_calib(K, RT)
    cs.set_calib(K, RT)
    tab = GoldenTable(g)
    pin = (lambda t: t.pin_memory()) if str(dev).startswith("cuda") else (lambda t: t)
    out = pin(torch.empty(B * STEP_OUT_DTYPE.itemsize, dtype=torch.uint8))
    for f in frames:
        kf = np.stack([pad_poses(kps[f][r], Pmax) for r in rot])
        nf = np.stack([inp["n_pose"][f][r] for r in rot]).astype(np.int32)
        cb.set_tracks(**tab.packed(f, B, Tmax))
        cs.set_tracks(**tab.packed(f, B, Tmax))
        ref = cb.step(kf, nf, f).copy()
        cs.step_host(pin(torch.from_numpy(np.ascontiguousarray(kf))), pin(torch.from_numpy(np.ascontiguousarray(nf))), f, out)
        got = out.numpy().view(STEP_OUT_DTYPE)
        assert got.tobytes() == ref.tobytes(), (name, f)
        assert len({ref[b].tobytes() for b in range(B)}) > 1, "the clips of this test must differ"
    cb.close()
    cs.close()


def _oracle_table(trk, Tmax):
    """The oracle tracker's alive-track table packed for ClipBatch.set_tracks (one clip)."""
    a = dict(n_trk=np.array([len(trk.tracks)], np.int32), ids=np.zeros((1, Tmax), np.int32), state=np.zeros((1, Tmax), np.int32),
             hits=np.zeros((1, Tmax), np.int32), tsu=np.zeros((1, Tmax), np.int32), length=np.zeros((1, Tmax), np.int32),
             param=np.zeros((1, Tmax, 68)), joints=np.zeros((1, Tmax, 54)), next_id=np.array([trk.next_id], np.int32))
    assert len(trk.tracks) <= Tmax
    for i, t in enumerate(trk.tracks):
        a["ids"][0, i], a["state"][0, i], a["hits"][0, i] = t.track_id, t.state, t.hits
        a["tsu"][0, i], a["length"][0, i] = t.time_since_update, len(t)
        a["param"][0, i] = t.params[-1].pack()
        a["joints"][0, i] = t.joints[-1].reshape(-1)
    return a


def run_side_by_side_with_oracle(dev, n_views, n_people, n_clips, n_frames, seed, Tmax, first_frame=3, shelf=False):
    """bench.py's workload, checked: `n_clips` distinct synthetic clips of the benchmarked shape, the oracle's tracker
    seeded from the generator's ground truth exactly as bench.py's CPU arm is (`_cpu_worker`), then `n_frames` tracked
    frames with the CUDA pipeline teacher-forced from the oracle's table before every frame. Asserts per clip-frame:
    kept poses / index layout identical, dst <= 1e-7 px, sim <= 1e-9, X_bin and the ALS stopping iteration bit-exact,
    matched (view, pose) sets, births, deaths, track ids and lifecycle counters identical. Returns the IK differences."""
    from multiview_motion_capture_b200 import synthetic as S
    from multiview_motion_capture_b200.clips import ClipBatch
    from helpers import golden
    st = dict(frames=0, dst=0.0, sim=0.0, dj=[], dparam=[], dcost=[], n=[], iters=[], nfev_same=0, solves=0)
    shelf_calib = None
    if shelf:
        gi, _ = golden("shelf")
        shelf_calib = (gi["K"], gi["RT"], gi["img_wh"][0])
    for ci in range(n_clips):
        c = S.make_clip(n_views, n_people, first_frame + n_frames, seed=seed, clip_idx=ci, shelf_calib=shelf_calib)
        C = len(c["K"])
        kps = S.body25_to_coco(c["kps25"])
        trk = o.Tracker(o.projections(c["K"], c["RT"]), c["K"], c["RT"])
        f0 = first_frame - 1
        for pi in range(n_people):
            prm = o.PoseParam(c["gt_root"][f0, pi].copy(), c["gt_euler"][f0, pi].copy(), trk.skel.side_bone_lens * c["gt_scale"][pi])
            joints, _ = o.forward_kinematics(trk.skel, prm.root, prm.euler, prm.bone_lens)
            trk.tracks.append(o.Track(pi, [f0], [prm], [joints], [[]], state=o.CONFIRMED, hits=3))
        trk.next_id = n_people
        cb = ClipBatch(1, C, n_people, max_tracks=Tmax, max_new=n_people, device=dev)
        cb.set_calib(c["K"][None], c["RT"][None])
        for f in range(first_frame, first_frame + n_frames):
            cb.set_tracks(**_oracle_table(trk, Tmax))
            before = [(t.track_id, t.params[-1]) for t in trk.tracks]
            a = trk.step(f, kps[f], c["n_pose"][f])
            rec = cb.step(kps[f][None], c["n_pose"][f][None], f)[0].copy()
            dst, sim, xb, dg = cb.read_matrices(0)
            tag = (n_views, n_people, ci, f)
            assert dst.shape == a.dst.shape, tag
            # 1e-7 px, relative 1e-12 where a distance is huge (a track re-projected from nearly behind a camera: 1e5 px)
            dd = float((np.abs(dst - a.dst) / np.maximum(1.0, np.abs(a.dst) * 1e-5)).max())
            ds = float(np.abs(sim - a.sim).max())
            st["dst"], st["sim"] = max(st["dst"], dd), max(st["sim"], ds)
            assert dd <= 1e-7 and ds <= 1e-9, tag + (dd, ds, float(np.abs(a.dst).max()))
            assert np.array_equal(xb, a.x_bin), tag + ("X_bin",)
            assert int(rec["als_iters"]) == a.n_iter, tag + ("ALS iterations", int(rec["als_iters"]), a.n_iter)
            n_alive = int(rec["n_alive"])
            tr = rec["tracks"][:n_alive]
            assert tr["track_id"].tolist() == [t.track_id for t in trk.tracks], tag + ("track ids",)
            state = np.stack([tr["state"], tr["hits"], tr["time_since_update"], tr["length"]], 1).reshape(-1, 4)
            assert np.array_equal(state, np.array([[t.state, t.hits, t.time_since_update, len(t)] for t in trk.tracks]).reshape(-1, 4)), tag
            assert sorted(rec["died_ids"][:rec["n_died"]].tolist()) == sorted(t.track_id for t in trk.dead if t.frame_idxs[-1] < f
                                                                               and t.track_id in [b[0] for b in before]), tag
            assert int(rec["n_dup_view"]) == a.n_dup_view, tag
            upd_o = [t for t in trk.tracks if t.frame_idxs[-1] == f]
            upd = tr[tr["updated"] > 0]
            assert upd["track_id"].tolist() == [t.track_id for t in upd_o], tag
            log = [e for e in trk.solve_log if e[0] in ("ik1", "ik2")]
            for u, (t, ot) in enumerate(zip(upd, upd_o)):
                sel = [tuple(x) for x in t["sel"][:t["n_sel"]].tolist()]
                assert sel == [tuple(x) for x in ot.views[-1]], tag + ("matched (view, pose)", u)
                st["dj"].append(float(np.abs(t["joints"].reshape(18, 3) - ot.joints[-1]).max()))
                st["dparam"].append(float(np.abs(t["param"] - ot.params[-1].pack()).max()))
                r1, r2 = log[2 * u][2], log[2 * u + 1][2]
                st["solves"] += 2
                st["nfev_same"] += int(t["nfev"][0] == r1.nfev) + int(t["nfev"][1] == r2.nfev)
                st["dcost"].append(float(abs(t["cost"][1] - r2.cost) / max(r2.cost, 1e-300)))
            st["frames"] += 1
            st["n"].append(int(dg[-1]))
            st["iters"].append(a.n_iter)
        cb.close()
    return st


def check_edge_cases(dev, full=True):
    """Edge cases of the per-frame path against the oracle's tracker on the same inputs (a 4-camera, 3-person golden scene,
    tracks seeded from the reference's table): an empty frame (no pose in any view: every track is marked missed and dies,
    max_age = 0), a frame after it (no tracks left: the float32 no-track path, births with fresh ids), a frame with poses
    in ONE view only (nothing can be matched by two views: no update, no birth), a ragged frame (one view empty, one view
    with a single pose), and poses that fail the filter (all-zero scores). Compared: X_bin, ALS iterations, alive ids,
    lifecycle counters, died ids, updated ids."""
    from multiview_motion_capture_b200.clips import ClipBatch
    inp, g = golden("synth_c4p3")
    kps_all = o.body25_to_coco(inp["kps25"])
    C, Pmax, Tmax = kps_all.shape[1], 4, 8
    cb = ClipBatch(1, C, Pmax, max_tracks=Tmax, max_new=4, device=dev)
    cb.set_calib(inp["K"][None], inp["RT"][None])
    tab = GoldenTable(g)
    trk = o.Tracker(o.projections(inp["K"], inp["RT"]), inp["K"], inp["RT"])
    f0 = 3
    pk = tab.packed(f0, 1, Tmax)
    cb.set_tracks(**pk)
    for i in range(int(pk["n_trk"][0])):
        prm = o.PoseParam.unpack(pk["param"][0, i].copy())
        L = int(pk["length"][0, i])
        trk.tracks.append(o.Track(int(pk["ids"][0, i]), list(range(f0 - L, f0)), [prm] * L, [pk["joints"][0, i].reshape(18, 3).copy()] * L,
                                  [[]] * L, state=int(pk["state"][0, i]), hits=int(pk["hits"][0, i]),
                                  time_since_update=int(pk["tsu"][0, i])))
    trk.next_id = int(pk["next_id"][0])

    def frame(kind, f):
        k = pad_poses(kps_all[f], Pmax).copy()
        n = inp["n_pose"][f].copy().astype(np.int32)
        if kind == "empty":
            n[:] = 0
        elif kind == "one_view":
            n[1:] = 0
        elif kind == "ragged":
            n[0] = 0
            n[1] = 1
        elif kind == "filtered":
            k[:, :, :, 2] = 0.0          # every pose fails filter_bad_pose (no keypoint above the score threshold)
        return k, n

    seen = []
    steps = [("normal", 3), ("one_view", 4), ("ragged", 5), ("empty", 6)]
    if full:
        steps += [("normal", 7), ("filtered", 8), ("normal", 8)]      # (re-births after everything died: 50-evaluation solves)
    for step, (kind, f) in enumerate(steps):
        k, n = frame(kind, f)
        fi = 100 + step
        ids_before = [t.track_id for t in trk.tracks]
        a = trk.step(fi, k, n)
        rec = cb.step(k[None], n[None], fi)[0].copy()
        _, _, xb, dg = cb.read_matrices(0)
        tag = (kind, f)
        assert xb.shape == a.x_bin.shape and np.array_equal(xb, a.x_bin), tag + ("X_bin",)
        if a.x_bin.size:
            assert int(rec["als_iters"]) == a.n_iter, tag + (int(rec["als_iters"]), a.n_iter)
        na = int(rec["n_alive"])
        tr = rec["tracks"][:na]
        assert tr["track_id"].tolist() == [t.track_id for t in trk.tracks], tag + (tr["track_id"].tolist(), [t.track_id for t in trk.tracks])
        st = np.stack([tr["state"], tr["hits"], tr["time_since_update"], tr["length"]], 1).reshape(-1, 4)
        assert np.array_equal(st, np.array([[t.state, t.hits, t.time_since_update, len(t)] for t in trk.tracks]).reshape(-1, 4)), tag
        died = sorted(set(ids_before) - {t.track_id for t in trk.tracks})
        assert sorted(rec["died_ids"][:rec["n_died"]].tolist()) == died, tag
        assert tr[tr["updated"] > 0]["track_id"].tolist() == [t.track_id for t in trk.tracks if t.frame_idxs[-1] == fi], tag
        assert rec["error"] == 0
        seen.append((kind, na, len(died)))
    cb.close()
    kinds = dict((k, (na, nd)) for k, na, nd in seen)
    assert kinds["empty"][0] == 0 and kinds["empty"][1] > 0          # everything died on the empty frame
    # (a one-view match is not solved but still counts as "matched": the track is not marked missed - the reference's quirk,
    #  src/motion_capture.py:925-934 - so nothing has to die on the single-view frame)
    return seen



def check_crowded_no_track_frame(dev, n_views=8, n_people=16, seed=1000, clip=0):
    """Frame 1 of a crowded synthetic clip (no tracks yet): the reference's float32 affinity is not discriminative there, its
    ALS stops at the 1000-iteration cap and merges many poses - several per view, of several people - into a few groups,
    and it builds each new track from ALL the poses of its group (src/motion_capture.py:618-624, 942-958). Against the oracle
    on the same input: X_bin and the iteration count bit-exact, the same groups with the same complete pose lists (the
    overflow table for the ones beyond MVMC_MAX_SEL), nothing truncated, the same track ids; the births' joints and costs are
    reported (fits of one skeleton to several people: ill-posed)."""
    from multiview_motion_capture_b200 import synthetic as S
    from multiview_motion_capture_b200._lib import MAX_SEL
    from multiview_motion_capture_b200.clips import ClipBatch
    c = S.make_clip(n_views, n_people, 3, seed=seed, clip_idx=clip)
    kps = S.body25_to_coco(c["kps25"])
    trk = o.Tracker(o.projections(c["K"], c["RT"]), c["K"], c["RT"])
    a = trk.step(1, kps[1], c["n_pose"][1])
    cb = ClipBatch(1, n_views, n_people, max_tracks=2 * n_people, max_new=2 * n_people, device=dev)
    cb.set_calib(c["K"][None], c["RT"][None])
    rec = cb.step(kps[1][None], c["n_pose"][1][None], 1)[0].copy()
    _, _, xb, _ = cb.read_matrices(0)
    assert np.array_equal(xb, a.x_bin) and int(rec["als_iters"]) == a.n_iter, "X_bin / iterations"
    born_ref = [g for g in a.new_groups if len(g) >= 2]
    tr = rec["tracks"][:int(rec["n_alive"])]
    assert rec["error"] == 0 and int(rec["n_truncated"]) == 0
    assert tr["track_id"].tolist() == [t.track_id for t in trk.tracks]
    big = cb.read_big_groups(0)
    sizes, dj, dcost = [], [], []
    for k, (t, gref, tref) in enumerate(zip(tr, born_ref, trk.tracks)):
        assert int(t["updated"]) == 2 and int(t["n_sel"]) == len(gref), (k, int(t["n_sel"]), len(gref))
        sel = big[k] if len(gref) > MAX_SEL else [tuple(x) for x in t["sel"][:len(gref)].tolist()]
        assert sel == [tuple(x) for x in gref], k
        sizes.append(len(gref))
        dj.append(float(np.abs(t["joints"].reshape(18, 3) - tref.joints[-1]).max()))
    cb.close()
    return dict(n=int(a.x_bin.shape[0]), als_iters=a.n_iter, group_sizes=sizes, n_big=len(big), joints_diff_m=dj)
