"""GPU tier: the clip-batch pipeline (mvmc_clips_*) — association + IK + track lifecycle on the device — against
the reference goldens, teacher-forced and free-running, plus size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

import mvmc_oracle as o
from helpers import WARM, GoldenTable, fkey, golden, golden_matches, pad_poses

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


from pipeline_checks import run_golden_clip as _run_dev


def _run(name, Pmax, Tmax, forced, B=1, max_new=8):
    return _run_dev(DEV, name, Pmax, Tmax, forced, B=B, max_new=max_new)


def test_shelf_teacher_forced_300_frames_identical_tracking(cuda):
    """All 300 Shelf frames, track table taken from the reference before every frame: X_bin, ALS iteration
    count, matches, births/deaths, track ids and lifecycle counters all identical to the reference."""
    st = _run("shelf", 8, 24, forced=True)
    dj = np.array(st["dj"])
    print(f"shelf forced: {st}"[:200], f"joints vs reference: median {np.median(dj)*1e3:.2f} mm p90 {np.percentile(dj,90)*1e3:.2f} mm")
    assert st["xbin"] == st["alive"] == st["upd"] == st["frames"] == 300
    assert st["iters"] == 300
    assert np.median(dj) <= 7e-3      # the reference against itself + 1 ulp on these solves: 4.2 mm (test_gpu_stages)


@pytest.mark.parametrize("name,Pmax,Tmax", [("synth_c4p3", 4, 8), ("synth_c8p6", 8, 12), ("synth_c8p12", 12, 16)])
def test_synthetic_teacher_forced(cuda, name, Pmax, Tmax):
    st = _run(name, Pmax, Tmax, forced=True, max_new=2 * Pmax)
    print(f"PARITY {name} vs reference, teacher-forced incl. the no-track frame 1: frames {st['frames']} X_bin {st['xbin']} "
          f"ALS-iterations {st['iters']} track-ids {st['alive']}; non-converged no-track frames (frame, X_bin identical): {st['unstable_frames']}")
    assert st["xbin"] == st["iters"] == st["alive"] == st["upd"] == st["frames"]
    # free-running from frame 1 (births from the no-track frame feed the next associations): the reference's track ids
    fr = _run(name, Pmax, Tmax, forced=False, max_new=2 * Pmax)
    print(f"PARITY {name} free-running from frame 1: frames {fr['frames']} X_bin {fr['xbin']} track-ids {fr['alive']} "
          f"alive-count {fr['n_alive']} first mismatch {fr['first_mismatch'] if fr['first_mismatch'] < 10**9 else None}")
    assert fr["alive"] == fr["frames"], "track ids of a free-running synthetic clip differ from the reference's"


@pytest.mark.parametrize("name,Pmax,Tmax", WARM)
def test_warm_goldens_teacher_forced(cuda, name, Pmax, Tmax):
    """Tracked frames at the BASELINE shapes (8 x 32: 6 frames, 8 x 16: 8, 8 x 12: 7, Shelf-shaped 5 x 4: 11) recorded from
    the REAL reference warm-started from ground truth (oracle/make_golden.py warm): X_bin, the ALS stopping iteration,
    matched (view, pose) sets, track ids and lifecycle counters identical on every frame; IK inside the envelope."""
    st = _run(name, Pmax, Tmax, forced=True, max_new=Pmax)
    dj = np.array(st["dj"])
    print(f"PARITY {name} vs reference, teacher-forced: frames {st['frames']} X_bin {st['xbin']} ALS-iterations {st['iters']} "
          f"track-ids {st['alive']} updated-sets {st['upd']}; IK joints median {np.median(dj)*1e3:.2f} mm p90 "
          f"{np.percentile(dj, 90)*1e3:.2f} mm max {dj.max()*1e3:.2f} mm over {len(dj)} solves")
    assert st["unstable_frames"] == []
    assert st["xbin"] == st["iters"] == st["alive"] == st["upd"] == st["frames"]
    assert np.median(dj) <= 7e-3


@pytest.mark.parametrize("views,people,clips,frames,Tmax,shelf", [(8, 32, 3, 5, 40, False), (8, 16, 3, 6, 24, False),
                                                                 (5, 4, 4, 12, 8, True)])
def test_benchmarked_shapes_side_by_side_with_oracle(cuda, views, people, clips, frames, Tmax, shelf):
    """bench.py's own workload with the oracle beside it: distinct clips seeded like bench.py's CPU arm,
    teacher-forced every frame; everything discrete bit-exact (asserted inside), IK reported against the envelope."""
    from pipeline_checks import run_side_by_side_with_oracle
    st = run_side_by_side_with_oracle(DEV, views, people, clips, frames, seed=1000, Tmax=Tmax, shelf=shelf)
    dj, dp = np.array(st["dj"]), np.array(st["dparam"])
    print(f"PARITY {views}x{people} vs oracle, {clips} clips x {frames} tracked frames = {st['frames']} clip-frames: n "
          f"{min(st['n'])}..{max(st['n'])}, ALS iterations {min(st['iters'])}..{max(st['iters'])} all identical, X_bin identical, "
          f"ids/lifecycle/matches identical; max |dst| diff {st['dst']:.2e} px, max |sim| diff {st['sim']:.2e}; IK: "
          f"{st['solves']} solves, nfev identical on {st['nfev_same']}, joints median {np.median(dj)*1e3:.2f} mm p90 "
          f"{np.percentile(dj, 90)*1e3:.2f} mm, params median {np.median(dp):.3f} p90 {np.percentile(dp, 90):.3f} (rad|m), "
          f"relative cost diff median {np.median(st['dcost']):.2e}")
    assert st["frames"] == clips * frames
    assert np.median(dj) <= 7e-3
    assert st["nfev_same"] >= 0.9 * st["solves"]


def test_shelf_free_running(cuda):
    """Free-running (our own IK output feeds the next association). The reference's IK is chaotic at the 1-ulp level
    (SURVEY.md 8c'), so identical tracking cannot be guaranteed: the oracle itself, re-run with projection matrices
    perturbed by 1 ulp, reproduces X_bin on 289/300 Shelf frames, the alive-id list on 193/300 (one early/late death
    renumbers every later track) and the number of alive tracks on 297/300. Report our rates; require determinism,
    the frames before the first bifurcation, and the count-level agreement."""
    st = _run("shelf", 8, 24, forced=False, B=3)
    print("shelf free-running identical-frame rates:", {k: v for k, v in st.items() if k != "dj"})
    assert st["replicas"] == st["frames"]          # bit-deterministic across CTAs / replicas
    assert st["first_mismatch"] > 40               # the oracle's free-running golden test covers frames 1..40
    assert st["xbin"] >= 0.6 * st["frames"]
    assert st["n_alive"] >= 0.8 * st["frames"]


def test_birth_joints_match_reference(cuda):
    """Frame 1 of every golden scene: tracks are born from triangulation + 50-nfev IK; well inside 1 cm of the reference
    (birth solves converge: status 2), ids in creation order."""
    from multiview_motion_capture_b200.clips import ClipBatch
    for name, Pmax, Tmax in [("shelf", 8, 24), ("synth_c4p3", 4, 8), ("synth_c8p6", 8, 12)]:
        inp, g = golden(name)
        kps = o.body25_to_coco(inp["kps25"])
        C = kps.shape[1]
        cb = ClipBatch(1, C, Pmax, max_tracks=Tmax, max_new=Pmax, device=DEV)
        cb.set_calib(inp["K"][None], inp["RT"][None])
        rec = cb.step(pad_poses(kps[1], Pmax)[None], inp["n_pose"][1][None], 1)[0].copy()
        k = fkey(1)
        n = int(rec["n_alive"])
        assert rec["tracks"]["track_id"][:n].tolist() == g[k + "alive_after"].tolist()
        dj = np.abs(rec["tracks"]["joints"][:n].reshape(-1, 18, 3) - g[k + "upd_joints"]).max(axis=(1, 2))
        print(name, "birth joints vs reference (m):", dj)
        assert np.median(dj) <= 1e-2
        cb.close()


def test_full_size_properties_8x32(cuda):
    """BASELINE shape (8 cameras x 32 people), 6 clips x 6 frames, no oracle (it needs ~1 min/frame there):
    replicated clips give bit-identical records; the assignment groups are pure w.r.t. the generator's ground-truth
    person ids; track ids are handed out in creation order; FK(params) == joints; every IK solve lowers its cost."""
    import torch
    from multiview_motion_capture_b200 import stages as S, synthetic as syn
    from multiview_motion_capture_b200.clips import ClipBatch
    nF = 8
    clips = [syn.make_clip(8, 32, nF + 1, seed=77, clip_idx=i) for i in range(3)]
    B = 6
    idx = [0, 1, 2, 0, 1, 2]
    kps = np.stack([syn.body25_to_coco(clips[i]["kps25"]) for i in idx], 1)
    n_pose = np.stack([clips[i]["n_pose"] for i in idx], 1)
    cb = ClipBatch(B, 8, 32, max_tracks=64, max_new=64, device=DEV)
    cb.set_calib(np.stack([clips[i]["K"] for i in idx]), np.stack([clips[i]["RT"] for i in idx]))
    seen_ids = [set() for _ in range(B)]
    for f in range(1, nF + 1):
        recs = cb.step(kps[f], n_pose[f], f).copy()
        assert (recs["error"] == 0).all()
        for b in range(3):
            assert recs[b].tobytes() == recs[b + 3].tobytes(), (f, b, "replica differs")
        for b in range(3):
            rec = recs[b]
            n = int(rec["n_alive"])
            tr = rec["tracks"][:n]
            ids = tr["track_id"].tolist()
            assert ids == sorted(ids) and len(set(ids)) == n
            new = [i for i, u in zip(ids, tr["updated"]) if u == 2]
            assert all(i > max(seen_ids[b], default=-1) for i in new)
            seen_ids[b].update(ids)
            gt = clips[b]["gt_person"][f]
            pure = 0
            for t in tr[tr["updated"] > 0]:
                who = {int(gt[v, p]) for v, p in t["sel"][:t["n_sel"]]}
                pure += int(len(who) == 1)
                assert t["n_sel"] >= 2
            # frame 1 has no tracks: the reference's float32 affinity merges dozens of poses of different people into
            # giant groups there (cut to MVMC_MAX_SEL poses, n_truncated); from the first tracked frames on groups are clean
            if f >= 4:
                assert rec["n_truncated"] == 0
                assert 24 <= n <= 44, (f, b, n)
                assert pure >= 0.9 * (tr["updated"] > 0).sum(), (f, b, pure)
            upd = tr[tr["updated"] > 0]
            j = S.fk(torch.as_tensor(upd["param"].copy(), device=DEV)).cpu().numpy()
            assert np.abs(j.reshape(len(upd), 54) - upd["joints"]).max() <= 1e-12
            assert (upd["cost"][:, 1] <= upd["cost"][:, 0] * (1 + 1e-12)).all()
            assert np.isfinite(upd["param"]).all()
            # joints close to the generator's ground truth (noise 2 px at ~6 m => a few cm)
            if f >= 5:
                gtj = clips[b]["gt_joints"][f]
                err = [np.abs(t["joints"].reshape(18, 3)[1:15] - gtj[int(gt[t["sel"][0, 0], t["sel"][0, 1]])][1:15]).max()
                       for t in upd]
                assert np.median(err) < 0.15, (f, b, np.median(err))
    cb.close()


@pytest.mark.gpu
def test_clip_streams_equal_clip_batch(cuda):
    """Groups of clips on their own CUDA streams (ClipStreams, mvmc_clips_step_host_async): byte-identical records."""
    from pipeline_checks import check_clip_streams_equal_clip_batch
    check_clip_streams_equal_clip_batch(DEV, name="synth_c8p6", Pmax=8, Tmax=16, max_new=8, frames=(2, 3, 4), B=5, groups=3)


def test_edge_cases_against_the_oracle(cuda):
    """Empty / single-view / ragged / all-filtered frames and the return to the no-track path after every track died."""
    from pipeline_checks import check_edge_cases
    print("edge cases (kind, alive after, died):", check_edge_cases(DEV))


def test_crowded_no_track_frame_births_from_all_poses(cuda):
    """8 x 32, frame 1 (a group of 76 poses among others): groups of more than MVMC_MAX_SEL poses are born from all their poses,
    like the reference."""
    from pipeline_checks import check_crowded_no_track_frame
    r = check_crowded_no_track_frame(DEV, 8, 32, seed=1000, clip=1)
    print(f"PARITY crowded no-track frame (8x32, n = {r['n']}, ALS {r['als_iters']} iterations): X_bin bit-exact, groups of "
          f"{r['group_sizes']} poses identical ({r['n_big']} beyond MVMC_MAX_SEL, none truncated), track ids identical; "
          f"birth joints vs oracle (m): {[round(x, 4) for x in r['joints_diff_m']]}")
    assert r["n_big"] >= 1
