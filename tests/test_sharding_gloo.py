"""Multi-process path on CPU: world_size 2 over gloo. Each rank tracks its shard of replicated/independent clips through
the clip pipeline (kernel emulator), results are all_gathered and must equal a single-process run of all clips."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import EMU_LIB, ROOT, GoldenTable, golden, pad_poses

from multiview_motion_capture_b200 import sharding


def test_shard_clips_partition():
    for n, w in [(4096, 8), (7, 2), (5, 8), (0, 3)]:
        parts = [sharding.shard_clips(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert all(sharding.clip_owner(int(c), w) == r for r, p in enumerate(parts) for c in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_clips(4, 2, 2)


def _inputs(n_clips):
    """n_clips clips = the c4p3 synthetic golden scene at a frame that depends on the clip index (so clips differ)."""
    import mvmc_oracle as o
    inp, g = golden("synth_c4p3")
    kps = o.body25_to_coco(inp["kps25"])
    frames = [2 + (c % 3) for c in range(n_clips)]
    return inp, g, kps, frames


def _track(clip_ids):
    """One tracking frame for the given global clip ids (the track table of the frame taken from the golden run: the
    emulator then needs tens, not a thousand, matcher iterations per clip); returns their records."""
    from multiview_motion_capture_b200 import _lib
    from multiview_motion_capture_b200.clips import ClipBatch
    _lib.use_library(EMU_LIB, device="cpu")
    inp, g, kps, frames = _inputs(max(clip_ids) + 1)
    B = len(clip_ids)
    cb = ClipBatch(B, 4, 4, max_tracks=8, max_new=4, device="cpu")
    cb.set_calib(np.repeat(inp["K"][None], B, 0), np.repeat(inp["RT"][None], B, 0))
    packs = [GoldenTable(g).packed(frames[c], 1, 8) for c in clip_ids]
    cb.set_tracks(**{k: np.concatenate([p[k] for p in packs]) for k in packs[0]})
    k = np.stack([pad_poses(kps[frames[c]], 4) for c in clip_ids])
    n = np.stack([inp["n_pose"][frames[c]] for c in clip_ids])
    rec = cb.step(k, n, 1).copy()
    packed, count = cb.pack_records(8, clip0=0)
    packed, count = packed.clone(), count.clone()
    cb.close()
    return rec, packed, count


def _worker(rank, world, port, n_clips, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_clips(n_clips, rank, world)
        rec, packed, count = _track(mine.tolist())
        allrec = sharding.gather_records(rec, n_clips)
        slowest = sharding.reduce_max(float(rank + 1))
        # the compact track records, gathered on rank 0 (equal shard sizes: pad the short shard with an empty clip)
        per = -(-n_clips // world)
        if packed.shape[0] < per:
            packed = torch.cat([packed, torch.zeros((per - packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype)])
            count = torch.cat([count, torch.zeros(per - count.shape[0], dtype=count.dtype)])
        recs, cnts = sharding.gather_track_records(packed, count, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, "gathered.npy"), allrec.view(np.uint8))
            np.save(os.path.join(out_dir, "max.npy"), np.array([slowest]))
            np.save(os.path.join(out_dir, "packed.npy"), recs.numpy())
            np.save(os.path.join(out_dir, "counts.npy"), cnts.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_single_process(tmp_path, emu):
    from multiview_motion_capture_b200._lib import STEP_OUT_DTYPE
    n_clips, world = 3, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_clips, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy").view(STEP_OUT_DTYPE).reshape(n_clips)
    ref, ref_packed, ref_count = _track(list(range(n_clips)))
    assert float(np.load(tmp_path / "max.npy")[0]) == 2.0
    for c in range(n_clips):
        n = int(ref[c]["n_alive"])
        assert n == int(got[c]["n_alive"]) and n >= 2
        assert got[c].tobytes() == ref[c].tobytes(), c      # bit-identical records wherever a clip is tracked
    # compact records (mvmc_clips_pack_records + gather_track_records): clip c sits at [c % world, c // world]
    packed, counts = np.load(tmp_path / "packed.npy"), np.load(tmp_path / "counts.npy")
    for c in range(n_clips):
        n = int(ref_count[c])
        assert n == int(counts[c % world, c // world]) == int((ref[c]["tracks"]["updated"][:int(ref[c]["n_alive"])] > 0).sum())
        a, b = packed[c % world, c // world, :n], ref_packed[c, :n].numpy()
        assert np.array_equal(a[:, 1:], b[:, 1:])             # (column 0 is the clip id relative to each call's clip0)
        upd = ref[c]["tracks"][:int(ref[c]["n_alive"])]
        upd = upd[upd["updated"] > 0]
        assert np.array_equal(a[:, 1], upd["track_id"]) and np.array_equal(a[:, 6:74], upd["param"]) and np.array_equal(a[:, 74:], upd["joints"])


class _FakeBatch:
    """Stands in for a ClipBatch: pack_records writes a pattern that depends on (rank, step)."""
    def __init__(self, B, rank):
        self.B, self.rank, self.step = B, rank, 0

    def pack_records(self, cap, clip0=0, rec=None, count=None):
        rec.zero_()
        rec[:, 0, 5] = 2.0                                   # one solved track per clip
        rec[:, 0, 6] = 1000.0 * self.rank + self.step        # a "parameter"
        count.fill_(1)
        self.step += 1
        return rec, count


def _gather_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []
        g = sharding.RecordGatherer(_FakeBatch(3, rank), cap=2, device="cpu", consume=lambda r: seen.append(r[:, :, 0, 6].clone()))
        for _ in range(5):
            g.submit()
        g.finish()
        recs, used = g.last()
        if rank == 0:
            np.save(os.path.join(out_dir, "seen.npy"), torch.stack(seen).numpy())
            np.save(os.path.join(out_dir, "used.npy"), used.numpy())
        else:
            assert seen == []
    finally:
        dist.destroy_process_group()


def test_record_gatherer_async_double_buffered(tmp_path):
    """sharding.RecordGatherer over gloo, world size 2: every step's records of both ranks reach rank 0 exactly once, in
    order, through the two alternating buffers."""
    world = 2
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    seen = np.load(tmp_path / "seen.npy")          # [steps, world, B]
    assert seen.shape == (5, 2, 3)
    for s in range(5):
        for r in range(2):
            assert (seen[s, r] == 1000.0 * r + s).all(), (s, r, seen[s, r])
    assert (np.load(tmp_path / "used.npy") == 1).all()
