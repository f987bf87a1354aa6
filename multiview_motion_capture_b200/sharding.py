"""Multi-GPU sharding of the capture path: by clip, nothing else.

Clips share no state and the frames of one clip are sequentially dependent (track table, IK warm start), so the only
parallel axis across GPUs is the clip: rank r of G owns clips r, r+G, r+2G, ... One process per GPU
(torch.distributed: NCCL on GPUs, gloo in the CPU test tier); there is no collective inside the per-frame loop. After a
batch of frames the ranks exchange fixed-stride result records (or just their summaries) with one all_gather."""
import numpy as np
import torch
import torch.distributed as dist


def shard_clips(n_clips: int, rank: int, world: int) -> np.ndarray:
    """Global indices of the clips rank `rank` of `world` processes owns (round robin, as BASELINE config 5)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    return np.arange(rank, n_clips, world)


def clip_owner(clip: int, world: int) -> int:
    return clip % world


def gather_records(local: np.ndarray, n_clips: int, device=None):
    """all_gather of per-clip fixed-stride records. `local` [n_local, ...] holds this rank's clips in shard order; returns
    [n_clips, ...] in global clip order on every rank. Works without an initialised process group (world of one)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local.copy()
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-n_clips // world)
    item = local.dtype.itemsize * int(np.prod(local.shape[1:], dtype=np.int64))
    buf = np.zeros((per, item), dtype=np.uint8)
    buf[:len(local)] = np.ascontiguousarray(local).view(np.uint8).reshape(len(local), item)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros((n_clips,) + local.shape[1:], dtype=local.dtype)
    for r, p in enumerate(parts):
        idx = shard_clips(n_clips, r, world)
        rows = p.cpu().numpy()[:len(idx)]
        out[idx] = rows.reshape(-1).view(local.dtype).reshape((len(idx),) + local.shape[1:])
    return out


def reduce_max(value: float, device=None) -> float:
    """max over ranks (device timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_track_records(rec: torch.Tensor, count: torch.Tensor, dst: int = 0):
    """The path's one collective: every rank's compact track records of a step (ClipBatch.pack_records: rec [B,cap,128]
    f64, count [B] i32, device tensors) gathered on rank `dst` over NCCL (gloo in the CPU test tier). Ranks must hold the
    same B (pad the last shard). Returns (recs [world,B,cap,128], counts [world,B]) on `dst`, (None, None) elsewhere;
    without a process group the inputs come back with a leading axis of one."""
    if not (dist.is_available() and dist.is_initialized()):
        return rec[None], count[None]
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        recs = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
        cnts = torch.empty((world,) + tuple(count.shape), dtype=count.dtype, device=count.device)
        dist.gather(rec, list(recs.unbind(0)), dst=dst)
        dist.gather(count, list(cnts.unbind(0)), dst=dst)
        return recs, cnts
    dist.gather(rec, None, dst=dst)
    dist.gather(count, None, dst=dst)
    return None, None


class RecordGatherer:
    """Per-step gather of a clip batch's track records on rank `dst`, off the batch's critical path.

    After a step the batch packs the tracks it solved into 1 KB records (ClipBatch.pack_records) and the records of all
    ranks are gathered on `dst` with an ASYNCHRONOUS torch.distributed.gather: NCCL's stream waits for the pack, but the
    batch's own stream does not wait for the collective, so the next step starts at once (a synchronous gather makes every
    clip group wait, every step, for the slowest rank's same group: measured 510 instead of 376 ms per step on two GPUs).
    Two record buffers alternate; before a buffer is packed again its stream waits for the gather that last read it.
    A record row with column 5 (views used) == 0 is padding: counts need no collective of their own."""

    def __init__(self, batch, cap, device, dst=0, consume=None):
        """consume(records [world,B,cap,128]) is called on `dst`, on the batch's stream, for every gathered step once its
        gather has landed (before its buffer is reused, and from finish() for the last two)."""
        self.cb, self.cap, self.dst, self.consume = batch, cap, dst, consume
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        B = batch.B
        self.rec = [torch.zeros((B, cap, 128), dtype=torch.float64, device=device) for _ in range(2)]
        self.cnt = [torch.zeros((B,), dtype=torch.int32, device=device) for _ in range(2)]
        self.out = [torch.zeros((self.world, B, cap, 128), dtype=torch.float64, device=device) if (self.world > 1 and self.rank == dst)
                    else None for _ in range(2)]
        self.work = [None, None]
        self.n = 0

    def submit(self, clip0=0):
        """Call on the batch's stream right after a step. Returns the buffer index used."""
        i = self.n & 1
        self.n += 1
        self._landed(i)
        self.cb.pack_records(self.cap, clip0=clip0, rec=self.rec[i], count=self.cnt[i])
        if self.world > 1:
            self.work[i] = dist.gather(self.rec[i], list(self.out[i].unbind(0)) if self.rank == self.dst else None, dst=self.dst,
                                       async_op=True)
        else:
            self.work[i] = True          # nothing to gather on one GPU: the packed records are the result
        return i

    def _landed(self, i):
        if self.work[i] is not None:
            if self.work[i] is not True:
                self.work[i].wait()      # (a stream-side wait: the gather that last read this buffer)
            if self.consume is not None and self.rank == self.dst:
                self.consume(self.out[i] if self.out[i] is not None else self.rec[i][None])
            self.work[i] = None

    def finish(self):
        for i in ((self.n & 1), ((self.n + 1) & 1)):      # older buffer first
            self._landed(i)

    def last(self):
        """(records [world,B,cap,128], rows used [world,B]) of the last submitted step on `dst` (after finish())."""
        i = (self.n - 1) & 1
        recs = self.out[i] if self.out[i] is not None else self.rec[i][None]
        return recs, (recs[..., 5] > 0).sum(-1)
