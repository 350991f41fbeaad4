"""Frame sharding across GPUs: frames are independent units, so rank g owns the contiguous range
[g*T/G, (g+1)*T/G) and additionally ingests the one-frame halo `start-1` as ICP target.  No collective is on the
data path; the only exchange is a host gather of the per-frame normal equations."""
import numpy as np


def frame_range(total_frames, world, rank):
    """(first frame, number of frames) owned by `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(total_frames, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def halo_range(total_frames, world, rank):
    """(first scan to ingest, number of scans): the owned frames plus the previous scan."""
    start, count = frame_range(total_frames, world, rank)
    return start - 1, count + 1


def gather_rows(dist, local, dst=0, group=None):
    """Gather [count_r, ...] arrays of every rank to `dst` (torch.distributed, any backend) and concatenate in
    rank order.  Returns the concatenation on dst, None elsewhere."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local))
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64), group=group)
    mx = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype)
    pad[: t.shape[0]] = t
    out = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst, group=group)
    if rank != dst:
        return None
    return np.concatenate([o[: int(c.item())].numpy() for o, c in zip(out, counts)], axis=0)


class RowGather:
    """The same gather, asynchronous and with the shapes fixed once: `start(local)` posts the exchange and returns at once (so the
    next batch can be issued to the GPU while the rows travel), `wait()` completes it and returns the concatenation on dst (None
    elsewhere).  `local` must stay untouched between the two calls.  Row counts are exchanged once, at construction."""

    def __init__(self, dist, rows, shape_tail, dtype, dst=0, group=None):
        import torch
        self.dist, self.dst, self.group = dist, dst, group
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([rows], dtype=torch.int64), group=group)
        self.counts = [int(c.item()) for c in counts]
        mx = max(self.counts)
        tdt = torch.from_numpy(np.zeros(1, dtype)).dtype
        self.pad = torch.zeros((mx,) + tuple(shape_tail), dtype=tdt)
        self.out = [torch.zeros_like(self.pad) for _ in range(self.world)] if self.rank == dst else None
        self.work = None

    def start(self, local):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(local))
        self.pad[: t.shape[0]] = t
        self.work = self.dist.gather(self.pad, self.out, dst=self.dst, group=self.group, async_op=True)

    def wait(self):
        if self.work is None:
            return None
        self.work.wait()
        self.work = None
        if self.rank != self.dst:
            return None
        return np.concatenate([o[:c].numpy() for o, c in zip(self.out, self.counts)], axis=0)
