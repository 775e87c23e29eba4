"""Frame-batch sharding across the GPUs of one box (BASELINE.json configs[3]).

Frames are independent units, so the data path has no collective: rank g of G processes the contiguous block
[g*B/G, (g+1)*B/G) of a B-frame batch with its own handles, streams and device arenas.  The only exchange is one
all-gather of the per-frame statistics {n_kp, n_lines, n_pt_matches, n_ln_matches} (16 B per frame) at the end of a
step: NCCL over NVLink on the GPU box, gloo in the CPU tests.  Frame-to-frame matching across a shard edge uses the
neighbour rank's last frame, which the owner of the edge simply recomputes (one extra frame per rank, no 64 KB send)."""
import numpy as np


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous block of rank `rank`: the first n_frames % world ranks get one frame more."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_with_halo(n_frames: int, rank: int, world: int):
    """(first frame to process, first frame owned, end): ranks > 0 also process the frame before their block so that the
    first owned frame can be matched against its predecessor without any transfer."""
    s, e = shard_range(n_frames, rank, world)
    return (max(s - 1, 0) if s < e else s), s, e


def gather_frame_stats(stats, group=None):
    """All-gather of per-rank (n_local, 4) int32 statistics into the (n_total, 4) table of the whole batch, in frame order.
    `stats` is a torch tensor on the device the process group communicates on (CUDA for NCCL, CPU for gloo).  Ragged shards
    are padded to the longest shard for the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return stats
    world = dist.get_world_size(group)
    n_local = torch.tensor([stats.shape[0]], dtype=torch.int64, device=stats.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    padded = torch.zeros((m, stats.shape[1]), dtype=stats.dtype, device=stats.device)
    padded[:stats.shape[0]] = stats
    out = torch.zeros((world * m, stats.shape[1]), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)], 0)


class StatsGather:
    """The per-step form of gather_frame_stats for a batch whose size is known to every rank (the bench's and a production
    loop's case): the shard sizes follow from shard_range, so there is no count exchange and no host synchronisation -- one
    asynchronous all_gather_into_tensor per step into a buffer allocated once; table() waits for the last one and returns the
    (n_total, cols) table in frame order.  Ragged shards are padded to the longest one."""

    def __init__(self, n_total, cols, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.counts = [e - s for s, e in (shard_range(n_total, r, self.world) for r in range(self.world))]
        self.m = max(self.counts) if self.counts else 0
        self.ragged = any(c != self.m for c in self.counts)
        self.out = torch.zeros((self.world * self.m, cols), dtype=dtype, device=device)
        self.pad = torch.zeros((self.m, cols), dtype=dtype, device=device) if self.ragged else None
        self.work = None

    def start(self, stats):
        """Enqueue the gather of this rank's (n_local, cols) table; returns at once."""
        import torch.distributed as dist
        if stats.shape[0] != self.counts[self.rank]:
            raise ValueError("rank %d holds %d rows, its shard has %d" % (self.rank, stats.shape[0], self.counts[self.rank]))
        if not self.on:
            self.out[:stats.shape[0]].copy_(stats)
            return
        if self.work is not None:
            self.work.wait()                # the previous gather of this object has read `stats` / `pad` before they are rewritten
        src = stats
        if self.ragged:
            self.pad[:stats.shape[0]].copy_(stats)
            src = self.pad
        self.work = dist.all_gather_into_tensor(self.out, src.contiguous(), group=self.group, async_op=True)

    def table(self):
        import torch
        if self.work is not None:
            self.work.wait()
            self.work = None
        if not self.ragged:
            return self.out
        return torch.cat([self.out[r * self.m:r * self.m + self.counts[r]] for r in range(self.world)], 0)


def frame_checksum(img: np.ndarray) -> np.ndarray:
    """Cheap deterministic 4-int summary of a frame; stands in for the GPU statistics in the CPU (gloo) tests."""
    a = img.astype(np.int64)
    return np.array([a.sum() % 65521, (a[::7, ::5] * 3).sum() % 65521, int(a.max()), int((a > 128).sum() % 65521)], np.int32)
