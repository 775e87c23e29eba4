"""CPU, world_size 2 and 3 over gloo: the frame-batch partition and the statistics all-gather (the only collective of the
path) reproduce the single-process table in frame order, including ragged shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdpl_slam_b200 import shard, synth


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard.shard_range(4096, 3, 8) == (1536, 2048)
    assert shard.shard_with_halo(8, 1, 2) == (3, 4, 8) and shard.shard_with_halo(8, 0, 2) == (0, 0, 4)
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = shard.shard_range(n_frames, rank, world)
        local = np.stack([shard.frame_checksum(synth.frame(100 + f, 48, 64, n_rect=3)) for f in range(s, e)]) if e > s \
            else np.zeros((0, 4), np.int32)
        table = shard.gather_frame_stats(torch.from_numpy(local))
        # the per-step form (shard sizes from shard_range: no count exchange, asynchronous) must give the same table
        g = shard.StatsGather(n_frames, 4, torch.int32, "cpu")
        for _ in range(2):
            g.start(torch.from_numpy(local))
        assert torch.equal(g.table(), table)
        if rank == 0:
            q.put(table.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 8), (2, 7), (3, 10)])
def test_gathered_table_equals_single_process(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([shard.frame_checksum(synth.frame(100 + f, 48, 64, n_rect=3)) for f in range(n_frames)])
    np.testing.assert_array_equal(table, ref)
