"""CPU: the N > 1 host-side path (stream sharding, max-over-ranks time, final host gather) with the gloo backend,
world size 2 -- no GPU involved; the per-rank engine is replaced by a deterministic per-stream function."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from irbaboon_b200 import sharding


def test_stream_ranges_partition_the_streams():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 8, 1024, 8192, 10001):
            r = [sharding.stream_range(g, world, n) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
            for tile in (2, 4, 8):
                a = [sharding.aligned_stream_range(g, world, n, tile) for g in range(world)]
                assert a[0][0] == 0 and a[-1][1] == n and all(a[i][1] == a[i + 1][0] for i in range(world - 1))
                assert all(b % tile == 0 for b, _ in a)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_engine(x):
    """stands in for one rank's engine: any per-stream (row-wise) function"""
    return np.cumsum(x, axis=1).astype(np.float32) * np.float32(0.5)


def _worker(rank, world, port, n_streams, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        x = rng.standard_normal((n_streams, B)).astype(np.float32)       # every rank can regenerate the global input
        mine = sharding.shard_streams(x, rank, world)
        b, e = sharding.stream_range(rank, world, n_streams)
        assert mine.shape[0] == e - b
        y = _fake_engine(mine)
        t = sharding.max_over_ranks(1.0 + rank)
        full = sharding.gather_streams(y, n_streams)
        dist.barrier()
        if rank == 0:
            q.put((t, np.array_equal(full, _fake_engine(x)), full.shape))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [8, 11])
def test_two_rank_gloo_shard_and_gather(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, same, shape = q.get()
    assert t == 2.0 and same and shape == (n_streams, 16)
