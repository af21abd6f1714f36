"""Host-side data-parallel logic on CPU: world_size-2 gloo processes (no GPU needed)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from news_recommendation_model_b200.dp import GradientBuckets, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # a flat gradient buffer per rank: bucket boundaries as DataParallel uses them
        n, head_begin = 1000, 640
        g = torch.arange(n, dtype=torch.float32) * (rank + 1)
        buckets = GradientBuckets()
        buckets.reduce(g, head_begin, n)      # head bucket first (overlaps the encoder backward on a GPU)
        buckets.reduce(g, 0, head_begin)
        buckets.wait()
        expect = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok_avg = torch.allclose(g, expect)
        # BatchNorm statistics: sums add up, row counts multiply
        sums = torch.full((528,), float(rank + 1), dtype=torch.float64)
        dist.all_reduce(sums)
        ok_stats = bool((sums == sum(range(1, world + 1))).all())
        # shards tile the batch without overlap
        lo, hi = shard_range(1001, rank, world)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        ok_shard = int(sum(s.item() for s in sizes)) == 1001 and abs(sizes[0].item() - sizes[-1].item()) <= 1
        if rank == 0:
            out.put((ok_avg, ok_stats, ok_shard))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_and_sharding():
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == (True, True, True)


def test_shard_range_covers_everything():
    for total in (1, 7, 1024, 8191):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b >= a
