"""world_size-2 gloo tests (CPU) of the data-parallel host logic: clip sharding, the global loss
normaliser and the bucketed gradient all-reduce order."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import product


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = __import__("importlib").import_module("3dspa_code_b200.dp")
    # 1. contiguous clip shards cover the global batch exactly once
    lo, hi = dp.shard_range(10, world, rank)
    # 2. bucketed all-reduce sums every element across ranks
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    dp.bucketed_allreduce(g, bucket_elems=96)
    # 3. global normaliser = max(sum over ALL ranks, 1)
    cnt = torch.tensor([3.0 if rank == 0 else 0.0])
    denom = dp.global_denominator(cnt)
    q.put((rank, lo, hi, g.clone(), denom))
    dist.destroy_process_group()


def test_dp_host_logic_gloo():
    product()
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, g, denom in outs:
        covered += list(range(lo, hi))
        assert torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3)
        assert denom == 3.0
    assert covered == list(range(10))


def test_shard_range_edges():
    dp = __import__("importlib").import_module("3dspa_code_b200.dp")
    assert [dp.shard_range(64, 8, r) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    spans = [dp.shard_range(5, 4, r) for r in range(4)]
    assert spans == [(0, 2), (2, 3), (3, 4), (4, 5)]
    assert dp.shard_range(1, 2, 1) == (1, 1)  # empty shard is allowed
