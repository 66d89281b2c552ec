"""N > 1 host logic on CPU: 2 ranks over gloo (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unsamflow_b200 import dist_util


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = dist_util.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = dist_util.shard_bounds(8, w, r)
    # each rank "processes" its shard; rank 1 is slower
    secs = 1.0 + r
    thr = dist_util.aggregate_throughput(hi - lo, secs)
    mx = dist_util.max_over_ranks([secs, 10.0 - r])
    # shards partition the batch: gather the bounds and check on every rank
    got = [None] * w
    dist.all_gather_object(got, (lo, hi))
    q.put((rank, thr, mx, got))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, thr, mx, got in out:
        assert thr == pytest.approx(8 / 2.0)          # 2 ranks x 4 pairs / slowest rank (2 s)
        assert mx == [2.0, 10.0]
        assert got == [(0, 4), (4, 8)]


def test_per_rank_batch_follows_reference():
    # train.py:82: batch_size // world_size; global 8 -> 8/4/2/1 at 1/2/4/8 GPUs
    assert [dist_util.per_rank_batch(8, w) for w in (1, 2, 4, 8)] == [8, 4, 2, 1]
    assert dist_util.per_rank_batch(9, 2) == 4  # drop_last
    with pytest.raises(ValueError):
        dist_util.per_rank_batch(4, 8)
    assert dist_util.max_over_ranks([1.5]) == [1.5]  # single process: identity
