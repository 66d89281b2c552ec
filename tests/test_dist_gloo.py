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


def _ddp_worker(rank, world, port, q):
    """DDP over gloo on CPU; the hot-path ops are swapped for the torch port in this TEST process only."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import unsamflow_b200.flow_loss as fl
    import unsamflow_b200.pwclite as pw
    from oracle import torch_port

    class PortCorr(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, a, b):
            return torch_port.cost_volume(a, b)

    pw.Correlation, pw.flow_warp, fl.flow_warp = PortCorr, torch_port.flow_warp, torch_port.flow_warp
    fl.get_occu_mask_backward = torch_port.occ_mask_backward
    from unsamflow_b200.train_step import TrainStep
    torch.set_num_threads(2)
    dist_util.init_from_env("gloo")
    ts = TrainStep(1, 64, 128, "cpu", seed=7, ddp=True, lr=0.0)  # lr 0: weights stay put, grads stay readable
    # local (non-DDP) gradient of this rank's own pair, same weights
    out = ts.model(ts.img1, ts.img2, with_bk=True)
    flows = [torch.cat([a, b], 1) for a, b in zip(out["flows_12"], out["flows_21"])]
    ts.loss_fn(flows, ts.img1, ts.img2).backward()
    probe = ts.model.flow_estimators.conv1[0].weight
    local = probe.grad.detach().clone()
    ts.model.zero_grad(set_to_none=True)
    # one DDP step (clip_grad_norm_ rescales, so read the gradient through a hook-free second backward)
    out = ts.net(ts.img1, ts.img2, with_bk=True)
    flows = [torch.cat([a, b], 1) for a, b in zip(out["flows_12"], out["flows_21"])]
    ts.loss_fn(flows, ts.img1, ts.img2).backward()
    ddp_grad = probe.grad.detach().clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean_local = sum(gathered) / world
    q.put((rank, float((ddp_grad - mean_local).abs().max()), float(mean_local.abs().max()),
           bool(torch.equal(gathered[0], gathered[1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradients_are_the_rank_average():
    """train.py:116-120 semantics: batch sharded over ranks, DDP all-reduce averages the gradients."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, err, scale, same in out:
        assert not same            # the ranks really saw different pairs
        assert err <= 1e-5 * max(scale, 1e-12) + 1e-9, (rank, err, scale)
