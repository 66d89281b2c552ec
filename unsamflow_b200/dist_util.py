"""Multi-GPU plumbing for the hot path: batch sharding and rank-aggregated timing.

The path shards by batch only: every correlation / warp output depends on one
sample, so ranks exchange nothing on the data path.  The reference shards the same way:
train.py:77-87 (DistributedSampler(drop_last=True), per-GPU batch = batch_size // world_size).
"""
import os

import torch
import torch.distributed as dist


def per_rank_batch(global_batch, world_size):
    """train.py:82 -- batch_size // world_size (the remainder is dropped, as the reference does)."""
    if world_size <= 0 or global_batch < world_size:
        raise ValueError(f"global batch {global_batch} cannot be split over {world_size} ranks")
    return global_batch // world_size


def shard_bounds(global_batch, world_size, rank):
    """Contiguous [lo, hi) slice of the (already shuffled) global batch owned by `rank`."""
    b = per_rank_batch(global_batch, world_size)
    return rank * b, (rank + 1) * b


def init_from_env(backend=None):
    """One process per GPU, rendezvous from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    return dist.get_rank(), dist.get_world_size()


def _reduce_device(device=None):
    """NCCL reduces CUDA tensors only: default to this rank's current GPU there, to the CPU under gloo."""
    if device is not None:
        return torch.device(device)
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (timings are reported as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=_reduce_device(device))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_throughput(units_per_rank, seconds_this_rank, device=None):
    """Whole-job throughput: units all ranks processed / slowest rank's time."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    (tmax,) = max_over_ranks([seconds_this_rank], device)
    return world * units_per_rank / tmax
