"""Multi-GPU plumbing: the env batch shards over ranks as contiguous, independent index ranges --
there is NO collective on the step path.  The only exchange is an optional all-reduce (sum) of the 16-entry
finished-episode statistics vector, once per logging interval.  One process per GPU (torchrun); backend nccl on
GPUs, gloo in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def env_shard(num_envs_total, rank, world_size):
    """Contiguous shard [lo, hi) of the global env index range owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(num_envs_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* variables; returns (rank, world, local)."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, world, local


def allreduce_stats(stats):
    """Sum of the per-rank episode-statistics vectors (a [16] float64 tensor, device or host). In place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value, device=None):
    """Max over ranks of a python float (the bench's step-time reduction)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
