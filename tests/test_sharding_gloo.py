"""CPU tier, world_size 2 over gloo: the env batch shards into independent contiguous ranges (no collective on the
step path) and the only exchange is the all-reduce of the episode-statistics vector."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from glgym.distributed import allreduce_stats, env_shard, init_from_env, max_over_ranks
    r, w, _ = init_from_env("gloo")
    lo, hi = env_shard(total, r, w)
    # per-rank "finished episode" statistics as the step kernel would accumulate them for its shard
    stats = torch.zeros(16, dtype=torch.float64)
    ids = torch.arange(lo, hi, dtype=torch.float64)
    stats[0], stats[1], stats[2] = hi - lo, ids.sum(), 5761.0 * (hi - lo)
    allreduce_stats(stats)
    t = max_over_ranks(1.0 + r)
    out[rank] = (lo, hi, stats.tolist(), t)
    dist.destroy_process_group()


def test_env_sharding_and_stats_allreduce_world2():
    total, world = 4099, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29533, total, out), nprocs=world, join=True)
    (lo0, hi0, s0, t0), (lo1, hi1, s1, t1) = out[0], out[1]
    assert (lo0, hi0, lo1, hi1) == (0, 2050, 2050, 4099)  # contiguous, disjoint, covers the batch
    assert s0 == s1  # all-reduce gives every rank the global sums
    assert s0[0] == total and s0[1] == total * (total - 1) / 2 and s0[2] == 5761.0 * total
    assert t0 == t1 == 2.0  # max over ranks


def test_env_shard_partition_properties():
    from glgym.distributed import env_shard
    for total in (1, 7, 4096, 262144, 4099):
        for world in (1, 2, 4, 8):
            bounds = [env_shard(total, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == total
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
