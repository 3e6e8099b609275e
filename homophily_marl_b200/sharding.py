"""Env-instance sharding across GPUs: contiguous blocks, keyed by GLOBAL env id.

Env instances are independent, so the step path needs no collective; a rank's
Philox draws depend only on (seed, global env id, tick), which makes trajectories
identical for any GPU count.
"""
from __future__ import annotations


def shard_range(global_envs: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the global env ids owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(global_envs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def weak_scaling_shard(envs_per_gpu: int, rank: int) -> tuple[int, int]:
    """bench.py's weak-scaling layout: every rank owns `envs_per_gpu` envs."""
    return rank * envs_per_gpu, (rank + 1) * envs_per_gpu
