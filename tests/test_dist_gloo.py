"""CPU suite, world_size 2 over gloo: the multi-GPU host logic -- contiguous env shards keyed by
global env id give the same trajectories as one process, with no collective on the step path; the only
cross-rank operations are the bench's barrier and MAX-reduction of the timed region."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lockstep as ls
    from homophily_marl_b200.sharding import shard_range
    from oracle import oracle as O
    spec = ls.spec_for("cleanup10", episode_limit=100)
    lo, hi = shard_range(B, rank, world)
    ob = O.OracleBatch.from_spec(spec, n_envs=hi - lo, seed=77, env_gid0=lo, random_spawn_point=True, spawn_rotation=None)
    ob.reset()
    rs = np.random.RandomState(5)
    rew = []
    dist.barrier()
    for t in range(T):
        act = rs.randint(0, spec.n_actions, size=(B, spec.n_agents)).astype(np.uint8)     # same global actions on every rank
        rew.append(ob.step(act[lo:hi], want_obs=False)["reward"].copy())
    # the bench's reduction: elapsed = MAX over ranks
    el = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    assert el.item() == world
    grids = [None] * world
    dist.all_gather_object(grids, (lo, hi, ob.grid.copy(), np.stack(rew)))
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"),
                 grid=np.concatenate([g[2] for g in grids]), reward=np.concatenate([g[3] for g in grids], axis=1),
                 bounds=np.array([[g[0], g[1]] for g in grids]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_oracle_equals_single_process(tmp_path):
    import lockstep as ls
    from oracle import oracle as O
    B, T, world = 10, 12, 2
    mp.spawn(_worker, args=(world, _free_port(), B, T, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "gathered.npz"))
    assert got["bounds"].tolist() == [[0, 5], [5, 10]]
    spec = ls.spec_for("cleanup10", episode_limit=100)
    whole = O.OracleBatch.from_spec(spec, n_envs=B, seed=77, env_gid0=0, random_spawn_point=True, spawn_rotation=None)
    whole.reset()
    rs = np.random.RandomState(5)
    rew = []
    for t in range(T):
        act = rs.randint(0, spec.n_actions, size=(B, spec.n_agents)).astype(np.uint8)
        rew.append(whole.step(act, want_obs=False)["reward"].copy())
    assert np.array_equal(got["grid"], whole.grid)
    assert np.array_equal(got["reward"], np.stack(rew))


def test_shard_range_partitions():
    from homophily_marl_b200.sharding import shard_range, weak_scaling_shard
    for B in (1, 7, 16384, 65536):
        for world in (1, 2, 3, 8):
            parts = [shard_range(B, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(16384, 3, 8) == (6144, 8192)                 # BASELINE configs[2]: 2048 envs per GPU
    assert weak_scaling_shard(4096, 2) == (8192, 12288)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
