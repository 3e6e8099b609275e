"""CPU suite: the C oracle against the golden traces generated from the unmodified reference
(tests/golden/make_golden.py) and against the published Philox4x32-10 known-answer vectors."""
import os

import numpy as np
import pytest

import lockstep as ls
from oracle import oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = sorted(k[:-4] for k in os.listdir(GOLDEN_DIR) if k.endswith(".npz"))


def load_golden(key):
    g = np.load(os.path.join(GOLDEN_DIR, key + ".npz"))
    return {k: g[k] for k in g.files}


def test_golden_present():
    assert set(KEYS) == set(ls.CONFIGS), KEYS


@pytest.mark.parametrize("key", KEYS)
def test_oracle_matches_reference_golden(key):
    g = load_golden(key)
    tr = ls.run_trace(ls.OracleBackend(key, random_spawn=bool(g["random_spawn"])),
                      ls.schedule_for(key, int(g["seed"])), int(g["steps"]))
    ls.assert_traces_equal(g, tr, key)
    assert tr["reward"].shape[0] == int(g["steps"])


@pytest.mark.parametrize("key", KEYS)
def test_golden_map_is_the_reference_map(key):
    from homophily_marl_b200 import mapspec
    g = load_golden(key)
    spec = ls.spec_for(key)
    assert mapspec.map_sha(spec.rows) == str(g["ascii_sha"]) == mapspec.MAP_SHA[spec.params.map_key]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert O.philox(ctr, key).tolist() == want


def test_episode_limit_and_returns():
    be = ls.OracleBackend("cleanup3", episode_limit=7)
    sched = ls.schedule_for("cleanup3", 3, teleport_every=0)
    tr = ls.run_trace(be, sched, 7)
    assert tr["done"].tolist() == [0] * 6 + [1]
    assert np.array_equal(be.o.ep_ret[0], tr["reward"].astype(np.int32).sum(axis=0))


def test_philox_mode_is_deterministic_and_shard_invariant():
    spec = ls.spec_for("harvest5", episode_limit=50)
    a = O.OracleBatch.from_spec(spec, n_envs=8, seed=7, env_gid0=0, random_spawn_point=True, spawn_rotation=None)
    lo = O.OracleBatch.from_spec(spec, n_envs=4, seed=7, env_gid0=0, random_spawn_point=True, spawn_rotation=None)
    hi = O.OracleBatch.from_spec(spec, n_envs=4, seed=7, env_gid0=4, random_spawn_point=True, spawn_rotation=None)
    for o in (a, lo, hi):
        o.reset()
    rs = np.random.RandomState(0)
    for t in range(40):
        act = rs.randint(0, spec.n_actions, size=(8, spec.n_agents)).astype(np.uint8)
        ra, rl, rh = a.step(act, threads=2), lo.step(act[:4]), hi.step(act[4:])
        for k in ("reward", "clean", "apple_cnt", "done", "obs"):
            assert np.array_equal(ra[k], np.concatenate([rl[k], rh[k]])), (t, k)
    assert np.array_equal(a.grid, np.concatenate([lo.grid, hi.grid]))
    assert len({a.grid[b].tobytes() for b in range(8)}) > 1      # envs diverge: draws are keyed per env
