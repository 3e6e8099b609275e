"""Dev-container only: the C oracle in lockstep with the UNMODIFIED reference (imported from /root/reference,
injected position-indexed draws).  Skipped wherever the reference is absent (the GPU box)."""
import os
import sys

import pytest

import lockstep as ls
from oracle import refshim

pytestmark = [pytest.mark.needs_reference,
              pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not present")]


def _ref_backend():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    return make_golden.RefBackend


@pytest.mark.parametrize("key,seed,rnd", [("cleanup5", 301, True), ("cleanup10_full", 302, False), ("cleanup3", 303, True),
                                          ("harvest5", 304, False), ("harvest10_full", 305, True)])
def test_oracle_lockstep_with_live_reference(key, seed, rnd):
    RefBackend = _ref_backend()
    kw = dict(teleport_every=6)
    ref = ls.run_trace(RefBackend(key, random_spawn=rnd), ls.schedule_for(key, seed, **kw), 90)
    ob = ls.OracleBackend(key, random_spawn=rnd)
    ora = ls.run_trace(ob, ls.schedule_for(key, seed, **kw), 90)
    ls.assert_traces_equal(ref, ora, key)
    assert int(ob.o.envs["error"][0]) == 0


def test_seeded_anchor_hashes_of_the_reference():
    """SURVEY Appendix C.3 anchor: unmodified reference, numpy/python RNG seeded with 0, 1000 steps of
    cleanup default5 -- proves the import shim runs the same reference the survey measured."""
    import hashlib
    import random
    import numpy as np
    env = refshim.make("cleanup", 5, "default5", 7)
    np.random.seed(0)
    random.seed(0)
    env.reset()
    acts = np.random.RandomState(123).randint(0, env.n_actions, size=(1000, 5))
    hA, hB = hashlib.sha256(), hashlib.sha256()
    for s in range(1000):
        r, term, info = env.step(acts[s])
        hA.update(env.world_map.astype("S1").tobytes())
        hA.update(env.get_agent_pos().astype(np.int16).tobytes())
        hA.update(r.astype(np.int8).tobytes())
        hA.update(info["clean_num"].astype(np.int8).tobytes())
        hB.update(np.uint8(np.stack(env.get_obs()) * 256).tobytes())
    assert hA.hexdigest()[:16] == "bd7111014f50657d" and hB.hexdigest()[:16] == "a45f7cb95a093744"
    assert env.rewards.tolist() == [-93.0, -114.0, -116.0, -86.0, -100.0]
    assert info["equality_metric"] == 0.9363457760314342
