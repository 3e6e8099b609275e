"""One-off long parity fuzz (not part of the suite): many envs x many steps with dense clustered restarts, CUDA vs oracle,
Philox mode, all six parity configurations.  Prints the number of env-steps compared per configuration."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import lockstep as ls  # noqa: E402
from homophily_marl_b200.batch_env import SSDBatchEnv  # noqa: E402
from oracle import oracle as O  # noqa: E402

B, T = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1500
total = 0
for key in ("cleanup5", "cleanup10_full", "cleanup3", "harvest5", "harvest10_full", "harvest10_n5"):
    name, mp, n, view, color = ls.CONFIGS[key]
    extra = dict(obs_color=color, random_spawn_point=True, random_spawn_rotation=None)
    env = SSDBatchEnv(name, B, n, map=mp, view_size=view, episode_limit=37, extra_args=extra, seed=99, env_gid_base=12345)
    ora = O.OracleBatch.from_spec(env.spec, n_envs=B, seed=99, env_gid0=12345, random_spawn_point=True, spawn_rotation=None)
    env.reset(); ora.reset(threads=16)
    rs = np.random.RandomState(7)
    t0 = time.time()
    for t in range(T):
        if t % 5 == 2:                                   # clustered restarts: collisions, swaps, chains, cycles, duplicates
            pos, orient = ls.cluster_states(rs, env.spec, B)
            env.load_state(pos_rc=pos, orient=orient)
            for b in range(B):
                ora.set_state(b, pos_rc=pos[b], orient=orient[b])
        act = rs.randint(0, env.n_actions, size=(B, n)).astype(np.uint8)
        if t % 3 == 0:
            act[:, : n // 2 + 1] = rs.randint(0, 4, size=(B, n // 2 + 1))
        env.step(torch.as_tensor(act, device=env.device))
        out = ora.step(act, threads=16)
        assert np.array_equal(env.reward.cpu().numpy(), out["reward"]), (key, t, "reward")
        assert np.array_equal(env.clean.cpu().numpy(), out["clean"]), (key, t, "clean")
        assert np.array_equal(env.obs_view().cpu().numpy(), out["obs"]), (key, t, "obs")
        assert np.array_equal(env.apple_cnt.cpu().numpy().view(np.uint16), out["apple_cnt"]), (key, t, "apple_cnt")
        if t % 10 == 0:
            assert np.array_equal(env.grid.cpu().numpy(), ora.grid), (key, t, "grid")
            assert np.array_equal(env.agent_pos.cpu().numpy(), ora.pos_rc), (key, t, "pos")
            g = env.grid_buf[:, :env.G]                     # the running cell counts equal a recount of the grid
            assert torch.equal(env.counts_buf, ((g == 2).sum(1) | ((g == 3).sum(1) << 16)).to(torch.int32)), (key, t, "counts")
        if out["done"].all():
            env.reset(); ora.reset(threads=16)
    assert int(ora.envs["error"].sum()) == 0
    total += B * T
    print(f"{key}: {B * T} env-steps bit-exact ({time.time() - t0:.0f} s)", flush=True)
print("TOTAL env-steps compared:", total)
