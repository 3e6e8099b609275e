"""Small all-paths workload for compute-sanitizer (memcheck / racecheck): both envs, both colour modes, all modes of the
kernel, generic and specialised geometry, injected draws."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from homophily_marl_b200.batch_env import SSDBatchEnv  # noqa: E402

rs = np.random.RandomState(0)
for name, mp, n, view, color, rows in (("cleanup", "default5", 5, 7, "simplified", None), ("harvest", "default5", 5, 15, "full", None),
                                       ("cleanup", "default10", 10, 7, "full", None),
                                       ("harvest", "x", 3, 4, "simplified", ["@@@@@@@", "@P A P@", "@ AAA @", "@P   A@", "@@@@@@@"])):
    B = 12
    env = SSDBatchEnv(name, B, n, map=mp, view_size=view, episode_limit=6, rows=rows, seed=3, want_state=True,
                      extra_args=dict(random_spawn_point=True, random_spawn_rotation=None, obs_color=color))
    env.reset()
    for t in range(8):
        act = torch.as_tensor(rs.randint(0, env.n_actions, size=(B, n)).astype(np.uint8), device=env.device)
        draws = None
        if t == 3:
            draws = dict(prio=rs.randint(0, 2 ** 32, size=(B, n), dtype=np.uint64), u_apple=rs.randint(0, 2 ** 32, size=(B, env.G), dtype=np.uint64),
                         u_waste=rs.randint(0, 2 ** 32, size=(B, env.G), dtype=np.uint64), wkey=rs.randint(0, 2 ** 32, size=(B, env.G), dtype=np.uint64))
        env.step(act, draws=draws, want_state=True)
        if bool(env.done[0].item()):
            mask = torch.zeros(B, dtype=torch.uint8)
            mask[::2] = 1
            env.reset(mask=mask)
    env.render(want_obs=True, want_state=True)
    torch.cuda.synchronize()
    env.close()
print("sanitizer case done")
