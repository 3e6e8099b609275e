"""Per-step cost of the PyMARL facade under the reference runner's call pattern (one env, Cleanup default3 / Harvest)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from homophily_marl_b200 import REGISTRY  # noqa: E402

out = {}
for name, mp, n, view in (("cleanup", "default3", 3, 7), ("cleanup", "default5", 5, 7), ("harvest", "default10", 5, 15)):
    env = REGISTRY[name](num_agents=n, map=mp, view_size=view, episode_limit=100, quiet=True)
    rs = np.random.RandomState(0)
    env.reset()
    for rep in range(2):
        t0 = time.perf_counter()
        steps = 0
        for ep in range(3):
            env.reset()
            term = False
            while not term:
                env.get_state(); env.get_avail_actions(); env.get_obs(); env.get_agent_pos(); env.get_agent_orientation()
                r, term, info = env.step(torch.as_tensor(rs.randint(0, 5, size=(n, 1)), device='cuda'))   # like the MAC's output
                env.get_agent_pos()
                steps += 1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out[f"{name}_{mp}_n{n}"] = {"us_per_env_step": dt / steps * 1e6, "agent_steps_per_s": steps * n / dt}
print(json.dumps(out, indent=1))
