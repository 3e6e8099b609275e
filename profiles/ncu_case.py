"""Workload for one ncu capture:  python profiles/ncu_case.py <harvest5|cleanup5|cleanup10|cleanup3|frontend7|frontend15> <envs> [reset|render]
Runs resets + ~40 plain launches of the chosen kernel (no graphs), so that `ncu -k regex:<kernel> -s 30 -c 1` lands on a warm one."""
import sys

import torch

sys.path.insert(0, ".")
from homophily_marl_b200.batch_env import SSDBatchEnv  # noqa: E402

CASES = {"harvest5": ("harvest", "default5", 5, 15), "cleanup5": ("cleanup", "default5", 5, 7),
         "cleanup10": ("cleanup", "default10", 10, 7), "cleanup3": ("cleanup", "default3", 3, 7)}
what, B = sys.argv[1], int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "step"
if what.startswith("frontend"):
    from homophily_marl_b200.frontend import ObsFrontEnd
    view = int(what[len("frontend"):])
    name, mp, n = ("cleanup", "default5", 5) if view == 7 else ("harvest", "default5", 5)
    env = SSDBatchEnv(name, B, n, map=mp, view_size=view, seed=1)
    env.reset()
    P = 2 * view - 1
    torch.manual_seed(0)
    mod = torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, 1), torch.nn.LeakyReLU(), torch.nn.Flatten(),
                              torch.nn.Linear(6 * P * P, 32), torch.nn.LeakyReLU())
    fe = ObsFrontEnd.from_module(mod, view, device=env.device)
    for i in range(8):
        fe.forward_env(env)
    torch.cuda.synchronize()
    sys.exit(0)
name, mp, n, view = CASES[what]
env = SSDBatchEnv(name, B, n, map=mp, view_size=view, episode_limit=1000, seed=1, want_state=True)
env.reset()
acts = torch.randint(0, env.n_actions, (64, B, n), device=env.device, dtype=torch.int32).to(torch.uint8)
ring = [env.new_obs_buffer() for _ in range(max(2, int(280e6 / (B * env.layout.obs_env_stride)) + 1))]
for i in range(40):
    if mode == "reset":
        env.reset(obs_out=ring[i % len(ring)])
    elif mode == "render":
        env.render(obs_out=ring[i % len(ring)], want_state=True)
    else:
        env.step(acts[i % 64], obs_out=ring[i % len(ring)])
torch.cuda.synchronize()
