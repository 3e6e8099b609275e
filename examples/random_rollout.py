"""Minimal use of the batched API: B Harvest envs, uniform random actions, episode returns at the end.

    python examples/random_rollout.py [n_envs] [episodes]
"""
import sys

import torch

sys.path.insert(0, ".")
from homophily_marl_b200 import SSDBatchEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
env = SSDBatchEnv("harvest", B, num_agents=5, map="default10", view_size=15, episode_limit=100,
                  extra_args=dict(random_spawn_point=True, random_spawn_rotation=None), seed=0)
for ep in range(episodes):
    env.reset()
    done = False
    while not done:
        obs = env.obs_view()                      # u8 [B, 5, 3, 31, 31] on the GPU; obs.float() / 256 is what get_obs() returns
        actions = torch.randint(0, env.n_actions, (B, env.n), device=env.device, dtype=torch.int32).to(torch.uint8)
        env.step(actions)                         # rewards in env.reward (int8 [B, 5]), termination in env.done (u8 [B])
        done = bool(env.done[0].item())           # all envs share episode_limit
    ret = env.ep_ret.float()
    print(f"episode {ep}: mean collective return {ret.sum(1).mean().item():.2f}, apples left {(env.apple_cnt.int() & 0xFFFF).float().mean().item():.1f}")
