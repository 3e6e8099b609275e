"""Times the step kernel with and without the observation render, and the render-only kernel (CUDA graphs, B200)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from homophily_marl_b200.batch_env import SSDBatchEnv  # noqa: E402

out = {}
import itertools
CASES = [("harvest", "default5", 5, 15, B) for B in (148, 592, 1184, 2368, 4096, 8192, 16384, 65536)] + \
        [("cleanup", "default5", 5, 7, B) for B in (592, 4096, 65536)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] == sys.argv[1]]
for name, mp, n, view, B in CASES:
    env = SSDBatchEnv(name, B, n, map=mp, view_size=view, episode_limit=1000, seed=1)
    env.reset()
    acts = torch.randint(0, env.n_actions, (64, B, n), device=env.device, dtype=torch.int32).to(torch.uint8)
    ring = [env.new_obs_buffer() for _ in range(max(2, int(280e6 / (B * env.layout.obs_env_stride)) + 1))]
    s = torch.cuda.Stream()

    def timeit(fn, steps=400):
        with torch.cuda.stream(s):
            for i in range(5):
                fn(i)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(steps):
                    fn(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            s.synchronize()
        return e0.elapsed_time(e1) / steps * 1e3

    key = f"{name}_b{B}"
    out[key] = {
        "step+obs_us": timeit(lambda i: env.step(acts[i % 64], obs_out=ring[i % len(ring)])),
        "step_no_obs_us": timeit(lambda i: env.step(acts[i % 64], want_obs=False)),
        "render_only_us": timeit(lambda i: env.render(obs_out=ring[i % len(ring)])),
    }
    env.close()
    del ring
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
