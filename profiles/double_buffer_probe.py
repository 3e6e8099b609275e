"""Does running the 4096 envs as G independent groups on G streams (double-buffered rollout) hide the logic phase of one
group under the observation stores of the other?  Same envs (global ids), same total work; CUDA-graph replay per group."""
import json
import sys

import torch

sys.path.insert(0, ".")
from homophily_marl_b200.batch_env import SSDBatchEnv  # noqa: E402

B, n, T = 4096, 5, 400
out = {}
for G in (1, 2, 4, 8, 16):
    for offset in (False,):
        if G == 1 and offset:
            continue
        Bg = B // G
        envs = [SSDBatchEnv("harvest", Bg, n, map="default5", view_size=15, episode_limit=100000, seed=1, env_gid_base=g * Bg) for g in range(G)]
        streams = [torch.cuda.Stream() for _ in range(G)]
        acts = [torch.randint(0, 8, (64, Bg, n), device="cuda", dtype=torch.int32).to(torch.uint8) for _ in range(G)]
        rings = [[e.new_obs_buffer() for _ in range(6)] for e in envs]
        graphs = []
        for g in range(G):
            e, s = envs[g], streams[g]
            with torch.cuda.stream(s):
                e.reset()
                for i in range(5):
                    e.step(acts[g][i], obs_out=rings[g][i % 6])
                s.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    for i in range(T):
                        e.step(acts[g][i % 64], obs_out=rings[g][i % 6])
                graphs.append(gr)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0 = torch.cuda.Event(enable_timing=True)
            ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
            e0.record()
            for g in range(G):
                streams[g].wait_event(e0)
                with torch.cuda.stream(streams[g]):
                    if offset and g > 0:
                        torch.cuda._sleep(int(9000 * g / G))          # ~ a fraction of a step, establishes the phase offset
                    graphs[g].replay()
                    ends[g].record(streams[g])
            torch.cuda.synchronize()
            best = min(best, max(e0.elapsed_time(x) for x in ends))
        out[f"G{G}_offset{int(offset)}"] = {"us_per_step_of_4096": best / T * 1e3, "agent_steps_per_s": B * n * T / (best * 1e-3)}
        del envs, rings, graphs
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
