"""Small batches at full speed: the B envs are stepped as G independent ranges, each on its own CUDA stream
(``SSDBatchEnv.step_range`` -> ``ssd_step_range``); prints the step time for G ranges and for one launch per step.

A launch of a few thousand envs is bound by the latency of its own chain (launch -> state loads -> logic -> observation
stores).  Independent ranges overlap one range's chain with another's stores, and consecutive env launches of a stream are
chained by programmatic dependent launch.  The trajectories do not depend on the split: random
draws are keyed by the global env id.  Actions come from a table here; a policy adds its own kernels between the env launches
of a range (``BatchedEpisodeRunner(env_groups=G)`` does that) and only pays off per range if it is a few fused kernels.

    python examples/ranges_rollout.py [n_envs] [ranges] [steps]
"""
import sys

import torch

sys.path.insert(0, ".")
from homophily_marl_b200 import SSDBatchEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
assert B % G == 0, "the ranges must divide the batch"
K = 50                                                                 # steps per CUDA graph (launching from Python is slower than the step)


def rollout(groups):
    env = SSDBatchEnv("harvest", B, num_agents=5, map="default5", view_size=15, episode_limit=steps + 4 * K, seed=0)
    dev, Bg = env.device, B // groups
    env.reset()
    torch.manual_seed(0)                                               # same actions for every split: the returns must agree
    acts = torch.randint(0, env.n_actions, (K, B, env.n), device=dev, dtype=torch.int32).to(torch.uint8)
    streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]

    def schedule(k):
        """k steps of every range: fork from the current stream, each range on its own stream, join."""
        main = torch.cuda.current_stream(dev)
        for r, s in enumerate(streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                for i in range(k):
                    env.step_range(acts[i], r * Bg, Bg)                # whole-batch action tensor; the range reads its own rows
        for s in streams:
            main.wait_stream(s)

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        schedule(2)                                                    # warm-up outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        schedule(K)
    graph.replay()                                                     # the first replay uploads the graph
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, steps // K)
    t0.record()
    for _ in range(reps):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    ms, done = t0.elapsed_time(t1), reps * K
    ret = env.ep_ret.float().sum(1).mean().item()
    env.close()
    return ms / done * 1e3, B * 5 * done / ms * 1e3, ret


for groups in (G, 1):
    us, rate, ret = rollout(groups)
    print(f"{B} envs as {groups} range(s): {us:.1f} us per step, {rate:.3e} agent-steps/s; mean collective return {ret:.2f}")
