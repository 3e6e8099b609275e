"""Evidence for "no host sync in the step loop" of BatchedEpisodeRunner: one 100-step rollout of B envs under torch.profiler,
CUDA runtime calls counted by name.  Synchronising calls (cudaStreamSynchronize / cudaDeviceSynchronize / cudaEventSynchronize and
device-to-host copies) may only appear a constant number of times per EPISODE (the done check and the returns at the end), not per step.

    python profiles/runner_sync_probe.py [B]     -> one JSON line
"""
import json
import os
import sys
from collections import Counter

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import refloop  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = refloop.load_config("cleanup", seed=0, use_cuda=True, save_model=False, runner="batched", batch_size_run=B, buffer_size=B,
                          buffer_cpu_only=False, fused_frontend=True, action_selector="epsilon_greedy_b200",
                          learner="homophily_learner_b200", test_nepisode=B, env_args=dict(num_agents=3, map="default3"))
c = refloop.build_components(cfg, backend="b200")
c.runner.run(test_mode=False)                                   # warm-up (lazy initialisation, weight packing)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    c.runner.run(test_mode=False)
    torch.cuda.synchronize()
calls, kernels, d2h = Counter(), Counter(), 0
for e in prof.events():
    name = e.name
    if name.startswith("cuda") and e.device_type == torch.autograd.DeviceType.CPU:
        calls[name] += 1
    if "Memcpy DtoH" in name:
        d2h += 1
    if e.device_type == torch.autograd.DeviceType.CUDA:
        for key in ("ssd_kernel", "obs_frontend_kernel", "frontend_finalize_kernel", "select_actions_kernel"):
            if key in name:
                kernels[key] += 1
steps = c.runner.episode_limit
sync = {k: v for k, v in calls.items() if "Synchronize" in k}
res = {"B": B, "steps_per_episode": steps, "sync_calls_per_episode": sync, "memcpy_dtoh_per_episode": d2h,
       "cudaLaunchKernel_per_step": round(calls.get("cudaLaunchKernel", 0) / (steps + 1), 1),
       "own_kernels_per_episode": dict(kernels),
       "verdict": "no per-step host synchronisation" if sum(sync.values()) + d2h < steps // 4 else "PER-STEP SYNC PRESENT"}
print(json.dumps(res))
