"""Data-parallel learner on real GPUs (NCCL): run under torchrun with >= 2 ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 profiles/dp_learner_check.py

Every rank builds the same batched Cleanup default3 rollout (same seed -> same replay buffer and the same sampled episodes),
then `DeviceHomophilyLearner.train` shards the sample by rank, averages the gradients of both Adam groups with ONE flat-bucket
NCCL all-reduce and steps.  Checks: parameters start identical (broadcast), stay bit-identical across ranks after every step,
move, and the averaged gradient equals the mean of the ranks' local gradients.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import refloop  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ["SSD_B200_DEVICE"] = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

B = 32
cfg = refloop.load_config("cleanup", seed=11, use_cuda=True, save_model=False, runner="batched", batch_size_run=B, buffer_size=2 * B,
                          batch_size=16, buffer_cpu_only=False, test_nepisode=B, learner="homophily_learner_b200",
                          action_selector="epsilon_greedy_b200", learner_log_interval=1,
                          env_args=dict(num_agents=3, map="default3", episode_limit=25))
c = refloop.build_components(cfg, backend="b200")
torch.cuda.set_device(local)
learner = c.learner
assert type(learner).__name__ == "DeviceHomophilyLearner" and learner._dist


def flat(params):
    return torch.cat([p.detach().reshape(-1).float() for p in params])


def same_on_all_ranks(v):
    out = [torch.zeros_like(v) for _ in range(world)]
    dist.all_gather(out, v)
    return all(torch.equal(out[0], o) for o in out)


learner._sync_start()
p0 = flat(learner.bucket.params)
ok_start = same_on_all_ranks(p0)
c.buffer.insert_episode_batch(c.runner.run(test_mode=False))
res = {"world": world, "params": int(p0.numel()), "start_identical": ok_start, "steps": []}
for step in range(2):
    np.random.seed(100 + step)                                  # ReplayBuffer.sample draws with NumPy: same episodes on every rank
    sample = c.buffer.sample(16)
    sample = sample[:, :sample.max_t_filled()]
    # the mean of the ranks' LOCAL gradients, computed without the bucket
    from homophily_marl_b200.learner import shard_episodes
    shard = shard_episodes(sample, rank, world)
    le, li, ls, _ = learner.losses(shard)
    learner.optimiser_inc.zero_grad()
    learner.optimiser_env.zero_grad()
    (li + le + ls * c.args.sim_loss_weight).backward()
    local_grad = flat([p.grad if p.grad is not None else torch.zeros_like(p) for p in learner.bucket.params])
    mean_grad = local_grad.clone()
    dist.all_reduce(mean_grad)
    mean_grad /= world
    learner.bucket.all_reduce_mean()
    bucket_grad = flat([p.grad for p in learner.bucket.params])
    grads_ok = torch.allclose(bucket_grad, mean_grad, rtol=0, atol=0)
    differs_locally = not torch.equal(local_grad, mean_grad)
    learner.train(sample, t_env=1000 * (step + 1), episode_num=B * (step + 1))
    p1 = flat(learner.bucket.params)
    res["steps"].append({"bucket_equals_mean_of_local_grads": bool(grads_ok), "local_grads_differ_between_ranks": bool(differs_locally),
                         "params_identical_after_step": same_on_all_ranks(p1), "max_param_change": float((p1 - p0).abs().max()),
                         "loss_value_env": float(c.logger.stats["loss_value_env"][-1][1])})
    p0 = p1
if rank == 0:
    print(json.dumps(res), flush=True)
    assert res["start_identical"] and all(s["bucket_equals_mean_of_local_grads"] and s["params_identical_after_step"]
                                          and s["max_param_change"] > 0 and np.isfinite(s["loss_value_env"]) for s in res["steps"])
dist.barrier()
dist.destroy_process_group()
