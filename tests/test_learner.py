"""Learner-side rows.

GPU: ``DeviceHomophilyLearner`` (incentive bookkeeping in one kernel, device-side similarity clusters) takes the same step as
the reference ``HomophilyLearner.cal_loss_and_step`` (src/learners/homophily_learner.py:51-247) on a recorded batch: same three
losses, same parameters after the two Adam steps.  CPU (gloo, world size 2): the flat-bucket gradient all-reduce and the
episode sharding of the data-parallel path.
"""
import copy
import os
import socket
import sys

import numpy as np
import pytest

from baseline import refloop

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(not refloop.available(), reason="reference sources absent (baseline/_ref)")
@pytest.mark.parametrize("others", [False, True])
def test_device_learner_takes_the_reference_step_on_a_recorded_batch(others):
    from homophily_marl_b200.learner import DeviceHomophilyLearner
    B = 24
    cfg = refloop.load_config("cleanup", seed=5, use_cuda=True, save_model=False, runner="batched", batch_size_run=B,
                              buffer_size=2 * B, batch_size=12, buffer_cpu_only=False, test_nepisode=B, consider_others_inc=others,
                              learner_log_interval=1, env_args=dict(num_agents=3, map="default3", episode_limit=30,
                                                                    extra_args=dict(disable_fire_action=False)))
    c = refloop.build_components(cfg, backend="b200")
    c.buffer.insert_episode_batch(c.runner.run(test_mode=False))
    sample = c.buffer.sample(12)
    sample = sample[:, :sample.max_t_filled()]
    assert (sample["clean_num"] > 0).any() and (sample["actions_inc"] != 0).any()      # the incentive terms are exercised
    sd0 = copy.deepcopy(c.mac.agent.state_dict())
    logs_ref = {k: float(v) for k, v in c.learner.cal_loss_and_step(sample).items()}
    after_ref = copy.deepcopy(c.mac.agent.state_dict())
    c.mac.agent.load_state_dict(sd0)
    ours = DeviceHomophilyLearner(c.mac, c.buffer.scheme, c.logger, c.args)
    logs = {k: float(v) for k, v in ours.cal_loss_and_step(sample).items()}
    assert set(logs) == set(logs_ref)
    for k in logs_ref:
        assert logs[k] == pytest.approx(logs_ref[k], rel=2e-5, abs=1e-7), k
    moved = 0.0
    for k, v in c.mac.agent.state_dict().items():
        assert torch.allclose(v, after_ref[k], rtol=1e-4, atol=2e-6), k
        moved = max(moved, (v - sd0[k]).abs().max().item())
    assert moved > 1e-4                                                               # the step really changed the parameters
    ours.train(sample, t_env=3000, episode_num=40)                                    # reference surface: logging + target update
    assert np.isfinite(c.logger.stats["loss_sim"][-1][1])
    c.runner.close_env()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from homophily_marl_b200.learner import FlatBucket, shard_episodes
    torch.manual_seed(rank)                                                           # ranks start with different parameters
    conv = torch.nn.Parameter(torch.randn(6, 3, 3, 3))
    env_w, inc_w, frozen = torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(4)), torch.nn.Parameter(torch.randn(3), requires_grad=False)
    bucket = FlatBucket([conv, inc_w, conv, env_w, frozen])                           # conv sits in both Adam groups: reduced once
    assert len(bucket.params) == 3 and bucket.numel == 162 + 35 + 4
    bucket.broadcast_params(0)
    conv.grad = torch.full_like(conv, float(rank + 1))
    env_w.grad = torch.full_like(env_w, 10.0 * (rank + 1))                            # inc_w has no gradient on this step
    bucket.all_reduce_mean()
    mean = sum(r + 1 for r in range(world)) / world

    class Episodes:                                                                   # the slicing surface of EpisodeBatch
        batch_size = 7

        def __getitem__(self, sl):
            return list(range(7))[sl]
    shard = shard_episodes(Episodes(), rank, world)
    torch.save({"conv": conv.data, "conv_grad": conv.grad, "env_grad": env_w.grad, "inc_grad": inc_w.grad, "mean": mean, "shard": shard},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_episode_shards_world2(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt")) for k in range(world)]
    assert torch.equal(r[0]["conv"], r[1]["conv"])                                    # parameters were broadcast from rank 0
    for k in range(world):
        assert torch.all(r[k]["conv_grad"] == r[k]["mean"]) and torch.all(r[k]["env_grad"] == 10 * r[k]["mean"])
        assert torch.all(r[k]["inc_grad"] == 0)                                       # missing gradients count as zeros on every rank
    assert r[0]["shard"] + r[1]["shard"] == list(range(7)) and len(r[0]["shard"]) == 4


@pytest.mark.gpu
@pytest.mark.skipif(not refloop.available(), reason="reference sources absent (baseline/_ref)")
def test_data_parallel_learner_on_two_gpus_nccl():
    """f4 on real hardware: two ranks, one flat-bucket NCCL all-reduce per learner step (profiles/dp_learner_check.py)."""
    import json
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "profiles", "dp_learner_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["world"] == 2 and res["start_identical"]
    for s in res["steps"]:
        assert s["bucket_equals_mean_of_local_grads"] and s["params_identical_after_step"] and s["local_grads_differ_between_ranks"]
