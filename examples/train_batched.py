"""The reference's own training loop (run.run_sequential) on the B200 stack: batched runner over the CUDA env, fused u8
observation front end, device epsilon-greedy, device learner.  MAC, agent network and replay buffer are the reference's.

    python baseline/fetch_ref.py                 # once, where the reference is mounted (copies it to baseline/_ref/)
    python examples/train_batched.py [envs] [env_steps]

Equivalent to adding these lines to the reference's registries and running
`python src/main.py --config=homophily --env-config=cleanup with runner=batched batch_size_run=256 ...`.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import refloop  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t_max = int(sys.argv[2]) if len(sys.argv) > 2 else 10 * B * 100
cfg = refloop.load_config("cleanup", seed=0, use_cuda=True, save_model=False, t_max=t_max,
                          runner="batched", batch_size_run=B, buffer_size=max(B, 1024), buffer_cpu_only=False,
                          fused_frontend=True, action_selector="epsilon_greedy_b200", learner="homophily_learner_b200",
                          test_nepisode=B, test_interval=t_max, log_interval=B * 100, runner_log_interval=B * 100,
                          learner_log_interval=B * 100, env_args=dict(num_agents=3, map="default3"))
r = refloop.run_training(cfg, backend="b200", log_level="INFO")
steps = (t_max // (B * 100) + 1) * B * 100
print(f"{steps} env steps in {r['seconds']:.1f} s = {steps / r['seconds']:.0f} env-steps/s")
for k in ("return_mean", "collective_return_mean", "equality_metric_mean", "loss_value_env", "loss_value_inc", "loss_sim", "epsilon"):
    if k in r["stats"]:
        print(f"  {k}: {r['stats'][k]:.4f}")
