import sys, time, json
sys.path.insert(0, ".")
import torch
from baseline import refloop
def run(**kw):
    B = 256
    cfg = refloop.load_config("cleanup", seed=0, use_cuda=True, save_model=False, t_max=50000, runner="batched", batch_size_run=B,
                              buffer_size=4 * B, buffer_cpu_only=False, test_nepisode=B, test_interval=50000, log_interval=1000,
                              runner_log_interval=1000, learner_log_interval=1000, env_args=dict(num_agents=3, map="default3"), **kw)
    c = refloop.build_components(cfg, backend="b200")
    ts = []
    for ep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        b = c.runner.run(test_mode=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        c.buffer.insert_episode_batch(b)
        s = c.buffer.sample(16); s = s[:, :s.max_t_filled()]
        c.learner.train(s, c.runner.t_env, (ep + 1) * B)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((round(t1 - t0, 3), round(t2 - t1, 3)))
    c.runner.close_env()
    return ts
print("plain            ", run())
print("fused            ", run(fused_frontend=True))
print("fused+sel        ", run(fused_frontend=True, action_selector="epsilon_greedy_b200"))
print("fused+sel+learner", run(fused_frontend=True, action_selector="epsilon_greedy_b200", learner="homophily_learner_b200"))
print("groups4 all      ", run(fused_frontend=True, action_selector="epsilon_greedy_b200", learner="homophily_learner_b200", env_groups=4))
