"""Unchanged single-env EpisodeRunner.run() on the CUDA facade vs on the reference env, same process, alternating episodes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baseline import refloop
res = {}
comps = {}
for backend in ("b200", "reference"):
    cfg = refloop.load_config("cleanup", seed=0, use_cuda=True, save_model=False, t_max=10**9, test_nepisode=1,
                              env_args=dict(num_agents=3, map="default3"))
    comps[backend] = refloop.build_components(cfg, backend=backend)
for backend, c in comps.items():
    for _ in range(2):
        c.runner.run(test_mode=False)
times = {"b200": [], "reference": []}
for rep in range(6):
    for backend, c in comps.items():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c.runner.run(test_mode=False)
        torch.cuda.synchronize(); times[backend].append(time.perf_counter() - t0)
for k, v in times.items():
    print(k, "ms per env step: min %.2f median %.2f" % (min(v) * 10, sorted(v)[len(v)//2] * 10))
# where does the facade's time go?
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); comps["b200"].runner.run(test_mode=False); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
