"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

Drives the UNMODIFIED upper layers of the reference (``run.run_sequential``, ``EpisodeRunner``, ``HomophilyMAC``,
``HomophilyLearner``, ``EpisodeBatch`` / ``ReplayBuffer``) from ``/root/reference/src`` (dev container) or from the
git-ignored copy ``baseline/_ref/src`` (GPU box, made by ``baseline/fetch_ref.py``), either on the reference's own CPU
env or on this repo's CUDA env registered under the same names -- the two-line swap of INTEGRATION.md done at run time
on the registry dicts, so that no reference file is edited.

What has to be stubbed for the reference to import in this image: matplotlib (replay plots only) and
pyclustering (learner's x-means; stand-in in ``baseline/stubs``, PARITY UNPINNED).  ``main.py`` itself cannot run
(sacred is absent; ``collections.Mapping``; ``yaml.load`` without a Loader), so the YAML merge of ``main.py:75-100`` is
restated here with ``yaml.safe_load`` and ``run.run_sequential`` is called directly, as ``run.run`` does (run.py:19-54).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import logging
import os
import sys
import time
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUBS = os.path.join(HERE, "stubs")


def ref_src() -> str | None:
    for cand in (os.path.join(os.environ.get("SSD_REFERENCE_ROOT", "/root/reference"), "src"),
                 os.path.join(HERE, "_ref", "src")):
        if os.path.isdir(os.path.join(cand, "envs", "ssd")):
            return cand
    return None


def available() -> bool:
    return ref_src() is not None


def import_reference() -> str:
    """Puts the reference's ``src`` (and stubs for the packages this image lacks) on sys.path."""
    src = ref_src()
    if src is None:
        raise RuntimeError("no reference sources: neither /root/reference/src nor baseline/_ref/src exists "
                           "(run `python baseline/fetch_ref.py` in the dev container)")
    need_stub = False
    for mod in ("matplotlib", "pyclustering"):
        if mod in sys.modules:
            continue
        try:
            importlib.import_module(mod)
        except Exception:
            need_stub = True
    if need_stub and STUBS not in sys.path:
        sys.path.append(STUBS)                                  # after site-packages: real packages win when present
    if src not in sys.path:
        sys.path.insert(0, src)
    return src


def _merge(d, u):
    for k, v in u.items():                                     # main.py:57-63 recursive_dict_update
        if isinstance(v, dict):
            d[k] = _merge(d.get(k, {}) or {}, v)
        else:
            d[k] = v
    return d


def load_config(env_config: str, alg_config: str = "homophily", seed: int = 0, **overrides) -> dict:
    """default.yaml <- envs/<env>.yaml <- algs/<alg>.yaml <- overrides (main.py:75-100); env_args.seed as main.py:34."""
    import yaml
    cfg_dir = os.path.join(import_reference(), "config")
    with open(os.path.join(cfg_dir, "default.yaml")) as f:
        cfg = yaml.safe_load(f)
    for sub, name in (("envs", env_config), ("algs", alg_config)):
        with open(os.path.join(cfg_dir, sub, name + ".yaml")) as f:
            cfg = _merge(cfg, yaml.safe_load(f))
    env_over = overrides.pop("env_args", None)
    cfg.update(overrides)
    if env_over:
        _merge(cfg["env_args"], env_over)
    cfg["seed"] = seed
    cfg["env_args"]["seed"] = seed
    cfg["use_tensorboard"] = False                              # tensorboard_logger is absent
    return cfg


def make_args(cfg: dict):
    """run.run (run.py:19-28): sanity check, namespace, device."""
    import torch
    run = importlib.import_module("run")
    log = logging.getLogger("refloop")
    cfg = run.args_sanity_check(dict(cfg), log)
    args = SimpleNamespace(**cfg)
    args.device = "cuda" if args.use_cuda else "cpu"
    args.unique_token = "refloop"
    if args.use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))      # one process per GPU under torchrun
    return args


def quiet_logger(level="WARNING"):
    from utils.logging import Logger, get_logger
    lg = get_logger()
    lg.setLevel(level)
    return Logger(lg)


def register_b200() -> None:
    """INTEGRATION.md section 1, applied to the live registry dicts: the reference's ``envs.REGISTRY`` entries are
    replaced by this repo's facade, and ``runners.REGISTRY['batched']`` is added.  No reference file changes."""
    import_reference()
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    import envs as ref_envs
    import runners as ref_runners
    from homophily_marl_b200 import pymarl_env
    from homophily_marl_b200.batched_runner import BatchedEpisodeRunner
    if "_reference_registry" not in ref_envs.__dict__:
        ref_envs._reference_registry = dict(ref_envs.REGISTRY)
    ref_envs.REGISTRY.update(pymarl_env.REGISTRY)
    ref_runners.REGISTRY["batched"] = BatchedEpisodeRunner
    import components.action_selectors as ref_selectors
    from homophily_marl_b200 import selectors
    ref_selectors.REGISTRY.update(selectors.REGISTRY)        # 'epsilon_greedy_b200': one kernel instead of seven torch launches
    import learners as ref_learners
    from homophily_marl_b200 import learner
    ref_learners.REGISTRY.update(learner.REGISTRY)           # 'homophily_learner_b200': incentive kernel, device clusters, DP all-reduce


def _python_bool_terminated(ctor):
    """The reference's ``terminated`` is a ``numpy.bool_``; with torch >= 2 / numpy >= 2 ``EpisodeBatch.update`` cannot turn
    ``[(np.bool_,)]`` into a uint8 tensor (episode_buffer.py:106, SURVEY 8b pitfall).  The env instance gets its ``step``
    wrapped to return a Python ``bool`` -- the only adaptation the reference needs to run its own loop in this image."""
    def make(**kw):
        env = ctor(**kw)
        inner = env.step

        def step(actions):
            reward, terminated, info = inner(actions)
            return reward, bool(terminated), info
        env.step = step
        return env
    return make


def register_reference_env() -> None:
    import_reference()
    import envs as ref_envs
    if "_reference_registry" not in ref_envs.__dict__:
        ref_envs._reference_registry = dict(ref_envs.REGISTRY)
    for k, ctor in ref_envs._reference_registry.items():
        ref_envs.REGISTRY[k] = _python_bool_terminated(ctor)


def run_training(cfg: dict, backend: str = "b200", log_level="WARNING") -> dict:
    """``run.run_sequential`` (run.py:81-244) end to end: rollouts + replay buffer + learner updates + test episodes.
    backend: 'b200' (CUDA env behind the reference's registry) or 'reference' (the reference's own NumPy env)."""
    import numpy as np
    import torch
    import_reference()
    if backend == "b200":
        register_b200()
    else:
        register_reference_env()
    run = importlib.import_module("run")
    np.random.seed(cfg["seed"])                                # main.py:32-33
    torch.manual_seed(cfg["seed"])
    args = make_args(cfg)
    logger = quiet_logger(log_level)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):           # env constructors print their map
        run.run_sequential(args=args, logger=logger)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = {k: v[-1][1] for k, v in logger.stats.items() if v}
    # run_sequential leaves `while t_env <= t_max`, so the number of env steps actually run is read from the episode
    # counter it logs (run.py:236-240) when available, else from t_max
    steps = None
    if logger.stats.get("episode"):
        steps = logger.stats["episode"][-1][0]
    return {"seconds": dt, "stats": stats, "t_env_logged": steps, "t_max": cfg["t_max"], "logger": logger}


# --------------------------------------------------------------------------- the reference's CPU env, timed
def _ref_env(env_name, map_name, n, view, episode_limit, extra=None):
    import_reference()
    import envs as ref_envs
    reg = ref_envs.__dict__.get("_reference_registry", ref_envs.REGISTRY)
    extra_args = dict(random_spawn_point=False, random_spawn_rotation=0, disable_rotation_action=True,
                      disable_fire_action=True, obs_color="simplified")
    extra_args.update(extra or {})
    with contextlib.redirect_stdout(io.StringIO()):
        env = reg[env_name](num_agents=n, render=False, episode_limit=episode_limit, is_replay=False, view_size=view,
                            map=map_name, extra_args=extra_args)
    if env_name == "harvest" and not hasattr(env, "SPAWN_PROB"):
        import envs.ssd.harvest as hv
        env.SPAWN_PROB = list(hv.SPAWN_PROB)
    return env


def _time_one(task):
    """One process: the unmodified ``MapEnv.step() + get_obs()`` (map_env.py:874-945), uniform random actions."""
    env_name, map_name, n, view, steps, warmup, seed, budget_s = task
    import random
    import numpy as np
    np.random.seed(seed)
    random.seed(seed)
    env = _ref_env(env_name, map_name, n, view, episode_limit=100)
    env.reset()
    rs = np.random.RandomState(seed)
    acts = rs.randint(0, env.n_actions, size=(steps + warmup, n))
    t = 0
    for s in range(warmup):
        _, term, _ = env.step(acts[s])
        env.get_obs()
        if term:
            env.reset()
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        _, term, _ = env.step(acts[warmup + s])
        env.get_obs()
        done += 1
        if term:
            env.reset()
        if (s & 15) == 15 and time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


def time_reference_env(env_name, map_name, n, view, steps=1000, warmup=50, procs=1, seed=0, budget_s=30.0) -> dict:
    """agent-steps/s of the reference's own env on `procs` host cores (independent single-env processes: the reference has
    no parallel runner, episode_runner.py:13).  Each process runs <= `steps` steps or `budget_s` seconds."""
    tasks = [(env_name, map_name, n, view, steps, warmup, seed + i, budget_s) for i in range(procs)]
    if procs == 1:
        res = [_time_one(tasks[0])]
        wall = res[0][1]
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            res = pool.map(_time_one, tasks)
        wall = max(r[1] for r in res)
    total_steps = sum(r[0] for r in res)
    return {"value": total_steps * n / wall, "env_steps": total_steps, "seconds": wall, "procs": procs,
            "ms_per_env_step_per_proc": 1e3 * sum(r[1] for r in res) / max(total_steps, 1)}


# --------------------------------------------------------------------------- components, as run_sequential builds them
def build_components(cfg: dict, backend: str = "b200", runner_name: str | None = None):
    """The objects ``run_sequential`` wires together before its training loop (run.py:84-135), returned to the caller
    so that a test can call ``runner.run()`` / ``learner.train()`` itself.  Same construction order, same scheme."""
    import torch as th
    import_reference()
    if backend == "b200":
        register_b200()
    else:
        register_reference_env()
    from components.episode_buffer import ReplayBuffer
    from components.transforms import OneHot
    from controllers import REGISTRY as mac_REGISTRY
    from learners import REGISTRY as le_REGISTRY
    from runners import REGISTRY as r_REGISTRY
    args = make_args(cfg)
    if runner_name:
        args.runner = runner_name
    logger = quiet_logger()
    with contextlib.redirect_stdout(io.StringIO()):
        runner = r_REGISTRY[args.runner](args=args, logger=logger)
    env_info = runner.get_env_info()
    args.n_agents, args.n_actions = env_info["n_agents"], env_info["n_actions"]
    args.state_shape, args.obs_shape = env_info["state_shape"], env_info["obs_shape"]
    if args.rgb_input:
        args.state_dims, args.obs_dims = env_info["state_dims"], env_info["obs_dims"]
    scheme = {
        "state": {"vshape": env_info["state_shape"]},
        "obs": {"vshape": env_info["obs_shape"], "group": "agents"},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
        "avail_actions": {"vshape": (env_info["n_actions"],), "group": "agents", "dtype": th.int},
        "reward": {"vshape": (1,) if not args.ind_reward else (args.n_agents,)},
        "terminated": {"vshape": (1,), "dtype": th.uint8},
        "clean_num": {"vshape": (args.n_agents,)},
        "apple_den": {"vshape": (args.n_agents,)},
        "agent_pos": {"vshape": (args.n_agents, 2)},
        "agent_orientation": {"vshape": (args.n_agents, 2)},
    }
    if "homophily" in args.name:
        scheme["actions_inc"] = {"vshape": (args.n_agents, 1), "group": "agents", "dtype": th.long}
    groups = {"agents": args.n_agents}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=args.n_actions if not args.action_double else args.n_actions * 2)])}
    buffer = ReplayBuffer(scheme, groups, args.buffer_size, env_info["episode_limit"] + 1, preprocess=preprocess,
                          device="cpu" if args.buffer_cpu_only else args.device)
    mac = mac_REGISTRY[args.mac](buffer.scheme, groups, args)
    runner.setup(scheme=scheme, groups=groups, preprocess=preprocess, mac=mac)
    learner = le_REGISTRY[args.learner](mac, buffer.scheme, logger, args)
    if args.use_cuda:
        learner.cuda()
    return SimpleNamespace(args=args, logger=logger, runner=runner, buffer=buffer, mac=mac, learner=learner, scheme=scheme,
                           groups=groups, env_info=env_info)
