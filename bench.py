#!/usr/bin/env python
"""Benchmark of the SSD env hot path (fused step + observation render).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, all host threads)

Metric (BASELINE.json): agent-steps/s, device-timed, + fraction of the HBM roofline.
A "step" is one fused step+obs pass over the whole resident batch (one kernel launch;
the synchronous episode reset every `episode_limit` steps is inside the timed region).
Default workload = BASELINE configs[1]: Harvest map=default5, 5 agents, 4096 envs per GPU,
uniform random actions, yaml-default extra_args.  Weak scaling: every rank owns 4096 envs
keyed by global env id; no collective on the step path.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (env, map, num_agents, view_size, envs_per_gpu, BASELINE.json config it is)
    "harvest5_b4096": ("harvest", "default5", 5, 15, 4096, "configs[1] Harvest default5, 5 agents, 4096 envs"),
    "cleanup5_b4096": ("cleanup", "default5", 5, 7, 4096, "configs[0] map/agents at B=4096"),
    "cleanup10_b2048": ("cleanup", "default10", 10, 7, 2048, "configs[2] Cleanup default10, 16384 envs over 8 GPUs"),
    "cleanup3_b4096": ("cleanup", "default3", 3, 7, 4096, "configs[3] map/agents at B=4096"),
    "harvest5_b65536": ("harvest", "default5", 5, 15, 65536, "configs[4] Harvest default5, 65536 envs obs stress"),
    "cleanup5_b65536": ("cleanup", "default5", 5, 7, 65536, "configs[0] map/agents at B=65536 (large-batch regime)"),
    "cleanup10_b16384": ("cleanup", "default10", 10, 7, 16384, "configs[2] whole 16384-env job on ONE GPU"),
}
L2_BYTES = 126 * 2 ** 20
METRIC = "agent-steps/sec (step+obs, device-timed)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=500)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="harvest5_b4096", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA graphs")
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--random-spawn", action="store_true")
    ap.add_argument("--obs-color", default="simplified", choices=["simplified", "full"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-groups-probe", action="store_true", help="skip the informational pipelined-groups measurement")
    ap.add_argument("--ring", type=int, default=0,
                    help="DIAGNOSTIC: force the number of observation ring buffers (1 = L2-resident stores; not a valid bench number)")
    ap.add_argument("--masked-actions", action="store_true",
                    help="draw actions from the yaml-masked set ({0-4,8} cleanup / {0-4} harvest) instead of all n_actions")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def extra_args(a):
    e = dict(obs_color=a.obs_color)
    if a.random_spawn:
        e.update(random_spawn_point=True, random_spawn_rotation=None)
    return e


def workload_config(a, world):
    env, mp, n, view, B, what = WORKLOADS[a.workload]
    B = a.envs or B
    return dict(workload=f"{a.workload}: {what}", env=env, map=mp, num_agents=n, view_size=view, envs_per_gpu=B,
                global_envs=B * world, episode_limit=100,
                actions="uniform over the yaml-masked set" if a.masked_actions else "uniform over all n_actions",
                extra_args="yaml defaults" if not a.random_spawn else "random spawn point+rotation",
                obs_color=a.obs_color, obs_format="u8 RGB planes, pixel rows padded to a multiple of 4 bytes")


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, 10 ms)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.mask, self.max_mhz, self._stop, self._thr, self.h = [], 0, None, threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def sample(self):
        """One synchronous sample (also called right after the timed work is enqueued, so short runs get one under load)."""
        if self.h is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(0.002)

    def __enter__(self):
        if self.h is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU arm
def oracle_batch(cfg, B, seed, gid0, a):
    from homophily_marl_b200 import mapspec
    from oracle import oracle as O
    spec = mapspec.compile_map(cfg["env"], cfg["map"], cfg["num_agents"], cfg["view_size"], cfg["episode_limit"], obs_color=a.obs_color)
    return spec, O.OracleBatch.from_spec(spec, n_envs=B, seed=seed, env_gid0=gid0, random_spawn_point=a.random_spawn,
                                         spawn_rotation=None if a.random_spawn else 0)


def time_oracle(cfg, a, steps, warmup, budget_s, threads):
    """Times the C port of the reference env (step + get_obs) on a bounded sample of the workload."""
    B_full = cfg["envs_per_gpu"]
    spec, probe = oracle_batch(cfg, min(256, B_full), a.seed, 0, a)
    rs = np.random.RandomState(a.seed)
    probe.reset(threads=threads)
    acts = rs.randint(0, spec.n_actions, size=(probe.B, spec.n_agents)).astype(np.uint8)
    out = probe.step(acts, threads=threads)
    t0 = time.perf_counter()
    for _ in range(4):
        probe.step(acts, threads=threads, out=out)
    per_env = (time.perf_counter() - t0) / (4 * probe.B)
    B = int(max(64, min(B_full, budget_s / max(steps + warmup, 1) / per_env)))
    spec, ob = oracle_batch(cfg, B, a.seed, 0, a)
    ob.reset(threads=threads)
    actions = rs.randint(0, spec.n_actions, size=(min(steps + warmup, 64), B, spec.n_agents)).astype(np.uint8)
    out, t = None, 0
    for s in range(warmup):
        out = ob.step(actions[s % len(actions)], threads=threads, out=out)
        t += 1
        if t % cfg["episode_limit"] == 0:
            ob.reset(threads=threads)
    t0 = time.perf_counter()
    for s in range(steps):
        out = ob.step(actions[(s + warmup) % len(actions)], threads=threads, out=out)
        t += 1
        if t % cfg["episode_limit"] == 0:
            ob.reset(threads=threads)
    dt = time.perf_counter() - t0
    return dict(value=B * spec.n_agents * steps / dt, seconds=dt, envs=B, steps=steps, ms_per_step=dt / steps * 1e3)


def run_reference(a):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cfg = workload_config(a, 1)
    threads = os.cpu_count() or 1
    r = time_oracle(cfg, a, a.steps, a.warmup, budget_s=150.0, threads=threads)
    sample = f"{r['envs']} of {cfg['envs_per_gpu']} envs x {a.steps} steps (step+get_obs), {threads} OpenMP threads"
    cfg["parallelism"] = f"cpu x{threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "agent-steps/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is pure Python and cannot travel to the GPU box; this arm times oracle/ssd_oracle.c, "
                    "the C port pinned bit-exact against it (the Python reference itself: ~2.5k agent-steps/s/core, BASELINE.md)"}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- CUDA arm
def pipelined_groups_probe(cfg, a, dev, rank, groups=4, steps=400):
    """Informational: the same envs stepped as `groups` independent sub-batches on separate streams (double-buffered
    rollout, what an asynchronous sampler does).  The logic phase of one group overlaps the observation stores of another."""
    import torch
    from homophily_marl_b200.batch_env import SSDBatchEnv
    B, n = cfg["envs_per_gpu"], cfg["num_agents"]
    if B % groups:
        return None
    Bg = B // groups
    envs = [SSDBatchEnv(cfg["env"], Bg, n, map=cfg["map"], view_size=cfg["view_size"], episode_limit=10 ** 6,
                        extra_args=extra_args(a), seed=a.seed, device=dev, env_gid_base=rank * B + g * Bg) for g in range(groups)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]
    ring_n = max(2, int(np.ceil(2.2 * L2_BYTES / (B * envs[0].layout.obs_env_stride))))
    graphs = []
    for g, (e, s) in enumerate(zip(envs, streams)):
        acts = torch.randint(0, e.n_actions, (32, Bg, n), device=dev, dtype=torch.int32).to(torch.uint8)
        ring = [e.new_obs_buffer() for _ in range(ring_n)]
        with torch.cuda.stream(s):
            e.reset()
            for i in range(3):
                e.step(acts[i], obs_out=ring[i % ring_n])
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(steps):
                    e.step(acts[i % 32], obs_out=ring[i % ring_n])
        graphs.append((gr, acts, ring))
    torch.cuda.synchronize()
    ms = float("inf")
    for _ in range(3):                                        # the first replay also uploads the graphs
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(groups)]
        e0.record()
        for g in range(groups):
            streams[g].wait_event(e0)
            with torch.cuda.stream(streams[g]):
                graphs[g][0].replay()
                ends[g].record(streams[g])
        torch.cuda.synchronize()
        ms = min(ms, max(e0.elapsed_time(x) for x in ends))
    for e in envs:
        e.close()
    return {"groups": groups, "steps": steps, "value": B * n * steps / (ms * 1e-3), "unit": "agent-steps/s",
            "us_per_step": ms / steps * 1e3, "what": f"{groups} independent groups of {Bg} envs on {groups} streams (same global env ids)"}


def run_b200(a):
    import torch
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from homophily_marl_b200.batch_env import SSDBatchEnv

    cfg = workload_config(a, world)
    B, n, limit = cfg["envs_per_gpu"], cfg["num_agents"], cfg["episode_limit"]
    env = SSDBatchEnv(cfg["env"], B, n, map=cfg["map"], view_size=cfg["view_size"], episode_limit=limit,
                      extra_args=extra_args(a), seed=a.seed, device=dev, env_gid_base=rank * B)
    obs_bytes = B * env.layout.obs_env_stride
    ring_n = a.ring if a.ring > 0 else max(2, int(np.ceil(2.2 * L2_BYTES / obs_bytes)))
    ring = [env.new_obs_buffer() for _ in range(ring_n)]
    cfg["l2"] = f"obs written to a ring of {ring_n} buffers x {obs_bytes / 2**20:.1f} MiB (> {L2_BYTES >> 20} MiB L2), like an episode buffer"
    cfg["parallelism"] = f"env-sharded x{world} (no collective on the step path)"
    n_act_slots = 2 * limit
    g = torch.Generator(device=dev).manual_seed(a.seed * 1000 + rank)
    actions = torch.randint(0, env.n_actions, (n_act_slots, B, n), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    if a.masked_actions:                                       # disable_rotation_action / disable_fire_action (yaml defaults)
        allowed = torch.tensor([0, 1, 2, 3, 4] + ([8] if env.n_actions == 9 else []), device=dev, dtype=torch.uint8)
        actions = allowed[torch.randint(0, len(allowed), (n_act_slots, B, n), generator=g, device=dev)]

    state = {"t": 0}

    def one_step():
        t = state["t"]
        if t % limit == 0:
            env.reset(obs_out=ring[t % ring_n])
        env.step(actions[t % n_act_slots], obs_out=ring[(t + 1) % ring_n])
        state["t"] = t + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    period = int(np.lcm.reduce([limit, n_act_slots, ring_n]))      # schedule repeats with this period
    use_graph = not a.no_graph
    with torch.cuda.stream(stream):
        for _ in range(max(a.warmup, 3)):
            one_step()
        stream.synchronize()
        graphs = []
        if use_graph:
            # capture the K timed steps as CUDA graphs of <= `chunk` steps (launch-bound inner loop -> graph)
            chunk = period if period <= 2000 else 1000
            t_first = state["t"]
            todo, cache = a.steps, {}
            while todo > 0:
                k = min(chunk, todo)
                key = (state["t"] % period, k)
                if key not in cache:
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=stream):
                        for _ in range(k):
                            one_step()
                    cache[key] = gr
                else:
                    state["t"] += k
                graphs.append(cache[key])
                todo -= k
            # capture does not execute: the env state is still at t_first
            state["t"] = t_first
        launches0 = env.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with ClockSampler(local_rank) as clk:
            ev0.record(stream)
            if use_graph:
                for gr in graphs:
                    gr.replay()
            else:
                for _ in range(a.steps):
                    one_step()
            ev1.record(stream)
            clk.sample()                                      # the GPU is still executing the enqueued steps here
            barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        if use_graph:
            n_resets = sum(1 for t in range(t_first, t_first + a.steps) if t % limit == 0)
            launches = a.steps + n_resets
            state["t"] = t_first + a.steps
        else:
            launches = env.launch_count - launches0

        # ---- e2e: the reference-facing call with HOST buffers (H2D actions, D2H results+obs every step)
        io = env.make_host_io(with_obs=True)
        h_actions = actions[: min(n_act_slots, 64)].cpu()
        e2e_steps = max(1, min(a.e2e_steps, a.steps))
        for s in range(3):
            io["actions"].copy_(h_actions[s % len(h_actions)])
            env.step_host(io)
        barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            io["actions"].copy_(h_actions[s % len(h_actions)])
            env.step_host(io)
            if bool(io["done"][0]):
                env.reset(obs=False)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

    times = torch.tensor([elapsed_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(times[0]), float(times[1])

    line = None
    if rank == 0:
        agent_steps = B * n * world * a.steps
        value = agent_steps / (elapsed_ms * 1e-3)
        alg_bytes = env.bytes_per_env_step() * B                       # per launch
        avg_launch_s = elapsed_ms * 1e-3 / a.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / avg_launch_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(a.workload, {}).get("dram_bytes_per_launch")
        d2h = B * (2 * n + 2 + 1) + obs_bytes
        line = {"metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                "clocks": clk.summary(), "gpu_launches": launches,
                "launch_mode": "cuda-graph" if use_graph else "python-loop",
                "e2e": {"value": B * n * world * e2e_steps / (e2e_ms * 1e-3), "unit": "agent-steps/s",
                        "h2d_bytes_per_step": B * n, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "what": "ssd_step_host: pinned actions H2D, kernel, reward/clean/apple_cnt/done/obs D2H, stream sync"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "ssd_kernel<MODE_STEP>", "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "bytes_per_env_step": env.bytes_per_env_step(),
                             "avg_launch_us": avg_launch_s * 1e6}}
        if world == 1 and not a.no_groups_probe:
            line["pipelined_groups"] = pipelined_groups_probe(cfg, a, dev, rank)
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = time_oracle(cfg, a, steps=20, warmup=3, budget_s=20.0, threads=threads)
            line["cpu_baseline"] = {"value": r["value"], "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{r['envs']} envs x 20 steps of the same workload (C port of the reference env, step+get_obs)"}
            r1 = time_oracle(cfg, a, steps=10, warmup=2, budget_s=6.0, threads=1)
            line["cpu_baseline_1core"] = {"value": r1["value"], "unit": "agent-steps/s", "cores": 1, "kind": "port",
                                          "sample": f"{r1['envs']} envs x 10 steps; the reference itself is single-env, single-core "
                                                    "(episode_runner.py:13) and ~50x slower than this C port (BASELINE.md)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))
