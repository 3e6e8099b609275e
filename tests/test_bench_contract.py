"""bench.py prints ONE JSON line with the contract's keys (CPU arm here; CUDA arm under -m gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       env={**os.environ, **(env or {})}, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    return lines


def test_reference_arm_line():
    lines = _run(["--impl", "reference", "--steps", "30", "--warmup", "1", "--workload", "cleanup3_b4096", "--envs", "64", "--no-train"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "agent-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline_port"]["kind"] == "port" and d["cpu_baseline_port"]["value"] > 0
    from baseline import refloop
    assert d["cpu_baseline"]["kind"] == ("reference" if refloop.available() else "port")
    assert set(d["config"]) == {"workload", "env", "map", "num_agents", "view_size", "envs_per_gpu", "global_envs", "episode_limit",
                                "actions", "extra_args", "obs_color", "obs_format"}
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "u8" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--envs", "64"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_reference_arm_times_the_reference_training_loop():
    from baseline import refloop
    if not refloop.available():
        pytest.skip("reference sources absent")
    d = json.loads(_run(["--impl", "reference", "--steps", "10", "--warmup", "1", "--workload", "cleanup3_b4096", "--train-t-max", "150"])[0])
    t = d["train_e2e"]
    assert t["unit"] == "env-steps/s" and t["reference"]["env_steps"] == 200 and t["reference"]["value"] > 0


@pytest.mark.gpu
def test_cuda_arm_line():
    lines = _run(["--steps", "300", "--warmup", "20", "--envs", "512", "--e2e-steps", "5", "--no-extra", "--no-train"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 300 and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["gpu_launches"] >= 300
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 512 * 5 and e["d2h_bytes_per_step"] > 512 * 5 * 3 * 31 * 31 and 0 < e["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    assert d["replays"] >= 5 and d["replay_ms"]["min"] <= d["replay_ms"]["median"] <= d["replay_ms"]["max"]
    assert "l2" not in d["config"] and "parallelism" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_env_range_rule():
    """Small batches are stepped as env ranges (profiles/r2_notes.md sections 2-3); every range size must divide the batch."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert [bench.auto_groups(b) for b in (512, 2048, 4096, 8192, 16384, 65536, 4095, 10000)] == [8, 8, 8, 4, 8, 1, 1, 1]
    for b in (512, 2048, 4096, 8192, 16384):
        assert b % bench.auto_groups(b) == 0
