"""Builds the front end with several stage geometries (-DFE_OC_PER_STAGE / FE_STAGES / FE_BSTAGES, optional diagnostics) into
build_variants/ (run here, no GPU), or times every built variant on the GPU (run with `time`): V=15 and V=7, 20 480 views,
plus the parity test of tests/test_gpu_frontend.py against each variant."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "g3_s2_b4": [],
    "g3_s3_b4": ["-DFE_STAGES=3"],
    "g2_s3_b6": ["-DFE_OC_PER_STAGE=2", "-DFE_STAGES=3", "-DFE_BSTAGES=6"],
    "g2_s4_b6": ["-DFE_OC_PER_STAGE=2", "-DFE_STAGES=4", "-DFE_BSTAGES=6"],
    "g2_s6_b8": ["-DFE_OC_PER_STAGE=2", "-DFE_STAGES=6", "-DFE_BSTAGES=8"],
    "g1_s6_b12": ["-DFE_OC_PER_STAGE=1", "-DFE_STAGES=6", "-DFE_BSTAGES=12"],
    "g1_s12_b16": ["-DFE_OC_PER_STAGE=1", "-DFE_STAGES=12", "-DFE_BSTAGES=16"],
}

if sys.argv[1:] == ["build"]:
    from homophily_marl_b200 import _build
    os.makedirs(os.path.join(ROOT, "build_variants"), exist_ok=True)
    for name, flags in VARIANTS.items():
        out = os.path.join(ROOT, "build_variants", f"libssd_b200_{name}.so")
        _build.build(force=True, extra_flags=flags, out=out)
        print("built", out, flush=True)
elif sys.argv[1:] == ["time"]:
    import torch
    from homophily_marl_b200.frontend import ObsFrontEnd
    out = {}
    for view, rows in ((15, 20480), (15, 18944), (7, 20480), (15, 2560)):
        N = 2 * view + 1
        RP = (N + 3) // 4 * 4
        PS, AS = N * RP, (3 * N * RP + 15) // 16 * 16
        P = N - 2
        torch.manual_seed(0)
        mod = torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, 1), torch.nn.LeakyReLU(), torch.nn.Flatten(), torch.nn.Linear(6 * P * P, 32), torch.nn.LeakyReLU())
        fe = ObsFrontEnd.from_module(mod, view, device="cuda:0")
        buf = torch.randint(0, 256, (rows * AS,), dtype=torch.int32).to(torch.uint8).cuda()
        for _ in range(3):
            fe.forward(buf, rows, AS, PS, RP)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fe.forward(buf, rows, AS, PS, RP)
        e1.record()
        torch.cuda.synchronize()
        out[f"V{view}x{rows}"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    print(out)
else:
    for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libssd_b200_*.so"))):
        env = {**os.environ, "SSD_B200_LIB": lib}
        r = subprocess.run(["timeout", "200", sys.executable, __file__, "time"], capture_output=True, text=True, env=env)
        res = r.stdout.strip() or ("ERR " + r.stderr[-300:])
        t = ""
        if "g_full" in lib and os.environ.get("FE_TESTS"):
            r = subprocess.run(["timeout", "300", sys.executable, "-m", "pytest", "tests/test_gpu_frontend.py", "-m", "gpu", "-q", "-x"],
                               capture_output=True, text=True, env=env, cwd=ROOT)
            t = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:]
        print(os.path.basename(lib), res, "|", t, flush=True)
