"""Times the fused front end for 20480 agent views at several split-K factors (SSD_B200_FRONTEND_SPLIT)."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from homophily_marl_b200.frontend import ObsFrontEnd
    view, rows = int(sys.argv[1]), int(sys.argv[2])
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
    print("%.1f" % (e0.elapsed_time(e1) / 20 * 1e3))
else:
    for view in (7, 15):
        for rows in (20480, 2560):
            res = {}
            for split in (1, 2, 3, 4, 6, 8, 12, 16):
                r = subprocess.run([sys.executable, __file__, str(view), str(rows)], capture_output=True, text=True,
                                   env={**os.environ, "SSD_B200_FRONTEND_SPLIT": str(split)})
                res[split] = r.stdout.strip() or r.stderr[-200:]
            print(f"view {view} rows {rows}: us by split {res}", flush=True)
