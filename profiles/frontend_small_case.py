"""Launches the front end on 2 560 agent views (V=15) five times: workload for an `ncu --metrics gpu__time_duration.sum` launch list."""
import sys
import torch
sys.path.insert(0, ".")
from homophily_marl_b200.frontend import ObsFrontEnd
view, rows = 15, int(sys.argv[1]) if len(sys.argv) > 1 else 2560
N = 2 * view + 1
RP = (N + 3) // 4 * 4
PS, AS = N * RP, (3 * N * RP + 15) // 16 * 16
P = N - 2
torch.manual_seed(0)
mod = torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, 1), torch.nn.LeakyReLU(), torch.nn.Flatten(), torch.nn.Linear(6 * P * P, 32), torch.nn.LeakyReLU())
fe = ObsFrontEnd.from_module(mod, view, device="cuda:0")
buf = torch.randint(0, 256, (rows * AS,), dtype=torch.int32).to(torch.uint8).cuda()
for _ in range(5):
    fe.forward(buf, rows, AS, PS, RP)
torch.cuda.synchronize()
