"""Write-only vs copy HBM bandwidth on this GPU (context for the obs-write dominated step kernel)."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for mb in (1024, 2048):
    n = mb * 2 ** 20
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    b = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn, bytes_moved in (("fill", lambda: a.fill_(7), n), ("copy", lambda: b.copy_(a), 2 * n)):
        for _ in range(3):
            fn()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"{name}_{mb}MiB_GBps"] = bytes_moved / best / 1e6
print(json.dumps(out))
