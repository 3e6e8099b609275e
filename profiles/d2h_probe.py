"""Diagnoses the end-to-end (host-buffer) path at N ranks: concurrent pinned D2H copies of one observation batch per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/d2h_probe.py

Per rank: 61 MB (4096 Harvest envs x 5 agents x 2976 B) device -> pinned host, 30 copies, all ranks released by one barrier.
Variants: plain; host threads bound to the GPU's NUMA-local CPUs (nvmlDeviceSetCpuAffinity) BEFORE the pinned allocation
(first touch); the copy split into 4 chunks on 4 streams.  Prints one JSON line on rank 0 (aggregate and per-rank GB/s)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 4096 * 5 * 2976
src = torch.randint(0, 255, (NBYTES,), dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(chunks, reps=30):
    host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    host.zero_()                                               # first touch on the current CPU set
    streams = [torch.cuda.Stream() for _ in range(chunks)]
    step = (NBYTES + chunks - 1) // chunks
    for _ in range(3):
        host.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for c, s in enumerate(streams):
            with torch.cuda.stream(s):
                host[c * step:(c + 1) * step].copy_(src[c * step:(c + 1) * step], non_blocking=True)
        for s in streams:
            s.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return NBYTES * reps / dt / 1e9


def gather(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world == 1:
        return [v]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x) for x in out]


res = {"world": world, "mb_per_copy": NBYTES / 1e6, "cpus": os.cpu_count()}
res["plain"] = gather(run(1))
res["chunks4"] = gather(run(4))
try:
    import pynvml
    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
    h = pynvml.nvmlDeviceGetHandleByIndex(phys)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    res["affinity_cpus"] = gather(float(len(os.sched_getaffinity(0))))
    res["numa_bound"] = gather(run(1))
    res["numa_bound_chunks4"] = gather(run(4))
except Exception as e:
    res["affinity_error"] = f"{type(e).__name__}: {e}"[:200]
if rank == 0:
    for k in ("plain", "chunks4", "numa_bound", "numa_bound_chunks4"):
        if k in res:
            res[k] = {"per_rank_gbs": [round(x, 2) for x in res[k]], "aggregate_gbs": round(sum(res[k]), 1)}
    try:
        res["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        pass
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
