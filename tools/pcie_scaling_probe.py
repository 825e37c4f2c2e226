#!/usr/bin/env python
"""Aggregate device->host copy rate of N ranks copying 2 MB results concurrently, with and without binding
each rank's CPU threads (and therefore its first-touched pinned memory) to the NUMA node of its GPU.

    torchrun --nproc-per-node N tools/pcie_scaling_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("gloo")


def measure(tag):
    nbytes, reps = 2 << 20, 400
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dsts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for d in dsts:
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(reps):
        dsts[i & 3].copy_(src, non_blocking=True)
    s1.record()
    torch.cuda.synchronize()
    gbs = nbytes * reps / (s0.elapsed_time(s1) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        print(json.dumps({"case": tag, "ranks": world, "aggregate_d2h_gbs": round(float(t), 1), "rank0_gbs": round(gbs, 1)}), flush=True)


measure("no affinity (cpus %s)" % sorted(os.sched_getaffinity(0))[:4])
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local).uuid)).encode())
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 4)
    cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
    allowed = sorted(set(cpus) & os.sched_getaffinity(0))
    if rank == 0 or rank == world - 1:
        print(f"rank {rank}: NVML ideal cpus {cpus[:6]}..{cpus[-3:] if cpus else []} ({len(cpus)}), allowed here {allowed[:8]}", flush=True)
    if allowed:
        os.sched_setaffinity(0, allowed)
        measure("bound to the GPU's NUMA cpus")
except Exception as e:  # pragma: no cover
    if rank == 0:
        print("affinity probe failed:", repr(e), flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i 'numa\\|^CPU(s)' | head -6")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
