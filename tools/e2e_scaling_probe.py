#!/usr/bin/env python
"""Which part of the host end-to-end step loses throughput when N ranks share one host?  Every rank times the
same K steps in several variants (copies only, kernels only, fewer lanes) and rank 0 prints the aggregate.

    torchrun --nproc-per-node N tools/e2e_scaling_probe.py [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

rank, local, world = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("gloo")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 600
ring = bench.build_ring(dev, rank, 32, "map")
for p in ring:
    hb = p._host_buffers()
    hb["x_d1"].copy_(p.host_inputs[0])
    for s, t in zip(p.scales, p.host_inputs[1]):
        hb["src"][s].copy_(t)
streams = [torch.cuda.Stream() for _ in range(32)]


def capture(plan, h2d, kern, d2h):
    hb = plan._host_buffers()

    def enqueue():
        if h2d:
            plan._in_dev.copy_(hb["packed"], non_blocking=True)
        if kern:
            plan.run()
        if d2h:
            hb["depth"].copy_(plan.depth, non_blocking=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        enqueue()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        enqueue()
    return g


def run(tag, h2d, kern, d2h, lanes, plans=32):
    graphs = [capture(p, h2d, kern, d2h) for p in ring[:plans]]
    ss = streams[:lanes]

    def go(k):
        for i in range(k):
            with torch.cuda.stream(ss[i % lanes]):
                graphs[i % len(graphs)].replay()
    go(64)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    go(K)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, t_issue], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"case": tag, "ranks": world, "lanes": lanes, "plans": plans, "maps_per_s": round(world * K * 16 / float(t[0])),
                          "us_per_step_per_rank": round(float(t[0]) / K * 1e6, 1), "host_issue_us_per_step": round(float(t[1]) / K * 1e6, 1)}), flush=True)
    del graphs


run("full e2e", True, True, True, 32)
run("full e2e", True, True, True, 8)
run("full e2e", True, True, True, 2)
run("full e2e", True, True, True, 8, plans=8)      # 8 pinned result buffers per rank (16 MB) instead of 32 (64 MB)
run("full e2e", True, True, True, 4, plans=4)
run("D2H only", False, False, True, 4, plans=4)
run("D2H only", False, False, True, 32)
run("H2D + D2H", True, False, True, 32)
run("kernels only", False, True, False, 32)
run("H2D + kernels", True, True, False, 32)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
