#!/usr/bin/env python
"""Throughput of the fusion path against the number of reference batches (groups of 16 images) per launch.

    python tools/launch_rate_probe.py [lanes]

If maps/s grows with the images per launch while the work per image is unchanged, the single-batch
configuration is bound by the kernel launch rate of the device front end, not by the SMs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops  # noqa: F401,E402
from md_rdm_b200.fusion import FusionPlan, capture_lane  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
R = torch.ops.rdm
for groups in (1, 2, 4, 8):
    N = 16 * groups
    plans = []
    for b in range(lanes * 2):
        x_d1, rel, weights = bench.synthetic_batch(N, bench.SCALES, seed=1234 + b)
        plan = FusionPlan(N, bench.SCALES, "raw", group=16, device=dev, want_bins=True)
        rel_d = [r.to(dev) for r in rel]
        srcs = [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel_d]
        plan.load_inputs(x_d1.to(dev), srcs, torch.cat([w.reshape(-1) for w in weights]).to(dev))
        plans.append(plan)
    graphs = [capture_lane(plans[j::lanes]) for j in range(lanes)]
    def run(reps):
        for _ in range(reps):
            for g, st in graphs:
                with torch.cuda.stream(st):
                    g.replay()
    run(3)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    s0.record(cur)
    for g, st in graphs:
        st.wait_event(s0)
    reps = 20
    run(reps)
    for g, st in graphs:
        ev = torch.cuda.Event()
        ev.record(st)
        cur.wait_event(ev)
    s1.record(cur)
    torch.cuda.synchronize()
    ms = s0.elapsed_time(s1)
    maps = reps * len(plans) * N
    print(json.dumps({"groups_per_launch": groups, "lanes": lanes, "maps_per_s": round(maps / ms * 1e3), "us_per_batch16": round(ms * 1e3 / (maps / 16), 2)}), flush=True)
    del graphs, plans
    torch.cuda.empty_cache()
