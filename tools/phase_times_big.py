#!/usr/bin/env python
"""Per-kernel durations with several reference batches (groups of 16) per launch: the kernels' throughput
when the chip is full.   python tools/phase_times_big.py [groups] [raw|map]"""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops  # noqa: F401,E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 8
source = sys.argv[2] if len(sys.argv) > 2 else "raw"
dev = torch.device("cuda:0")
R = torch.ops.rdm
N = 16 * groups
ring = []
for b in range(max(64 // groups, 4)):
    x_d1, rel, weights = bench.synthetic_batch(N, bench.SCALES, seed=1234 + b)
    plan = FusionPlan(N, bench.SCALES, source, group=16, device=dev, want_bins=True, flags=bench.PLAN_FLAGS)
    rel_d = [r.to(dev) for r in rel]
    srcs = [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel_d] if source == "raw" else rel_d
    plan.load_inputs(x_d1.to(dev), srcs, torch.cat([w.reshape(-1) for w in weights]).to(dev))
    ring.append(plan)
out = {"source": source, "groups_per_launch": groups, "ring": len(ring)}
for name, mask in FusionPlan.phase_masks().items():
    out[name + "_us_per_batch16"] = round(bench.time_serial([(lambda p=p, m=mask: p.run_als_phase(m)) for p in ring], 100) * 1e6 / groups, 2)
out["tail_us_per_batch16"] = round(bench.time_serial([(lambda p=p: p.run_tail()) for p in ring], 100) * 1e6 / groups, 2)
out["step_us_per_batch16"] = round(bench.time_serial([(lambda p=p: p.run()) for p in ring], 100) * 1e6 / groups, 2)
print(json.dumps(out))
