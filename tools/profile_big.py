#!/usr/bin/env python
"""ncu driver: one FusionPlan with several reference batches per launch, a few steps.
    python tools/profile_big.py [groups] [steps]"""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops  # noqa: F401,E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
R = torch.ops.rdm
N = 16 * groups
x_d1, rel, weights = bench.synthetic_batch(N, bench.SCALES, seed=1234)
plan = FusionPlan(N, bench.SCALES, "raw", group=16, device=dev, want_bins=True, flags=bench.PLAN_FLAGS)
rel_d = [r.to(dev) for r in rel]
srcs = [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel_d]
plan.load_inputs(x_d1.to(dev), srcs, torch.cat([w.reshape(-1) for w in weights]).to(dev))
for _ in range(steps):
    plan.run()
torch.cuda.synchronize()
print("done", float(plan.depth.sum()))
