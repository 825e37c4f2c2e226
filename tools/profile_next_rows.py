#!/usr/bin/env python
"""ncu driver for the SURVEY 8f rows beside the fusion path: ground-truth preparation, conv head fused with the pair build,
the training step's fused pieces (tail backward, component loss), the DORN head / ordinal loss / SID labels.
    python tools/profile_next_rows.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops  # noqa: F401,E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402
from md_rdm_b200.training import TrainingStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
R = torch.ops.rdm
g = torch.Generator().manual_seed(3)
y226, logits = bench.synthetic_gt(B, 5)
y226 = y226.to(dev)
cplan = FusionPlan(B, (16, 32), "map", device=dev, want_bins=False)
cf = {16: torch.randn(B, 1664, 16, 16, generator=g).to(dev), 32: torch.randn(B, 832, 32, 32, generator=g).to(dev)}
cw = {16: torch.randn(1664, generator=g).to(dev) * 0.01, 32: torch.randn(832, generator=g).to(dev) * 0.01}
cb = {16: torch.ones(1).to(dev) * 2, 32: torch.ones(1).to(dev) * 2}
ts = TrainingStep(B, bench.SCALES, device=dev)
_, rel, weights = bench.synthetic_batch(B, bench.SCALES, seed=1603)
ts.load(rel, bench.synthetic_gt(B, 1604)[0], logits, torch.cat([w.reshape(-1) for w in weights]))
for _ in range(2):
    R.gt_prepare(y226)
    cplan.enqueue_conv_heads(cf, cw, cb)
    ts.step()
torch.cuda.synchronize()
print("done")
