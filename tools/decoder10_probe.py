#!/usr/bin/env python
"""Per-call time of a plan that includes decoder 10 (128x128, RN:61): shared ALS launches + the composed tail."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402

dev = torch.device("cuda:0")
for scales in ((8, 16, 32, 64), (8, 16, 32, 64, 128)):
    x_d1, rel, weights = bench.synthetic_batch(16, scales, seed=7)
    plans = []
    for i in range(2):
        p = FusionPlan(16, scales, "map", device=dev, want_bins=False)
        p.load_inputs(x_d1.to(dev), [r.to(dev) for r in rel], torch.cat([w.reshape(-1) for w in weights]).to(dev))
        plans.append(p)
    t_all = bench.time_serial([(lambda p=p: p.run()) for p in plans], 16)
    t_als = bench.time_serial([(lambda p=p: p.run_als()) for p in plans], 16)
    print(scales, "per call %.1f us, of which ALS launches %.1f us, tail %.1f us (%s)" % (
        t_all * 1e6, t_als * 1e6, (t_all - t_als) * 1e6, "composed from the stand-alone kernels" if plans[0].composed_tail else "one fused launch"))
