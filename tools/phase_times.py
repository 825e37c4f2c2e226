#!/usr/bin/env python
"""Per-launch durations of the fusion path's kernels, one stream, CUDA-graph timed (run under gpurun).

    python tools/phase_times.py [raw|map] [ring]

The launches of rdm_als_fused_phases one by one (FusionPlan.phase_masks), then the tail kernel and the whole step.
"""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402

source = sys.argv[1] if len(sys.argv) > 1 else "raw"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
ring = bench.build_ring(dev, 0, n, source, bench.BATCH)
from md_rdm_b200.fusion import FusionPlan  # noqa: E402
out = {"source": source, "ring": n}
for name, mask in FusionPlan.phase_masks().items():
    out[name + "_us"] = round(bench.time_serial([(lambda p=p, m=mask: p.run_als_phase(m)) for p in ring], 400) * 1e6, 2)
out["tail_us"] = round(bench.time_serial([(lambda p=p: p.run_tail()) for p in ring], 400) * 1e6, 2)
out["step_us"] = round(bench.time_serial([(lambda p=p: p.run()) for p in ring], 400) * 1e6, 2)
print(json.dumps(out))
