#!/usr/bin/env python
"""Per-launch durations of the fusion path's kernels, one stream, CUDA-graph timed (run under gpurun).

    python tools/phase_times.py [raw|map] [ring]

Phase bits of rdm_als_fused_phases: 4 = compact page form, 8 = ALS on compact pages, 16 = dense ALS,
2 = select; then the tail kernel and the whole step.
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
ring = bench.build_ring(dev, 0, n, source)
out = {"source": source, "ring": n}
for name, mask in (("sparsify", 4), ("als_sparse", 8), ("als_dense", 16), ("iterate_all", 1), ("select", 2)):
    out[name + "_us"] = round(bench.time_serial([(lambda p=p, m=mask: p.run_als_phase(m)) for p in ring], 400) * 1e6, 2)
out["tail_us"] = round(bench.time_serial([(lambda p=p: p.run_tail()) for p in ring], 400) * 1e6, 2)
out["step_us"] = round(bench.time_serial([(lambda p=p: p.run()) for p in ring], 400) * 1e6, 2)
print(json.dumps(out))
