#!/usr/bin/env python
"""Warm, graph-timed durations of the training step's small ops at batch 16 (one stream, 8 calls per graph)."""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops as ops  # noqa: E402
from md_rdm_b200.training import TrainingStep  # noqa: E402

B = 16
dev = torch.device("cuda:0")
R = torch.ops.rdm
ts = TrainingStep(B, bench.SCALES, device=dev)
_, rel, weights = bench.synthetic_batch(B, bench.SCALES, seed=1603)
y_raw, logits = bench.synthetic_gt(B, 1604)
ts.load(rel, y_raw, logits, torch.cat([w.reshape(-1) for w in weights]))
ts.step()
torch.cuda.synchronize()
with torch.no_grad():
    x_d1, ord_ = R.dorn_regression(ts.logits)
    relm = ts.plan.run_als()
    final, yhat, A = R.fuse_tail(x_d1, [relm[s] for s in ts.scales], ts.weights.detach(), True)
    y, pyr, ord_t = R.gt_prepare(ts.y_raw)
    g = torch.randn_like(final)
    cases = {
        "dorn_regression": lambda: R.dorn_regression(ts.logits),
        "gt_prepare": lambda: R.gt_prepare(ts.y_raw),
        "component_loss": lambda: R.component_loss(yhat, pyr, ts.plan.kmax),
        "ordinal_loss": lambda: R.ordinal_loss(ord_, ord_t),
        "mse_loss(torch)": lambda: torch.nn.functional.mse_loss(final, y),
        "fuse_tail(want_A)": lambda: R.fuse_tail(x_d1, [relm[s] for s in ts.scales], ts.weights.detach(), True),
        "fuse_tail_bwd": lambda: R.fuse_tail_bwd(g, list(A)),
        "run_als(overlap)": lambda: ts.plan.run_als(),
    }
    out = {}
    for name, fn in cases.items():
        out[name] = round(bench.time_serial([fn] * 8, 64) * 1e6, 2)
print(json.dumps(out))
