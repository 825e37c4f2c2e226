#!/usr/bin/env python
"""GPU check: one FusionPlan carrying many reference batches per launch (n_images = 16 x groups, group = 16) must
equal separate batch-16 plans bit for bit (every arg-min group is independent), and hold no NaN.
    python tools/check_grouped.py [groups] [raw|map]"""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
import md_rdm_b200.ops  # noqa: F401,E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 29
source = sys.argv[2] if len(sys.argv) > 2 else "raw"
dev = torch.device("cuda:0")
R = torch.ops.rdm
N = 16 * groups
x_d1, rel, weights = bench.synthetic_batch(N, bench.SCALES, seed=1234)
w = torch.cat([t.reshape(-1) for t in weights]).to(dev)


def run(n, xs, rs):
    plan = FusionPlan(n, bench.SCALES, source, group=16, device=dev, want_bins=False)
    rel_d = [r.to(dev) for r in rs]
    srcs = [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel_d] if source == "raw" else rel_d
    plan.load_inputs(xs.to(dev), srcs, w)
    plan.run()
    torch.cuda.synchronize()
    return plan


big = run(N, x_d1, rel)
print("NaN in depth:", int(torch.isnan(big.depth).sum()), "in yhat:", int(torch.isnan(big.yhat).sum()),
      {s: int(torch.isnan(big.rel[s]).sum()) for s in bench.SCALES})
bad = 0
for g in range(groups):
    sl = slice(16 * g, 16 * g + 16)
    small = run(16, x_d1[sl], [r[sl] for r in rel])
    for s in bench.SCALES:
        if not torch.equal(small.rel[s], big.rel[s][sl]) or not torch.equal(small.kstar[s].view(-1), big.kstar[s][g].view(-1)):
            bad += 1
            print("group", g, "scale", s, "differs: k*", small.kstar[s].view(-1).tolist(), big.kstar[s][g].view(-1).tolist(),
                  float((small.rel[s] - big.rel[s][sl]).abs().max()))
    if not torch.equal(torch.nan_to_num(small.depth, nan=-7.0), torch.nan_to_num(big.depth[sl], nan=-7.0)):
        bad += 1
        print("group", g, "depth differs", float((small.depth - big.depth[sl]).abs().max()), "nan small", int(torch.isnan(small.depth).sum()))
print("groups", groups, "mismatches", bad)
sys.exit(1 if bad else 0)
