#!/usr/bin/env python
"""Small end-to-end case touching every kernel once (for compute-sanitizer memcheck under gpurun)."""
import math
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import md_rdm_b200.computations as cp  # noqa: E402
import md_rdm_b200.ops  # noqa: E402,F401
from md_rdm_b200 import _cabi  # noqa: E402
from md_rdm_b200.codebooks import default_quantization  # noqa: E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402
from md_rdm_b200.ops import fuse_tail_autograd  # noqa: E402

R = torch.ops.rdm
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
q = default_quantization()
B = 3
for source in ("map", "raw"):
    scales = (8, 16, 32, 64)
    plan = FusionPlan(B, scales, source, device=dev, want_bins=True, want_values=True, want_A=True)
    rel = [torch.exp(0.3 * torch.randn(B, 1, s, s, generator=g)).to(dev) for s in scales]
    srcs = rel if source == "map" else [R.pair_v1(r) if r.shape[2] == 8 else R.pair_id(r)[0] for r in rel]
    plan.load_inputs(torch.randint(1, 90, (B, 1, 8, 8), generator=g).to(dev), srcs, torch.rand(plan.n_weights).to(dev))
    plan.run()
    plan.run_host(plan.x_d1.cpu(), [s.cpu() for s in srcs])
# replay path (k* >= 2), 64-row units
u = torch.exp(0.8 * torch.randn(4, 64, 1, generator=g))
Rq = (u * u.transpose(1, 2)).float().to(dev)
R.als_rank1(Rq, _cabi.SRC_VAL_F32, 64, 8, 30, 4, None, None, False, False)
# stand-alone ops incl. the GT path and every backward
y = (0.5 + 9.5 * torch.rand(B, 1, 226, 226, generator=g, dtype=torch.float64)).to(dev)
y128 = cp.resize(y, 128).requires_grad_(True)
comps = cp.decompose_depth_map([], R.gm_normalize(y128), 7)[::-1]
logs = [cp.make_matrix([c], True).view(c.shape) for c in comps]
w = [torch.rand(1, 1, device=dev, requires_grad=True) for _ in comps]
pred = cp.make_pred(w, [l.view(B, 1, -1) for l in logs], True, False)
out = cp.recombination(pred)
(out ** 2).mean().backward()
thr, lvl = q.device_tables(16, dev)
R.lloyd_quantize(torch.rand(1001, dtype=torch.float64, device=dev) + 0.5, thr, lvl)
R.lloyd_quantize(torch.rand(1003, device=dev) + 0.5, thr, lvl)
cp.als_step(torch.rand(2, 256, 64, device=dev), torch.rand(2, 64, 1, device=dev), True)
cp.multi_upsample(torch.rand(2, 1, 4, 4, device=dev), 3)
x_d1 = torch.randint(1, 90, (B, 1, 8, 8), generator=g).to(dev)
wflat = torch.rand(4 + 3 + 4, device=dev, requires_grad=True)
d, _ = fuse_tail_autograd(x_d1, [torch.rand(B, 1, 8, 8, device=dev) + 0.5, torch.rand(B, 1, 16, 16, device=dev) + 0.5], wflat)
(d ** 2).mean().backward()
torch.cuda.synchronize()
print("sanitize case done", float(out.sum()), float(wflat.grad.sum()))
