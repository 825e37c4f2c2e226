#!/usr/bin/env python
"""Per-stage kernel timings (stand-alone ops) on one GPU: achieved algorithmic GB/s for the
streaming kernels (pair build, Lloyd, decomposition, recombination, fused tail) at batch 16.

    python tools/stage_bench.py            (under gpurun)
"""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import md_rdm_b200.ops  # noqa: E402,F401
from md_rdm_b200.codebooks import default_quantization  # noqa: E402

R = torch.ops.rdm
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)


def timeit(fn, reps=200, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()          # evict L2 (buffer larger than L2) so inputs come from HBM
        s.record()
        fn()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps * 1e-3


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
q = default_quantization()
out = {}
x8 = torch.exp(0.3 * torch.randn(B, 1, 8, 8, generator=g)).to(dev)
t = timeit(lambda: R.pair_v1(x8), flush=flush)
out["pair_v1"] = dict(us=t * 1e6, gbs=(B * (256 + 16384)) / t / 1e9)
for s in (16, 32, 64, 128):
    x = torch.exp(0.3 * torch.randn(B, 1, s, s, generator=g)).to(dev)
    P = (s // 16) ** 2
    t = timeit(lambda: R.pair_id(x), flush=flush)
    out[f"pair_id_{s}"] = dict(us=t * 1e6, gbs=B * (4 * s * s + 8 * (s // 2) ** 2 + P * 256 * 64 * 8) / t / 1e9)
    raw, _ = R.pair_id(x)
    thr, lvl = q.device_tables(s, dev)
    t = timeit(lambda: R.lloyd_quantize(raw, thr, lvl), flush=flush)
    out[f"lloyd_f64_{s}"] = dict(us=t * 1e6, gbs=raw.numel() * (8 + 8 + 1) / t / 1e9)
raw8 = R.pair_v1(x8)
thr, lvl = q.device_tables(8, dev)
t = timeit(lambda: R.lloyd_quantize(raw8, thr, lvl), flush=flush)
out["lloyd_f32_8"] = dict(us=t * 1e6, gbs=raw8.numel() * 9 / t / 1e9)
y = (0.5 + 9.5 * torch.rand(B, 1, 128, 128, generator=g, dtype=torch.float64)).to(dev)
t = timeit(lambda: R.decompose(R.gm_normalize(y), False), flush=flush)
out["gt_normalize+decompose_128"] = dict(us=t * 1e6, gbs=B * (131072 * 3 + 174760) / t / 1e9)
comps = [torch.randn(B, 1, 2 ** k, 2 ** k, generator=g).to(dev) for k in range(8)]
t = timeit(lambda: R.recombination(comps, 7), flush=flush)
out["recombination_128"] = dict(us=t * 1e6, gbs=B * (131072 + 4 * 21845) / t / 1e9)
# SURVEY 8f "next" row: DORN head + ordinal loss (K = 90 ordinal bins on the 8x8 map)
xh = torch.randn(B, 180, 8, 8, generator=g).to(dev)
t = timeit(lambda: R.dorn_regression(xh), flush=flush)
out["dorn_regression"] = dict(us=t * 1e6, gbs=B * (180 * 64 * 4 + 90 * 64 * 8 + 64 * 8) / t / 1e9)
_, ordp = R.dorn_regression(xh)
tgt = torch.randint(0, 90, (B, 1, 8, 8), generator=g).to(dev)
t = timeit(lambda: R.ordinal_loss(ordp, tgt), flush=flush)
out["ordinal_loss"] = dict(us=t * 1e6, gbs=B * (90 * 64 * 8 + 64 * 4) / t / 1e9)
print(json.dumps(out, indent=1))
