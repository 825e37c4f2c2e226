#!/usr/bin/env python
"""Do the four kernels of the path overlap when they run in different lanes?  For every kernel alone and for every
pair: L lanes (streams), each replaying a CUDA graph of one kernel over its own 4 resident batches; the time per batch
of a pair run together against the sum / max of the two alone.   python tools/overlap_matrix.py [lanes_per_kernel]"""
import json
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
from md_rdm_b200.fusion import FusionPlan  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
ring = bench.build_ring(dev, 0, 8 * L, "raw", bench.BATCH)
for p in ring:
    p.run()
torch.cuda.synchronize()
masks = dict(FusionPlan.phase_masks())
kinds = list(masks) + ["tail"]


def lane_graph(kind, plans, stream):
    def body():
        for p in plans:
            p.run_tail() if kind == "tail" else p.run_als_phase(masks[kind])
    with torch.cuda.stream(stream):
        body()
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        body()
    return g


def run(kinds_now, reps=12):
    lanes = []
    for i, kind in enumerate(kinds_now):
        for l in range(L):
            plans = ring[(i * L + l) * 4:(i * L + l) * 4 + 4]
            s = torch.cuda.Stream()
            lanes.append((s, lane_graph(kind, plans, s)))
    cur = torch.cuda.current_stream()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for timed in (False, True):
        torch.cuda.synchronize()
        if timed:
            start.record(cur)
        for s, g in lanes:
            s.wait_stream(cur)
        for _ in range(reps):
            for s, g in lanes:
                with torch.cuda.stream(s):
                    g.replay()
        for s, g in lanes:
            cur.wait_stream(s)
        if timed:
            end.record(cur)
    torch.cuda.synchronize()
    return start.elapsed_time(end) * 1e3 / (reps * L * 4)   # us per batch of EACH kind


if os.environ.get("OVERLAP_ONLY"):   # quick form: the page ALS alone, with the sparsify kernel, and all four
    a = run(["als_sparse"])
    b = run(["als_sparsify", "als_sparse"])
    c = run(kinds)
    print(json.dumps({"lanes_per_kernel": L, "als_sparse": round(a, 2), "als_sparsify+als_sparse": round(b, 2), "all_four": round(c, 2)}))
    sys.exit(0)
alone = {k: run([k]) for k in kinds}
out = {"lanes_per_kernel": L, "alone_us_per_batch": {k: round(v, 2) for k, v in alone.items()}, "pairs": {}}
for i, a in enumerate(kinds):
    for b in kinds[i + 1:]:
        t = run([a, b])
        out["pairs"][f"{a}+{b}"] = {"together": round(t, 2), "sum": round(alone[a] + alone[b], 2), "max": round(max(alone[a], alone[b]), 2)}
t = run(kinds)
out["all_four"] = {"together": round(t, 2), "sum": round(sum(alone.values()), 2)}
print(json.dumps(out, indent=1))
