#!/usr/bin/env python
"""Minimal driver for ncu: a few fusion steps on one resident batch (run under gpurun).

    python tools/profile_als.py [steps]
"""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
ring = bench.build_ring(dev, 0, 2, "raw", bench.BATCH)
for i in range(steps):
    ring[i % 2].run()
torch.cuda.synchronize()
print("done", float(ring[0].depth.sum()))
