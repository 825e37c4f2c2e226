#!/usr/bin/env python
"""Run one pass over a ring of resident batches on several stream branches with the RDM_TIMING build
(per-CTA clock64 printouts) to see how co-resident CTAs affect each other."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
import torch  # noqa: E402

import bench  # noqa: E402
from md_rdm_b200.fusion import capture_ring  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
ring = bench.build_ring(dev, 0, 8, "raw")
g = capture_ring(ring, n_streams)
torch.cuda.synchronize()
print("=== measured pass", flush=True)
g.replay()
torch.cuda.synchronize()
