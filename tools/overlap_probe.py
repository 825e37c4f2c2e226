import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench
from md_rdm_b200.fusion import FusionPlan
dev = torch.device("cuda:0")
ring = bench.build_ring(dev, 0, 8, "raw", bench.BATCH)
for ov in (False, True):
    t = bench.time_serial([(lambda p=p, ov=ov: p.run(overlap=ov)) for p in ring], 200)
    print("overlap", ov, round(t * 1e6, 2), "us per call")
