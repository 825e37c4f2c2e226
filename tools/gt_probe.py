import sys; sys.path.insert(0, "/root/repo")
import torch, bench
import md_rdm_b200.ops
R = torch.ops.rdm
y, _ = bench.synthetic_gt(16, 5)
y = y.cuda()
for _ in range(3):
    out = R.gt_prepare(y)
torch.cuda.synchronize()
