#!/usr/bin/env python
"""Golden vectors for the SURVEY 8f "next" row (DORN head + ordinal loss) from the UNMODIFIED reference:
`Ordinal_Layer.DornOrdinalRegression` (network/RDM_Net.py:313-345), `loss.Ordinal_Loss.calc` (loss.py:17-59)
with autograd gradients.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
from oracle import fusion_ref as fr  # noqa: E402


class _Q:   # Ordinal_Layer only stores the quantizer; the DORN branch never touches it
    pass


def main():
    import loss as ref_loss
    import network.RDM_Net as rn
    rn.use_cuda = False
    g = torch.Generator().manual_seed(808)
    N, K, H, W = 3, 90, 8, 8
    x = 2.0 * torch.randn(N, 2 * K, H, W, generator=g)
    x[0, :8] = torch.tensor([-1.0, 0.0, 1e-9, 2e4, 5.0, 5.0, 1e-8, 1e4]).view(8, 1, 1)     # clamp edges, ties
    x = x.requires_grad_(True)
    layer = rn.Ordinal_Layer(1, True, _Q())
    decode, ord_c1 = layer(x)
    depth = 0.5 + 9.5 * torch.rand(N, 1, H, W, generator=g, dtype=torch.float64)
    target = fr.depth2label_sid(depth)
    lossv = ref_loss.Ordinal_Loss().calc(ord_c1, target, cuda=False)
    lossv.backward()
    d2, o2 = fr.dorn_regression(x.detach())
    assert torch.equal(d2, decode) and torch.equal(o2, ord_c1.detach())
    l2 = fr.ordinal_loss(o2, target)
    print("oracle vs reference: decode equal, ord bit-equal, loss", float(lossv), float(l2))
    assert abs(float(lossv) - float(l2)) <= 1e-6 * abs(float(lossv))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "dorn_loss.npz"), x=x.detach().numpy(), decode=decode.numpy(),
                        ord=ord_c1.detach().numpy(), target=target.numpy(), loss=np.array(float(lossv), dtype=np.float32),
                        grad_x=x.grad.numpy())


if __name__ == "__main__":
    main()
