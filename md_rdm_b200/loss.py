"""Drop-in for `Ordinal_Loss` of the reference's `loss.py` (loss.py:8-59) - SURVEY 8f "next" row: the
other loss of the training step, fused into one reduction kernel (forward) and one elementwise kernel
(backward) instead of a Python loop over the K ordinal channels, an (N,K,H,W) index tensor and two
boolean masks."""
from __future__ import annotations

import torch

from . import ops  # noqa: F401  (registers torch.ops.rdm.*)


class Ordinal_Loss():
    def __init__(self):
        self.loss = 0.0

    def calc(self, ord_labels, target, cuda):
        """ord_labels (N,K,H,W) f64 ordinal probabilities, target (N,1,H,W) integer SID labels."""
        self.loss = torch.ops.rdm.ordinal_loss(ord_labels.double(), target)
        return self.loss


def depth2label_sid(depth, K=90.0, alpha=0.02, beta=10.0, cuda=True):
    """utils.py:195-211 (caller glue; elementwise torch on the tensor's device, f32 scalars as in the reference)."""
    dev = depth.device
    a, b, k = torch.tensor(alpha, device=dev), torch.tensor(beta, device=dev), torch.tensor(K, device=dev)
    label = k * torch.log(depth / a) / torch.log(b / a)
    return torch.max(label, torch.zeros(label.shape, device=dev)).int()
