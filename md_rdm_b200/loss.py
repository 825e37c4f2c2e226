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
    """utils.py:195-211 in one elementwise kernel (f32 scalars K, alpha, beta as in the reference; an f64 map is
    processed in f64, an f32 map in f32).  Other dtypes are widened to f64 first."""
    if depth.dtype not in (torch.float32, torch.float64):
        depth = depth.double()
    return torch.ops.rdm.depth2label_sid(depth, float(K), float(alpha), float(beta))
