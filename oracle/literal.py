"""Loop-for-loop CPU restatement of the reference's relative-decoder tail.

TEST / BASELINE INFRASTRUCTURE -- see `oracle/__init__.py`.  This module restates the two
stages whose cost in the reference is interpreted Python loops, with the SAME loop structure
and the same torch calls per iteration, so that timing it on the GPU box's host cores stands
in for timing the reference itself (which is Python and cannot travel to the GPU box):

* pair build  -- network/RDM_Net.py:259-284 (`sparse_comparison_id`): a double loop over the
  256 pixels of a page, each building a ones-like area with a 3x3 window copied in
  (network/computations.py:269-295) and one broadcast multiply;
* Lloyd       -- network/RDM_Net.py:286-311: 40 whole-tensor compares into a label tensor, a
  sum, and then ONE PYTHON ITERATION PER MATRIX ENTRY writing `inv[int(idx)]`.

ALS, decomposition, weighting and recombination are already vectorised torch in the reference
(network/computations.py:38-155, 368-421, 512-528), so the restatement in `fusion_ref` has
their cost profile; `relative_decoder_tail_literal` chains the two.

Parity status: PINNED -- tools/make_golden.py asserts these functions bit-equal to the
unmodified reference methods; tests/test_oracle_literal.py asserts them bit-equal to the
vectorised oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from . import fusion_ref as fr


def lloyd_literal(x: torch.Tensor, thresholds: torch.Tensor, levels: torch.Tensor) -> torch.Tensor:
    """RN:286-311.  `thresholds`/`levels` are the (40,)/(41,) f64 tables; python floats are
    compared against the tensor, so the compare runs in x's dtype exactly as in the reference."""
    q = [float(v) for v in thresholds]
    inv = [float(v) for v in levels]
    labels = torch.zeros(tuple(x.shape) + (40,))
    for i in range(40):
        labels[..., i] = (x >= q[i])
    indices = torch.flatten(torch.sum(labels, -1))
    flat = torch.flatten(x.clone())
    for i in range(flat.shape[0]):          # the reference's per-element loop (RN:296-297, RN:309-310)
        flat[i] = inv[int(indices[i])]
    return flat.view(x.shape)


def pair_v1_literal(d3: torch.Tensor) -> torch.Tensor:
    """RN:244-252 (already two whole-tensor ops in the reference)."""
    B, C, H, W = d3.size()
    flat = d3.view(B, C, H * W)
    return torch.matmul(flat.view(B, H * W, C), torch.pow(flat, -1)).view(B, H * W, H * W)


def pair_id_literal(dn: torch.Tensor, dn_1: torch.Tensor) -> torch.Tensor:
    """RN:259-280 + CP:269-295: one (B,1,64) row per page pixel, concatenated."""
    B, C, H, W = dn.size()
    page = dn.view(B, H, W)
    rows = []
    for r in range(H):
        for c in range(W):
            r0 = int(min(max(np.floor(r / 2), 0), dn_1.shape[2] - 3))
            c0 = int(min(max(np.floor(c / 2), 0), dn_1.shape[3] - 3))
            area = torch.ones_like(dn_1)
            for rr in (r0, r0 + 1, r0 + 2):
                area[:, :, rr, c0:c0 + 3] = dn_1[:, :, rr, c0:c0 + 3]
            area = area.view(B, 1, dn_1.shape[2] * dn_1.shape[3])
            rows.append(page[:, r, c].view(B, 1, 1) * torch.pow(area, -1))
    return torch.cat(rows, 1)


def relative_decoder_tail_literal(x: torch.Tensor, books) -> torch.Tensor:
    """RN:358-396 (non-DORN `Ordinal_Layer.forward`) with the literal pair build and Lloyd."""
    s = x.shape[2]
    q, lv = books[s]
    if s == 8:
        vals = lloyd_literal(pair_v1_literal(x), q, lv)
        return fr.als_rank1(vals, fr.LIMIT_8)[0]
    dn_1 = fr.resize(x, s // 2)
    outs = []
    for page, parent in fr.split_pages(x, dn_1):
        vals = lloyd_literal(pair_id_literal(page, parent), q, lv)
        outs.append(fr.als_rank1(vals, fr.LIMIT_PAGE)[0])
    return outs[0] if s == 16 else fr.retile_pages(outs)


def fusion_forward_literal(x_d1, rel_maps, weights, books):
    """The whole path (RN:103-133 + network/module.py:132) with the literal decoder tails."""
    import math
    filled = [relative_decoder_tail_literal(x, books) for x in rel_maps]
    rows = [fr.decompose(fr.gm_normalize(x_d1), 3)]
    for f in filled:
        rows.append(fr.decompose(f, int(math.log2(f.shape[2])), relative_map=True))
    y_hat = fr.make_pred(weights, fr.fine_detail_matrices(rows))
    return fr.recombination(y_hat)
