"""Drop-ins for the fusion-path classes of the reference's `network/RDM_Net.py` (RN):
`Ordinal_Layer` (RN:237-396), `Quantization` (RN:397-442), `Weights` (RN:443-491).

Same constructors, method names and return conventions; the arithmetic runs in the sm_100a
kernels behind `torch.ops.rdm.*`, including the DORN branch of `Ordinal_Layer` (RN:313-345), the
producer of the ordinary map (SURVEY 8f "next" row; `md_rdm_b200/loss.py` has the matching loss).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from . import computations as cp
from .codebooks import Quantization  # noqa: F401  (re-exported: same name as the reference class)

R = torch.ops.rdm

LIMIT_8 = 30      # CP:38 default, RN:364 passes none
LIMIT_PAGE = 100  # RN:378, RN:392


class Ordinal_Layer(nn.Module):
    def __init__(self, decoder_id, DORN, quantizer, flags: int = 0):
        """`flags` (not in the reference): `_cabi.ALS_TRUE_TRANSPOSE | ALS_TRUE_GM | ALS_CORRECT_TILING`, the
        "paper-correct" knobs of SURVEY 8f rank 4; 0 = the reference's behaviour."""
        super().__init__()
        self.quant = quantizer
        self.id = decoder_id - 3
        self.dorn = DORN
        self.flags = int(flags)

    # ------------------------------------------------------------------ codebooks
    def _tables(self, id, device):
        """Device f64 (thresholds[40], levels[41]) for scale id 3..7 from whatever quantizer
        object the caller supplied (the reference's attribute surface is all that is used)."""
        if hasattr(self.quant, "device_tables"):
            return self.quant.device_tables(1 << id, device)
        if id == 3:
            q, inv = self.quant.depth_ratio_008_008_quant, self.quant.depth_ratio_008_008_quant_inv
        else:
            q, inv = self.quant.get_with_id(id)
        cache = self.__dict__.setdefault("_tab_cache", {})
        key = (id, str(device))
        if key not in cache:
            cache[key] = (torch.as_tensor(q, dtype=torch.float64).reshape(-1).to(device),
                          torch.as_tensor(inv, dtype=torch.float64).reshape(-1).to(device))
        return cache[key]

    # ------------------------------------------------------------------ RN:244-257
    def sparse_comparison_v1(self, d_3):
        B, C, H, W = d_3.size()
        sparse_m = R.pair_v1(d_3.float())
        depth_labels = torch.empty(B, H * W, H * W, 0)        # shape carrier only (RN:254)
        return self.LloydQuantization(depth_labels, sparse_m)

    # ------------------------------------------------------------------ RN:259-284
    def sparse_comparison_id(self, dn, dn_1):
        B, C, H, W = dn.size()
        sparse_m = R.pair_pages(dn, dn_1)
        depth_labels = torch.empty(B, H * W, (H // 2) * (W // 2), 0)
        return self.LloydQuantization(depth_labels, sparse_m, id=self.id)

    # ------------------------------------------------------------------ RN:286-311
    def LloydQuantization(self, labels, relative_depths, id=3):
        """bin = #{i: x >= q_i} in x's dtype, x <- inv[bin]; quantises `relative_depths` in place
        (the reference writes through `torch.flatten`, a view for contiguous input) and returns
        it viewed as labels.shape[:3].  `labels` is only a shape carrier here."""
        N, C, W = labels.shape[0], labels.shape[1], labels.shape[2]
        thr, lvl = self._tables(id, relative_depths.device)
        values, _ = R.lloyd_quantize(relative_depths, thr, lvl)
        if relative_depths.is_contiguous():
            relative_depths.copy_(values)
            return relative_depths.view(N, C, W)
        return values.view(N, C, W)

    # ------------------------------------------------------------------ RN:313-345 (SURVEY 8f "next" row)
    def DornOrdinalRegression(self, x):
        """(decode_c (N,1,H,W) int64, ord_c1 (N,K,H,W) f64): one kernel instead of two clones, cat,
        clamp, double, softmax, clone, compare, sum."""
        return R.dorn_regression(x.float())

    # ------------------------------------------------------------------ RN:347-396
    def forward(self, x):
        if self.dorn:
            return self.DornOrdinalRegression(x)
        B, C, side, _ = x.size()
        if side != (1 << self.id):
            raise RuntimeError(f"Ordinal_Layer id {self.id + 3} expects {1 << self.id}x{1 << self.id} maps, got {side}")
        thr, lvl = self._tables(self.id, x.device)
        # pair build + Lloyd + ALS + normalisation (+ page split / re-tiling for id > 4) in one
        # fused call; the pair matrix never exists in HBM.  One arg-min group per call (CP:172-173).
        rows, limit = (64, LIMIT_8) if self.id == 3 else (256, LIMIT_PAGE)
        return R.als_rank1(x.float(), _cabi.SRC_MAP_F32, rows, side, limit, B, thr, lvl, False, False, self.flags)[0]


class Weights(nn.Module):
    """RN:443-491: one non-negative weight vector per fine-detail slot (d0, f1..f7)."""

    def __init__(self, vector_sizes, use_cuda, relative_only):
        super().__init__()
        self.use_cuda = use_cuda
        self.relative_only = relative_only
        dev = "cuda" if use_cuda else "cpu"
        names = ["d0", "f1", "f2", "f3", "f4", "f5", "f6", "f7"]
        for name, size in zip(names, vector_sizes):
            setattr(self, name, nn.Parameter(torch.abs(torch.randn((size, 1))).to(dev)))
        self.weight_list = [getattr(self, n) for n in names]
        for weight_vector in self.weight_list:
            if weight_vector.shape[0] == 0:
                weight_vector.requires_grad = False

    def update(self, weight_index, lr, gradient):
        self.weight_list[weight_index] = self.weight_list[weight_index] - lr * gradient

    def get(self, index):
        return self.weight_list[index]

    def flat(self):
        """All weights concatenated in slot order (what the fused tail kernel takes)."""
        return torch.cat([w.reshape(-1) for w in self.weight_list if w.numel()])

    def forward(self, x):
        return cp.make_pred(self.weight_list, x, self.use_cuda, self.relative_only)
