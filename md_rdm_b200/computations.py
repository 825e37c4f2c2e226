"""Drop-in for the reference's `network/computations.py` ("cp") on the fusion path.

Same function names, positional order, defaults, return dtypes/shapes and list-mutation
side effects as the reference (az16/MD_RDM, CP = network/computations.py), so
`network/RDM_Net.py` and `network/module.py` can `import md_rdm_b200.computations as cp`.
Every arithmetic step runs in a hand-written sm_100a kernel of librdm_b200.so through
`torch.ops.rdm.*`; the `cuda` arguments are kept for signature compatibility - tensors must
already live on a CUDA device and there is no CPU path.

Only list plumbing that the reference itself does with views / `torch.cat` (split_matrix,
reconstruct, the per-scale MSE glue of optimize_components) stays in torch.
"""
from __future__ import annotations

import math

import torch

from . import _cabi, ops

__all__ = [
    "quadratic_als", "alternating_least_squares", "als_step", "matmul", "rmse", "split_matrix", "reconstruct", "quick_gm",
    "get_resized_area", "find_nans", "resize", "upsample", "multi_upsample", "decompose_depth_map", "recombination",
    "relative_fine_detail_matrix", "idx_from_size", "make_matrix", "make_pred", "optimize_components", "squared_err",
]

R = torch.ops.rdm


def _als(sparse_m, rows, side, limit):
    B, H, W = sparse_m.size()
    if (H, W) != (rows, 64):
        raise RuntimeError(f"ALS kernel supports ({rows},64) matrices here, got ({H},{W})")
    if sparse_m.dtype == torch.float32:
        kind = _cabi.SRC_VAL_F32
    elif sparse_m.dtype == torch.float64:
        kind = _cabi.SRC_VAL_F64          # `.float()` (CP:40 / CP:106) is applied on load
    else:
        sparse_m, kind = sparse_m.float(), _cabi.SRC_VAL_F32
    # the whole call is one arg-min group: rmse is a mean over the batch (CP:172-173)
    filled = R.als_rank1(sparse_m, kind, rows, side, limit, B, None, None, False, False)[0]
    return filled


def quadratic_als(sparse_m, cuda, n=3, limit=30, debug=False):
    """CP:38-85: (B,64,64) -> (B,1,8,8) f32."""
    if n != 3:
        raise RuntimeError("quadratic_als: only n=3 (64x64 matrices) exists on this path")
    return _als(sparse_m, 64, 8, limit)


def alternating_least_squares(sparse_m, n, cuda, limit=30, debug=False):
    """CP:95-155: (B,256,64) -> (B,1,16,16) f32."""
    if n != 4:
        raise RuntimeError("alternating_least_squares: only n=4 (256x64 page matrices) exists on this path")
    return _als(sparse_m, 256, 16, limit)


def matmul(t1, t2):
    return torch.matmul(t1, t2)


def rmse(m1, m2):
    return torch.mean((m1 - m2) ** 2) ** 0.5


def als_step(ratings, fixed_tensor, cuda, regularization_term=0.05):
    """CP:175-193 for a rank-1 factor (fixed_tensor (B,n,1)): (ratings @ f) / (f^T f + reg)."""
    if fixed_tensor.size(2) != 1:
        raise RuntimeError("als_step: only rank-1 factors (B,n,1) exist on this path")
    return R.als_step(ratings.float(), fixed_tensor.float(), float(regularization_term))


def split_matrix(d_n, d_n_1):
    """CP:201-216: row-major lists of 16x16 pages and their 8x8 parent pages (views)."""
    ratio = int(d_n.shape[2] / 16)
    first, second = [], []
    for i in range(ratio):
        for j in range(ratio):
            r_s, c_s = 16 * i, 16 * j
            first.append(d_n[:, :, r_s:r_s + 16, c_s:c_s + 16])
            second.append(d_n_1[:, :, r_s // 2:r_s // 2 + 8, c_s // 2:c_s // 2 + 8])
    return first, second


def reconstruct(splits, correct_tiling=False):
    """CP:218-238, as written: every block-column repeats the vertical stack of the first
    `ratio` pages (pages >= ratio never reach the output).  `correct_tiling=True` (not in the
    reference) puts page i*ratio+j at block (i, j), the inverse of split_matrix."""
    ratio = int(len(splits) ** (1 / 2))
    if correct_tiling:
        return torch.cat([torch.cat(splits[i * ratio:(i + 1) * ratio], 3) for i in range(ratio)], 2)
    rows = [torch.cat(splits[0:ratio], 2) for _ in range(ratio)]
    return torch.cat(rows, dim=3)


def quick_gm(t, rc):
    """CP:244-255 (int32 SID labels, MOD:126, are widened to int64: same values, same f32 result)."""
    if t.dtype in (torch.int32, torch.int16, torch.uint8):
        t = t.long()
    return R.quick_gm(t, int(rc))


def get_resized_area(r_s, r_e, c_s, c_e, dn_1):
    """CP:269-295: ones everywhere except rows r_s, r_s+1, r_e x columns c_s:c_e copied from
    dn_1, flattened to (B,1,H*W).  Pure indexing; the kernels never materialise it (the window
    test is `in_window` in csrc/rdm_common.cuh)."""
    B, C, H, W = dn_1.size()
    area = torch.ones_like(dn_1)
    for r in (r_s, r_s + 1, r_e):
        area[:, :, r, c_s:c_e] = dn_1[:, :, r, c_s:c_e]
    return area.view(B, 1, H * W)


def find_nans(container):
    for tensor in container:
        if torch.any(tensor.isnan()):
            return True
    return False


def resize(depth_map, newsize):
    """CP:308-311: `.double()` + bicubic(align_corners=False)."""
    if isinstance(newsize, (tuple, list)):
        oh, ow = int(newsize[0]), int(newsize[1])
    else:
        oh = ow = int(newsize)
    if depth_map.dtype not in (torch.float32, torch.float64):
        depth_map = depth_map.double()
    H, W = depth_map.shape[2], depth_map.shape[3]
    if H == W and oh == ow and 2 * oh == H:
        return R.resize_half(depth_map)
    return R.resize_bicubic(depth_map, oh, ow)


def upsample(depth_map):
    """CP:357-360."""
    return R.upsample_nearest(depth_map, 1)


def multi_upsample(depth_map, n):
    """CP:362-366 (n == 0 returns the input unchanged, dtype included)."""
    if n == 0:
        return depth_map
    elif n > 0:
        return R.upsample_nearest(depth_map, int(n))


def decompose_depth_map(container, dn, n, relative_map=False):
    """CP:368-392: appends [F_n, ..., F_1 (, D_0)] (f64) to `container` and returns it."""
    if n == 0:
        if not relative_map:
            container.append(dn)
        return container
    side = dn.shape[2]
    if side != 2 ** n or dn.shape[3] != side:
        raise RuntimeError(f"decompose_depth_map: a {side}x{dn.shape[3]} map cannot be decomposed into n={n} levels")
    if dn.dtype not in (torch.float32, torch.float64):
        dn = dn.double()
    B = dn.shape[0]
    packed = R.decompose(dn.reshape(B, 1, side, side), bool(relative_map))
    comps = ops.unpack_pyramid(packed, B, side, bool(relative_map))      # [D_0?, F_1 .. F_n]
    container.extend(comps[::-1])
    return container


def recombination(list_of_components, n=7):
    """CP:394-421.  Pops the first one/two entries of the caller's list like the reference."""
    comps = list(list_of_components)
    has_d0 = comps[0].shape[2] == 1
    for _ in range(2 if has_d0 else 1):
        list_of_components.pop(0)
    dt = torch.float64 if any(c.dtype == torch.float64 for c in comps) else torch.float32
    comps = [c if c.dtype == dt else c.to(dt) for c in comps]
    return R.recombination(comps, int(n))


def idx_from_size(fine_detail_map):
    return int(math.log2(fine_detail_map.size(2)))


def make_matrix(list_of_candidates, cuda):
    """CP:464-484: log + stack -> (B,K,M) f64."""
    return R.log_stack(list(list_of_candidates))


def relative_fine_detail_matrix(fine_detail_rows, cuda):
    """CP:423-443: bucket by side, empty slots dropped."""
    slots = [[] for _ in range(8)]
    for row in fine_detail_rows:
        for fine_detail_map in row:
            slots[idx_from_size(fine_detail_map)].append(fine_detail_map)
    return [make_matrix(x, cuda) for x in slots if not len(x) == 0]


def make_pred(w, A, cuda, relative_only):
    """CP:512-528.  Replaces the entries of `A` in place, like the reference."""
    weights = w[1::] if relative_only else w
    for i in range(len(A)):
        B, M = A[i].shape[0], A[i].shape[2]
        side = int(math.sqrt(M))
        A[i] = R.make_pred(A[i].double(), weights[i]).view(B, 1, side, side)
    return A


def squared_err(yhat, y, cuda):
    """CP:530-544 (caller glue: per-scale torch MSELoss)."""
    sqr_err_list = []
    if yhat[0].shape[2] > y[0].shape[2]:
        y.pop(0)
    for i in range(len(yhat)):
        sqr_err_list.append(torch.nn.MSELoss()(yhat[i], y[i]))
    return sqr_err_list


def optimize_components(yhat, y, cuda):
    """CP:499-510: returns (yhat, detached sum of the per-scale MSEs)."""
    pred = yhat
    loss = squared_err(pred, y, cuda)
    return pred, torch.sum(torch.as_tensor(loss))
