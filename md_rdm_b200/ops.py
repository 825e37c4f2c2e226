"""torch custom ops (`torch.ops.rdm.*`) over the C ABI of librdm_b200.so.

Each op is a thin argument check + output allocation + one C call on the current CUDA
stream; the arithmetic lives in md_rdm_b200/csrc/*.cu.  Inputs must be CUDA tensors: there
is no CPU implementation and no fallback (a CPU tensor raises).  Autograd:

* `make_pred`, `recombination`, `fuse_tail`: real backward (the only gradients the
  reference's training loss needs flow through these to the `Weights` parameters);
* `pair_v1`, `pair_id`, `lloyd_quantize`, `als_rank1`: ZERO gradients (not None): in the
  reference Lloyd overwrites every ratio with a table constant (network/RDM_Net.py:296-297,
  309-310), so gradients upstream of it are exactly 0.

Reference citations: RN = network/RDM_Net.py, CP = network/computations.py.
"""
from __future__ import annotations

import math
from ctypes import c_void_p
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import AlsScale, check, i32_array, load, ptr_array


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[Tensor]) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _need_cuda(name: str, *ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(f"rdm::{name}: expected CUDA tensors (md_rdm_b200 has no CPU path), got device {t.device}")


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return _cabi.DT_F32
    if t.dtype == torch.float64:
        return _cabi.DT_F64
    if t.dtype == torch.int64:
        return _cabi.DT_I64
    raise RuntimeError(f"unsupported dtype {t.dtype} (f32, f64 or int64 expected)")


# ============================================================================ stage 1
@torch.library.custom_op("rdm::pair_v1", mutates_args=())
def pair_v1(d3: Tensor) -> Tensor:
    """RN:244-252: (B,1,8,8) f32 -> raw pair matrix (B,64,64) f32 (before Lloyd)."""
    _need_cuda("pair_v1", d3)
    if d3.dtype != torch.float32 or d3.numel() % 64 != 0:
        raise RuntimeError("rdm::pair_v1: expected an f32 tensor of 8x8 maps")
    d = d3.contiguous()
    B = d.numel() // 64
    out = torch.empty((B, 64, 64), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        check(load().rdm_pair_v1_f32(_p(d), B, _p(out), _stream()), "rdm_pair_v1_f32")
    return out


@pair_v1.register_fake
def _(d3):
    return d3.new_empty((d3.numel() // 64, 64, 64))


@torch.library.custom_op("rdm::resize_half", mutates_args=())
def resize_half(x: Tensor) -> Tensor:
    """CP:308-311 for newsize == side/2: (B,C,s,s) f32|f64 -> (B,C,s/2,s/2) f64."""
    _need_cuda("resize_half", x)
    if x.dim() != 4 or x.shape[2] != x.shape[3] or x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::resize_half: expected (B,C,s,s) f32/f64")
    s = x.shape[2]
    xc = x.contiguous()
    out = torch.empty((x.shape[0], x.shape[1], s // 2, s // 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_resize_half(_p(xc), int(x.dtype == torch.float64), x.shape[0] * x.shape[1], s, _p(out), _stream()),
              "rdm_resize_half")
    return out


@resize_half.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2), dtype=torch.float64)


@torch.library.custom_op("rdm::pair_id", mutates_args=())
def pair_id(dn: Tensor) -> Tuple[Tensor, Tensor]:
    """RN:259-280 + CP:201-216 + the CP:308 resize feeding them: (B,1,s,s) f32, s in {16..128}
    -> raw (B,P,256,64) f64 for all P=(s/16)^2 pages (row-major), parent map (B,1,s/2,s/2) f64."""
    _need_cuda("pair_id", dn)
    if dn.dtype != torch.float32 or dn.dim() != 4 or dn.shape[1] != 1 or dn.shape[2] != dn.shape[3]:
        raise RuntimeError("rdm::pair_id: expected (B,1,s,s) f32")
    B, s = dn.shape[0], dn.shape[2]
    P = (s // 16) ** 2
    d = dn.contiguous()
    raw = torch.empty((B, P, 256, 64), dtype=torch.float64, device=dn.device)
    parent = torch.empty((B, 1, s // 2, s // 2), dtype=torch.float64, device=dn.device)
    with torch.cuda.device(dn.device):
        check(load().rdm_pair_id_f64(_p(d), B, s, _p(raw), _p(parent), _stream()), "rdm_pair_id_f64")
    return raw, parent


@pair_id.register_fake
def _(dn):
    B, s = dn.shape[0], dn.shape[2]
    return (dn.new_empty((B, (s // 16) ** 2, 256, 64), dtype=torch.float64),
            dn.new_empty((B, 1, s // 2, s // 2), dtype=torch.float64))


@torch.library.custom_op("rdm::pair_pages", mutates_args=())
def pair_pages(dn: Tensor, dn_1: Tensor) -> Tensor:
    """RN:259-280 sparse_comparison_id before its Lloyd call, literal signature: dn (B,1,16,16) f32,
    dn_1 (B,1,8,8) f64 (caller-supplied parent page) -> raw (B,256,64) f64."""
    _need_cuda("pair_pages", dn, dn_1)
    if dn.shape[-2:] != (16, 16) or dn_1.shape[-2:] != (8, 8):
        raise RuntimeError("rdm::pair_pages: expected 16x16 pages and 8x8 parent pages")
    B = dn.numel() // 256
    if dn_1.numel() != B * 64:
        raise RuntimeError("rdm::pair_pages: page / parent count mismatch")
    raw = torch.empty((B, 256, 64), dtype=torch.float64, device=dn.device)
    with torch.cuda.device(dn.device):
        check(load().rdm_pair_pages_f64(_p(dn.float().contiguous()), _p(dn_1.double().contiguous()), B, _p(raw), _stream()),
              "rdm_pair_pages_f64")
    return raw


@pair_pages.register_fake
def _(dn, dn_1):
    return dn.new_empty((dn.numel() // 256, 256, 64), dtype=torch.float64)


@torch.library.custom_op("rdm::resize_bicubic", mutates_args=())
def resize_bicubic(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """CP:308-311 for an arbitrary size: (B,C,H,W) f32|f64 -> (B,C,out_h,out_w) f64."""
    _need_cuda("resize_bicubic", x)
    if x.dim() != 4 or x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::resize_bicubic: expected (B,C,H,W) f32/f64")
    B, C, H, W = x.shape
    out = torch.empty((B, C, out_h, out_w), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_resize_bicubic_f64(_p(x.contiguous()), int(x.dtype == torch.float64), B * C, H, W, out_h, out_w, _p(out),
                                            _stream()), "rdm_resize_bicubic_f64")
    return out


@resize_bicubic.register_fake
def _(x, out_h, out_w):
    return x.new_empty((x.shape[0], x.shape[1], out_h, out_w), dtype=torch.float64)


@torch.library.custom_op("rdm::upsample_nearest", mutates_args=())
def upsample_nearest(x: Tensor, times: int) -> Tensor:
    """CP:357-366: `.double()` + nearest x2, `times` times: (B,C,s,s) -> (B,C,s<<times,s<<times) f64."""
    _need_cuda("upsample_nearest", x)
    if x.dim() != 4 or x.shape[2] != x.shape[3]:
        raise RuntimeError("rdm::upsample_nearest: expected (B,C,s,s)")
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    B, C, s, _ = x.shape
    out = torch.empty((B, C, s << times, s << times), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_upsample_nearest_f64(_p(x.contiguous()), int(x.dtype == torch.float64), B * C, s, times, _p(out), _stream()),
              "rdm_upsample_nearest_f64")
    return out


@upsample_nearest.register_fake
def _(x, times):
    return x.new_empty((x.shape[0], x.shape[1], x.shape[2] << times, x.shape[3] << times), dtype=torch.float64)


# ============================================================================ stage 2
@torch.library.custom_op("rdm::lloyd_quantize", mutates_args=())
def lloyd_quantize(x: Tensor, thresholds: Tensor, levels: Tensor) -> Tuple[Tensor, Tensor]:
    """RN:286-311: (values like x, bins u8).  The compare runs in x's dtype."""
    _need_cuda("lloyd_quantize", x, thresholds, levels)
    if x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::lloyd_quantize: x must be f32 or f64")
    if thresholds.dtype != torch.float64 or thresholds.numel() != 40 or levels.dtype != torch.float64 or levels.numel() != 41:
        raise RuntimeError("rdm::lloyd_quantize: thresholds f64[40] and levels f64[41] expected")
    xc = x.contiguous()
    values = torch.empty_like(xc)
    bins = torch.empty(xc.shape, dtype=torch.uint8, device=x.device)
    fn = load().rdm_lloyd_quantize_f32 if x.dtype == torch.float32 else load().rdm_lloyd_quantize_f64
    with torch.cuda.device(x.device):
        check(fn(_p(xc), xc.numel(), _p(thresholds.contiguous()), _p(levels.contiguous()), _p(values), _p(bins), _stream()),
              "rdm_lloyd_quantize")
    return values, bins


@lloyd_quantize.register_fake
def _(x, thresholds, levels):
    return torch.empty_like(x), x.new_empty(x.shape, dtype=torch.uint8)


# ============================================================================ stage 3
_KIND_DTYPE = {_cabi.SRC_RAW_F64: torch.float64, _cabi.SRC_RAW_F32: torch.float32, _cabi.SRC_VAL_F32: torch.float32,
               _cabi.SRC_VAL_F64: torch.float64, _cabi.SRC_MAP_F32: torch.float32}


@torch.library.custom_op("rdm::als_rank1", mutates_args=())
def als_rank1(src: Tensor, kind: int, rows: int, side: int, limit: int, group: int, thresholds: Optional[Tensor],
              levels: Optional[Tensor], want_bins: bool, want_values: bool, flags: int = 0) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Lloyd (optional) + rank-1 ALS + batch-wide arg-min + gm normalisation + re-tiling for ONE
    scale (RN:358-396 minus DORN; CP:38-85, CP:95-155, CP:218-238).

    src: kind RAW_*/VAL_*: (N,P,rows,64) (or (N,rows,64) when P == 1); kind MAP_F32: (N,1,side,side).
    Returns (map (N,1,side,side) f32, pages (N,P,rows) f32, rmse record (N/group,P,limit+1) f32,
    kstar (N/group,P) i32, bins u8 (N,P,rows,64) or empty, values f32 (N,P,rows,64) or empty).
    flags: `_cabi.ALS_*` bits (rdm_als_scale_t.flags); 0 = the reference's behaviour."""
    _need_cuda("als_rank1", src, thresholds, levels)
    if src.dtype != _KIND_DTYPE[kind]:
        raise RuntimeError(f"rdm::als_rank1: src dtype {src.dtype} does not match kind {kind}")
    P = 1 if rows == 64 else (side // 16) ** 2
    s = src.contiguous()
    per = side * side if kind == _cabi.SRC_MAP_F32 else P * rows * 64
    if s.numel() % per != 0:
        raise RuntimeError("rdm::als_rank1: src size is not a whole number of images")
    N = s.numel() // per
    dev = s.device
    lib = load()
    ws = torch.empty((N * lib.rdm_als_ws_floats(rows, P, limit),), dtype=torch.float32, device=dev)
    out_map = torch.empty((N, 1, side, side), dtype=torch.float32, device=dev)
    pages = torch.empty((N, P, rows), dtype=torch.float32, device=dev)
    G = max(N // max(group, 1), 0)
    record = torch.empty((G, P, limit + 1), dtype=torch.float32, device=dev)
    kstar = torch.empty((G, P), dtype=torch.int32, device=dev)
    bins = torch.empty((N, P, rows, 64) if want_bins else (0,), dtype=torch.uint8, device=dev)
    values = torch.empty((N, P, rows, 64) if want_values else (0,), dtype=torch.float32, device=dev)
    sc = AlsScale(src=s.data_ptr(), src_kind=kind, rows=rows, pages=P, side=side, limit=limit, flags=int(flags),
                  thresholds=thresholds.data_ptr() if thresholds is not None else None,
                  levels=levels.data_ptr() if levels is not None else None,
                  bins_out=bins.data_ptr() if want_bins else None, values_out=values.data_ptr() if want_values else None,
                  pages_out=pages.data_ptr(), map_out=out_map.data_ptr(), ws=ws.data_ptr(),
                  record_out=record.data_ptr(), kstar_out=kstar.data_ptr())
    with torch.cuda.device(dev):
        check(lib.rdm_als_fused(sc, 1, N, group, _stream()), "rdm_als_fused")
    return out_map, pages, record, kstar, bins, values


@als_rank1.register_fake
def _(src, kind, rows, side, limit, group, thresholds, levels, want_bins, want_values, flags=0):
    P = 1 if rows == 64 else (side // 16) ** 2
    per = side * side if kind == _cabi.SRC_MAP_F32 else P * rows * 64
    N = src.numel() // per
    G = N // max(group, 1)
    f32 = dict(dtype=torch.float32)
    return (src.new_empty((N, 1, side, side), **f32), src.new_empty((N, P, rows), **f32),
            src.new_empty((G, P, limit + 1), **f32), src.new_empty((G, P), dtype=torch.int32),
            src.new_empty((N, P, rows, 64) if want_bins else (0,), dtype=torch.uint8),
            src.new_empty((N, P, rows, 64) if want_values else (0,), **f32))


@torch.library.custom_op("rdm::als_step", mutates_args=())
def als_step(ratings: Tensor, fixed: Tensor, reg: float) -> Tensor:
    """CP:175-193: ratings (B,H,W) f32, fixed (B,W,1) f32 -> (B,H,1) f32."""
    _need_cuda("als_step", ratings, fixed)
    if ratings.dtype != torch.float32 or fixed.dtype != torch.float32 or ratings.dim() != 3:
        raise RuntimeError("rdm::als_step: f32 (B,H,W) ratings and (B,W,1) fixed expected")
    B, H, W = ratings.shape
    if fixed.numel() != B * W:
        raise RuntimeError("rdm::als_step: fixed must be (B,W,1)")
    out = torch.empty((B, H, 1), dtype=torch.float32, device=ratings.device)
    with torch.cuda.device(ratings.device):
        check(load().rdm_als_step_f32(_p(ratings.contiguous()), _p(fixed.contiguous()), B, H, W, reg, _p(out), _stream()),
              "rdm_als_step_f32")
    return out


@als_step.register_fake
def _(ratings, fixed, reg):
    return ratings.new_empty((ratings.shape[0], ratings.shape[1], 1))


# ============================================================================ stage 4
@torch.library.custom_op("rdm::quick_gm", mutates_args=())
def quick_gm(t: Tensor, rc: int) -> Tensor:
    """CP:244-255: prod over dim 1 of pow(t, 1/rc^2); t (B,N,1) -> (B,1).  int64 -> f32."""
    _need_cuda("quick_gm", t)
    code = _dtype_code(t)
    B = t.shape[0]
    n = t.numel() // max(B, 1)
    out = torch.empty((B, 1), dtype=torch.float64 if code == _cabi.DT_F64 else torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        check(load().rdm_quick_gm(_p(t.contiguous()), code, B, n, rc, _p(out), _stream()), "rdm_quick_gm")
    return out


@quick_gm.register_fake
def _(t, rc):
    return t.new_empty((t.shape[0], 1), dtype=torch.float64 if t.dtype == torch.float64 else torch.float32)


@torch.library.custom_op("rdm::gm_normalize", mutates_args=())
def gm_normalize(x: Tensor) -> Tensor:
    """RN:117 / network/module.py:145-149: x / quick_gm(x.view(B,HW,1), H) for (B,1,s,s)."""
    _need_cuda("gm_normalize", x)
    code = _dtype_code(x)
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != x.shape[3]:
        raise RuntimeError("rdm::gm_normalize: expected (B,1,s,s)")
    out = torch.empty(x.shape, dtype=torch.float64 if code == _cabi.DT_F64 else torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_gm_normalize(_p(x.contiguous()), code, x.shape[0], x.shape[2], _p(out), _stream()), "rdm_gm_normalize")
    return out


@gm_normalize.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=torch.float64 if x.dtype == torch.float64 else torch.float32)


@torch.library.custom_op("rdm::gm_bwd", mutates_args=())
def gm_bwd(x: Tensor, rc: int, grad_gm: Optional[Tensor], grad_norm: Optional[Tensor]) -> Tensor:
    """Backward of quick_gm (grad_gm) and/or gm_normalize (grad_norm) for floating x viewed (B,n)."""
    B = x.shape[0]
    n = x.numel() // max(B, 1)
    xc = x.contiguous()
    gx = torch.empty_like(xc)
    gg = grad_gm.to(x.dtype).contiguous() if grad_gm is not None else None
    gn = grad_norm.to(x.dtype).contiguous() if grad_norm is not None else None
    with torch.cuda.device(x.device):
        check(load().rdm_gm_bwd(_p(xc), int(x.dtype == torch.float64), B, n, rc, _p(gg), _p(gn), _p(gx), _stream()), "rdm_gm_bwd")
    return gx


@gm_bwd.register_fake
def _(x, rc, grad_gm, grad_norm):
    return torch.empty_like(x)


def _gm_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.rc = inputs[1] if len(inputs) > 1 else int(inputs[0].shape[2])


def _quick_gm_backward(ctx, g):
    (t,) = ctx.saved_tensors
    if not t.is_floating_point():
        return None, None
    return torch.ops.rdm.gm_bwd(t, ctx.rc, g, None), None


def _gm_normalize_backward(ctx, g):
    (x,) = ctx.saved_tensors
    if not x.is_floating_point():
        return None
    return torch.ops.rdm.gm_bwd(x, ctx.rc, None, g)


quick_gm.register_autograd(_quick_gm_backward, setup_context=_gm_setup)
gm_normalize.register_autograd(_gm_normalize_backward, setup_context=_gm_setup)


def pyramid_len(side: int, relative_map: bool) -> int:
    return int(load().rdm_pyramid_len(side, int(relative_map)))


@torch.library.custom_op("rdm::decompose", mutates_args=())
def decompose(x: Tensor, relative_map: bool) -> Tensor:
    """CP:368-392 whole pyramid in one launch: (B,1,s,s) f32|f64 -> flat f64 buffer of B*len values,
    LEVEL-MAJOR: [D_0 (B,1) unless relative_map][F_1 (B,4)][F_2 (B,16)]...[F_n (B,s*s)]."""
    _need_cuda("decompose", x)
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != x.shape[3] or x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::decompose: expected (B,1,s,s) f32/f64")
    B, s = x.shape[0], x.shape[2]
    out = torch.empty((B * pyramid_len(s, relative_map),), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_decompose(_p(x.contiguous()), int(x.dtype == torch.float64), B, s, int(relative_map), _p(out), _stream()),
              "rdm_decompose")
    return out


@decompose.register_fake
def _(x, relative_map):
    n = int(math.log2(x.shape[2]))
    return x.new_empty((x.shape[0] * ((0 if relative_map else 1) + (4 ** (n + 1) - 4) // 3),), dtype=torch.float64)


@torch.library.custom_op("rdm::decompose_bwd", mutates_args=())
def decompose_bwd(x: Tensor, relative_map: bool, grad_pyramid: Tensor) -> Tensor:
    xc = x.contiguous()
    gin = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        check(load().rdm_decompose_bwd(_p(xc), int(x.dtype == torch.float64), x.shape[0], x.shape[2], int(relative_map),
                                       _p(grad_pyramid.double().contiguous()), _p(gin), _stream()), "rdm_decompose_bwd")
    return gin


@decompose_bwd.register_fake
def _(x, relative_map, grad_pyramid):
    return torch.empty_like(x)


def _decompose_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.relative = inputs[1]


def _decompose_backward(ctx, g):
    (x,) = ctx.saved_tensors
    if x.shape[2] < 2:
        return g.view(x.shape).to(x.dtype), None
    return torch.ops.rdm.decompose_bwd(x, ctx.relative, g), None


decompose.register_autograd(_decompose_backward, setup_context=_decompose_setup)


_SID_CONST = {}


def _sid_constants(K: float = 90.0, alpha: float = 0.02, beta: float = 10.0):
    """The f32 scalars utils.depth2label_sid holds (utils.py:196-198, 205), as Python floats: K, alpha and
    log(beta / alpha) evaluated by torch in f32 exactly like the reference does."""
    key = (K, alpha, beta)
    if key not in _SID_CONST:
        a, b, k = torch.tensor(alpha), torch.tensor(beta), torch.tensor(K)
        _SID_CONST[key] = (float(k), float(a), float(torch.log(b / a)))
    return _SID_CONST[key]


@torch.library.custom_op("rdm::gt_prepare", mutates_args=())
def gt_prepare(y_raw: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Ground-truth preparation of the training step in one launch (network/module.py:68, 74-78, 119-127, 134-149,
    utils.py:195-211): y_raw (B,1,H,W) f32|f64 -> (y (B,1,128,128) f64 masked, packed component targets (level-major,
    see `unpack_pyramid(packed, B, 128, False)`; slot 0 = the ordinal D_0), SID labels (B,1,8,8) int32)."""
    _need_cuda("gt_prepare", y_raw)
    if y_raw.dim() != 4 or y_raw.shape[1] != 1 or y_raw.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::gt_prepare: expected (B,1,H,W) f32 or f64")
    x = y_raw.contiguous()
    B, _, H, W = x.shape
    dev = x.device
    lib = load()
    y = torch.empty((B, 1, 128, 128), dtype=torch.float64, device=dev)
    pyr = torch.empty((B * lib.rdm_pyramid_len(128, 0),), dtype=torch.float64, device=dev)
    ord_t = torch.empty((B, 1, 8, 8), dtype=torch.int32, device=dev)
    K, a, lr = _sid_constants()
    with torch.cuda.device(dev):
        check(lib.rdm_gt_prepare(_p(x), 1 if x.dtype == torch.float64 else 0, B, H, W, K, a, lr, _p(y), _p(pyr), _p(ord_t), _stream()),
              "rdm_gt_prepare")
    return y, pyr, ord_t


@gt_prepare.register_fake
def _(y_raw):
    B = y_raw.shape[0]
    return (y_raw.new_empty((B, 1, 128, 128), dtype=torch.float64), y_raw.new_empty((B * 21845,), dtype=torch.float64),
            y_raw.new_empty((B, 1, 8, 8), dtype=torch.int32))


def unpack_pyramid(packed: Tensor, batch: int, side: int, relative_map: bool) -> List[Tensor]:
    """Dense views [D_0?, F_1, ..., F_n], each (B,1,2^k,2^k), of the level-major pyramid buffer."""
    n = int(math.log2(side))
    out, off = [], 0
    if not relative_map:
        out.append(packed[0:batch].view(batch, 1, 1, 1))
        off = batch
    for k in range(1, n + 1):
        m = 4 ** k
        out.append(packed[off:off + batch * m].view(batch, 1, 2 ** k, 2 ** k))
        off += batch * m
    return out


# ============================================================================ stage 5
@torch.library.custom_op("rdm::log_stack", mutates_args=())
def log_stack(cands: Sequence[Tensor]) -> Tensor:
    """CP:464-484 make_matrix: K candidates (B,1,s,s) f64 -> (B,K,s*s) f64 of logs."""
    _need_cuda("log_stack", *cands)
    B = cands[0].shape[0]
    M = cands[0].numel() // B
    cs = [c.double().contiguous() for c in cands]
    out = torch.empty((B, len(cs), M), dtype=torch.float64, device=cs[0].device)
    with torch.cuda.device(out.device):
        check(load().rdm_log_stack_f64(ptr_array([c.data_ptr() for c in cs]), len(cs), B, M, _p(out), _stream()),
              "rdm_log_stack_f64")
    return out


@log_stack.register_fake
def _(cands):
    B = cands[0].shape[0]
    return cands[0].new_empty((B, len(cands), cands[0].numel() // B), dtype=torch.float64)


@torch.library.custom_op("rdm::log_stack_bwd", mutates_args=())
def log_stack_bwd(cands: Sequence[Tensor], grad_out: Tensor) -> List[Tensor]:
    B = cands[0].shape[0]
    M = cands[0].numel() // B
    cs = [c.double().contiguous() for c in cands]
    gs = [torch.empty_like(c) for c in cs]
    with torch.cuda.device(grad_out.device):
        check(load().rdm_log_stack_bwd(ptr_array([c.data_ptr() for c in cs]), len(cs), B, M, _p(grad_out.double().contiguous()),
                                       ptr_array([g.data_ptr() for g in gs]), _stream()), "rdm_log_stack_bwd")
    return gs


@log_stack_bwd.register_fake
def _(cands, grad_out):
    return [torch.empty_like(c, dtype=torch.float64) for c in cands]


def _log_stack_setup(ctx, inputs, output):
    ctx.cands = list(inputs[0])


def _log_stack_backward(ctx, g):
    gs = torch.ops.rdm.log_stack_bwd(ctx.cands, g)
    return ([gi.to(c.dtype).view(c.shape) for gi, c in zip(gs, ctx.cands)],)


log_stack.register_autograd(_log_stack_backward, setup_context=_log_stack_setup)


@torch.library.custom_op("rdm::make_pred", mutates_args=())
def make_pred(A: Tensor, w: Tensor) -> Tensor:
    """CP:512-528 for one slot: A (B,K,M) f64, w (K,1) f32 -> (B,M) f32."""
    _need_cuda("make_pred", A, w)
    if A.dtype != torch.float64 or A.dim() != 3 or w.numel() != A.shape[1]:
        raise RuntimeError("rdm::make_pred: A (B,K,M) f64 and w with K entries expected")
    B, K, M = A.shape
    out = torch.empty((B, M), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(load().rdm_make_pred_f32(_p(A.contiguous()), _p(w.float().contiguous()), B, K, M, _p(out), _stream()),
              "rdm_make_pred_f32")
    return out


@make_pred.register_fake
def _(A, w):
    return A.new_empty((A.shape[0], A.shape[2]), dtype=torch.float32)


@torch.library.custom_op("rdm::make_pred_bwd", mutates_args=())
def make_pred_bwd(A: Tensor, w: Tensor, grad_out: Tensor) -> Tuple[Tensor, Tensor]:
    B, K, M = A.shape
    gA = torch.empty((B, K, M), dtype=torch.float64, device=A.device)
    gw = torch.empty((K,), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(load().rdm_make_pred_bwd(_p(A.contiguous()), _p(w.float().contiguous()), _p(grad_out.float().contiguous()), B, K, M,
                                       _p(gA), _p(gw), _stream()), "rdm_make_pred_bwd")
    return gA, gw


@make_pred_bwd.register_fake
def _(A, w, grad_out):
    return torch.empty_like(A), A.new_empty((A.shape[1],), dtype=torch.float32)


def _make_pred_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _make_pred_backward(ctx, g):
    A, w = ctx.saved_tensors
    gA, gw = torch.ops.rdm.make_pred_bwd(A, w, g)
    return gA, gw.view(w.shape).to(w.dtype)


make_pred.register_autograd(_make_pred_backward, setup_context=_make_pred_setup)


def _comp_list(comps: Sequence[Tensor]):
    dt = comps[0].dtype
    if dt not in (torch.float32, torch.float64) or any(c.dtype != dt for c in comps):
        raise RuntimeError("rdm::recombination: components must all be f32 or all f64")
    cs = [c.contiguous() for c in comps]
    return cs, [int(c.shape[2]) for c in cs], int(dt == torch.float64)


@torch.library.custom_op("rdm::recombination", mutates_args=())
def recombination(comps: Sequence[Tensor], n: int) -> Tensor:
    """CP:394-421: list [(d_0)?, f_1, f_2, ...] of (B,1,2^k,2^k) -> (B,1,2^n,2^n) f64."""
    _need_cuda("recombination", *comps)
    cs, sides, is64 = _comp_list(comps)
    B = cs[0].shape[0]
    out = torch.empty((B, 1, 2 ** n, 2 ** n), dtype=torch.float64, device=cs[0].device)
    with torch.cuda.device(out.device):
        check(load().rdm_recombination_f64(ptr_array([c.data_ptr() for c in cs]), i32_array(sides), len(cs), is64, B, n,
                                           _p(out), _stream()), "rdm_recombination_f64")
    return out


@recombination.register_fake
def _(comps, n):
    return comps[0].new_empty((comps[0].shape[0], 1, 2 ** n, 2 ** n), dtype=torch.float64)


@torch.library.custom_op("rdm::recombination_bwd", mutates_args=())
def recombination_bwd(grad_out: Tensor, sides: Sequence[int], is_f64: bool, n: int) -> List[Tensor]:
    B = grad_out.shape[0]
    dt = torch.float64 if is_f64 else torch.float32
    gs = [torch.empty((B, 1, s, s), dtype=dt, device=grad_out.device) for s in sides]
    with torch.cuda.device(grad_out.device):
        check(load().rdm_recombination_bwd(_p(grad_out.double().contiguous()), ptr_array([g.data_ptr() for g in gs]),
                                           i32_array(list(sides)), len(gs), int(is_f64), B, n, _stream()), "rdm_recombination_bwd")
    return gs


@recombination_bwd.register_fake
def _(grad_out, sides, is_f64, n):
    dt = torch.float64 if is_f64 else torch.float32
    return [grad_out.new_empty((grad_out.shape[0], 1, s, s), dtype=dt) for s in sides]


def _recomb_setup(ctx, inputs, output):
    comps, n = inputs
    ctx.sides = [int(c.shape[2]) for c in comps]
    ctx.is64 = comps[0].dtype == torch.float64
    ctx.n = n


def _recomb_backward(ctx, g):
    return torch.ops.rdm.recombination_bwd(g, ctx.sides, ctx.is64, ctx.n), None


recombination.register_autograd(_recomb_backward, setup_context=_recomb_setup)


# ---------------------------------------------------------------------------- fused stages 4+5
def tail_layout(sides: Sequence[int]):
    """Slot bookkeeping for decoder 1 + relative decoders of `sides`: (K per slot [8], weight
    offset per slot [8], kmax, total weights)."""
    K = [1, 1, 1, 1, 0, 0, 0, 0]
    for s in sides:
        for k in range(1, int(math.log2(s)) + 1):
            K[k] += 1
    off, acc = [], 0
    for k in range(8):
        off.append(acc)
        acc += K[k]
    kmax = max([3] + [int(math.log2(s)) for s in sides])
    return K, off, kmax, acc


@torch.library.custom_op("rdm::fuse_tail", mutates_args=())
def fuse_tail(x_d1: Tensor, rel: Sequence[Tensor], weights: Tensor, want_A: bool, bands: int = 0) -> Tuple[Tensor, Tensor, List[Tensor]]:
    """RN:117-133 + network/module.py:132 in one launch.  x_d1 (B,1,8,8) int64; rel: filled
    relative maps (B,1,s,s) f32 (s <= 64); weights: flat f32, slots concatenated [d0|f1|...].
    Returns (depth (B,1,128,128) f64, yhat packed (B, sum 4^k) f32, [A_k (B,K_k,4^k) f64] if want_A).
    bands: CTAs per image (1/2/4/8; 0 = chosen from the batch, best with many calls in flight; a call that is alone on the
    GPU is shorter with 4); the results do not depend on it."""
    _need_cuda("fuse_tail", x_d1, weights, *rel)
    if x_d1.dtype != torch.int64 or x_d1.numel() % 64:
        raise RuntimeError("rdm::fuse_tail: x_d1 must be int64 (B,1,8,8)")
    B = x_d1.numel() // 64
    sides = [int(r.shape[2]) for r in rel]
    K, _, kmax, nw = tail_layout(sides)
    if weights.dtype != torch.float32 or weights.numel() != nw:
        raise RuntimeError(f"rdm::fuse_tail: expected {nw} f32 weights for sides {sides}, got {weights.numel()} {weights.dtype}")
    if any(r.dtype != torch.float32 or r.shape[0] != B for r in rel):
        raise RuntimeError("rdm::fuse_tail: relative maps must be f32 (B,1,s,s)")
    dev = x_d1.device
    rc = [r.contiguous() for r in rel]
    depth = torch.empty((B, 1, 128, 128), dtype=torch.float64, device=dev)
    yhat = torch.empty((B, (4 ** (kmax + 1) - 1) // 3), dtype=torch.float32, device=dev)
    A = [torch.empty((B, K[k], 4 ** k), dtype=torch.float64, device=dev) for k in range(kmax + 1)] if want_A else []
    a_ptrs = ptr_array([A[k].data_ptr() if (want_A and k <= kmax) else None for k in range(8)])
    with torch.cuda.device(dev):
        check(load().rdm_fuse_tail_bands(_p(x_d1.contiguous()), ptr_array([r.data_ptr() for r in rc]), i32_array(sides), len(rc),
                                         _p(weights.contiguous()), B, _p(yhat), _p(depth), c_void_p(0), a_ptrs, int(bands), _stream()),
              "rdm_fuse_tail_bands")
    return depth, yhat, A


@fuse_tail.register_fake
def _(x_d1, rel, weights, want_A, bands=0):
    B = x_d1.numel() // 64
    sides = [int(r.shape[2]) for r in rel]
    K, _, kmax, _ = tail_layout(sides)
    A = [x_d1.new_empty((B, K[k], 4 ** k), dtype=torch.float64) for k in range(kmax + 1)] if want_A else []
    return (x_d1.new_empty((B, 1, 128, 128), dtype=torch.float64),
            x_d1.new_empty((B, (4 ** (kmax + 1) - 1) // 3), dtype=torch.float32), A)


@torch.library.custom_op("rdm::fuse_tail_bwd", mutates_args=())
def fuse_tail_bwd(grad_depth: Tensor, A: Sequence[Tensor]) -> Tensor:
    """Gradient of fuse_tail's depth with respect to the flat weights, two launches (pooled gradients of the
    recombination, then every weight's reduction); bit-identical to recombination_bwd + make_pred_bwd slot by slot."""
    _need_cuda("fuse_tail_bwd", grad_depth, *A)
    B = A[0].shape[0]
    kmax = len(A) - 1
    K = [int(a.shape[1]) for a in A]
    dev = grad_depth.device
    g = grad_depth.double().contiguous()
    As = [a.contiguous() for a in A]
    ws = torch.empty((B * ((4 ** (kmax + 1) - 1) // 3),), dtype=torch.float32, device=dev)
    gw = torch.empty((sum(K),), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(load().rdm_fuse_tail_bwd(_p(g), ptr_array([a.data_ptr() for a in As]), i32_array(K), kmax, B, _p(ws), _p(gw), _stream()),
              "rdm_fuse_tail_bwd")
    return gw


@fuse_tail_bwd.register_fake
def _(grad_depth, A):
    return grad_depth.new_empty((sum(int(a.shape[1]) for a in A),), dtype=torch.float32)


@torch.library.custom_op("rdm::component_loss", mutates_args=())
def component_loss(yhat: Tensor, target_pyramid: Tensor, kmax: int) -> Tensor:
    """CP:499-510 in one launch: sum over k <= kmax of the MSE between y_hat_k (packed, fuse_tail) and the level-major
    target pyramid (gt_prepare / decompose_packed); f64 scalar, no gradient (the reference detaches it)."""
    _need_cuda("component_loss", yhat, target_pyramid)
    if yhat.dtype != torch.float32 or target_pyramid.dtype != torch.float64 or yhat.dim() != 2:
        raise RuntimeError("rdm::component_loss: yhat must be packed (B, sum 4^k) f32 and the targets an f64 pyramid")
    B = yhat.shape[0]
    if yhat.shape[1] != (4 ** (kmax + 1) - 1) // 3 or target_pyramid.numel() < B * ((4 ** (kmax + 1) - 1) // 3):
        raise RuntimeError("rdm::component_loss: shapes do not match kmax")
    out = torch.empty((), dtype=torch.float64, device=yhat.device)
    with torch.cuda.device(yhat.device):
        check(load().rdm_component_loss(_p(yhat.contiguous()), _p(target_pyramid.contiguous()), B, kmax, _p(out), _stream()),
              "rdm_component_loss")
    return out


@component_loss.register_fake
def _(yhat, target_pyramid, kmax):
    return yhat.new_empty((), dtype=torch.float64)


def split_yhat(yhat: Tensor, kmax: int) -> List[Tensor]:
    """Views of the packed y_hat as the reference's list of (B,1,2^k,2^k) f32 tensors."""
    B = yhat.shape[0]
    out, off = [], 0
    for k in range(kmax + 1):
        m = 4 ** k
        out.append(yhat[:, off:off + m].view(B, 1, 2 ** k, 2 ** k))
        off += m
    return out


def fuse_tail_autograd(x_d1: Tensor, rel: Sequence[Tensor], weights: Tensor, bands: int = 0):
    """fuse_tail with gradients to `weights` (and to y_hat consumers): the forward is the single
    fused launch; the backward is rdm_fuse_tail_bwd on the saved A (or, when y_hat carries a gradient,
    recombination_bwd and make_pred_bwd slot by slot).  bands: see fuse_tail."""
    return _FuseTailFn.apply(x_d1, weights, int(bands), *rel)


class _FuseTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_d1, weights, bands, *rel):
        depth, yhat, A = torch.ops.rdm.fuse_tail(x_d1, list(rel), weights, True, bands)
        ctx.set_materialize_grads(False)   # y_hat feeds only the detached component loss: its gradient stays None
        sides = [int(r.shape[2]) for r in rel]
        K, off, kmax, _ = tail_layout(sides)
        ctx.layout = (K, off, kmax)
        ctx.n_rel = len(rel)
        ctx.save_for_backward(weights, *A)
        return depth, yhat

    @staticmethod
    def backward(ctx, g_depth, g_yhat):
        weights, *A = ctx.saved_tensors
        K, off, kmax = ctx.layout
        B = A[0].shape[0]
        sides = [2 ** k for k in range(kmax + 1)]
        if g_depth is not None and g_yhat is None:   # the training step: two launches
            return (None, torch.ops.rdm.fuse_tail_bwd(g_depth, list(A)).to(weights.dtype), None) + tuple(None for _ in range(ctx.n_rel))
        if g_depth is not None:
            gs = torch.ops.rdm.recombination_bwd(g_depth, sides, False, 7)
        else:
            gs = [torch.zeros((B, 1, s, s), dtype=torch.float32, device=weights.device) for s in sides]
        if g_yhat is not None:
            gs = [g + gy for g, gy in zip(gs, split_yhat(g_yhat, kmax))]
        gw = torch.zeros_like(weights)
        for k in range(kmax + 1):
            _, gwk = torch.ops.rdm.make_pred_bwd(A[k], weights[off[k]:off[k] + K[k]], gs[k].reshape(B, -1))
            gw[off[k]:off[k] + K[k]] = gwk
        # relative maps / x_d1: zero gradients (numerically severed by Lloyd / integer input)
        return (None, gw, None) + tuple(None for _ in range(ctx.n_rel))


# ---------------------------------------------------------------------------- zero-gradient ops
def _zero_grad_backward(n_tensor_inputs: int, n_inputs: int):
    def backward(ctx, *grads):
        out = [torch.zeros(s, dtype=d, device=dev) if s is not None else None for (s, d, dev) in ctx.meta]
        return tuple(out) + (None,) * (n_inputs - n_tensor_inputs)
    return backward


def _zero_setup_1(ctx, inputs, output):
    t = inputs[0]
    ctx.meta = [(t.shape, t.dtype, t.device) if t.is_floating_point() else (None, None, None)]


pair_v1.register_autograd(_zero_grad_backward(1, 1), setup_context=_zero_setup_1)
pair_id.register_autograd(_zero_grad_backward(1, 1), setup_context=_zero_setup_1)


def _zero_setup_2(ctx, inputs, output):
    ctx.meta = [(t.shape, t.dtype, t.device) if t.is_floating_point() else (None, None, None) for t in inputs[:2]]


pair_pages.register_autograd(_zero_grad_backward(2, 2), setup_context=_zero_setup_2)
lloyd_quantize.register_autograd(_zero_grad_backward(1, 3), setup_context=_zero_setup_1)
als_rank1.register_autograd(_zero_grad_backward(1, 11), setup_context=_zero_setup_1)


# ============================================================================ SURVEY 8f "next": DORN head + ordinal loss
@torch.library.custom_op("rdm::dorn_regression", mutates_args=())
def dorn_regression(x: Tensor) -> Tuple[Tensor, Tensor]:
    """RN:313-345: x (N,2K,H,W) f32 -> (decode (N,1,H,W) int64, ord (N,K,H,W) f64)."""
    _need_cuda("dorn_regression", x)
    if x.dim() != 4 or x.shape[1] % 2 or x.dtype != torch.float32:
        raise RuntimeError("rdm::dorn_regression: expected (N,2K,H,W) f32")
    N, C, H, W = x.shape
    decode = torch.empty((N, 1, H, W), dtype=torch.int64, device=x.device)
    ord_ = torch.empty((N, C // 2, H, W), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        check(load().rdm_dorn_regression_f32(_p(x.contiguous()), N, C // 2, H * W, _p(decode), _p(ord_), _stream()), "rdm_dorn_regression_f32")
    return decode, ord_


@dorn_regression.register_fake
def _(x):
    N, C, H, W = x.shape
    return x.new_empty((N, 1, H, W), dtype=torch.int64), x.new_empty((N, C // 2, H, W), dtype=torch.float64)


@torch.library.custom_op("rdm::dorn_regression_bwd", mutates_args=())
def dorn_regression_bwd(x: Tensor, ord_: Tensor, grad_ord: Tensor) -> Tensor:
    N, C, H, W = x.shape
    gx = torch.empty_like(x, memory_format=torch.contiguous_format)
    with torch.cuda.device(x.device):
        check(load().rdm_dorn_regression_bwd(_p(x.contiguous()), _p(ord_.contiguous()), _p(grad_ord.double().contiguous()), N, C // 2, H * W,
                                             _p(gx), _stream()), "rdm_dorn_regression_bwd")
    return gx


@dorn_regression_bwd.register_fake
def _(x, ord_, grad_ord):
    return torch.empty_like(x)


def _dorn_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], output[1])


def _dorn_backward(ctx, g_decode, g_ord):
    x, ord_ = ctx.saved_tensors
    if g_ord is None:
        return torch.zeros_like(x)
    return torch.ops.rdm.dorn_regression_bwd(x, ord_, g_ord)


dorn_regression.register_autograd(_dorn_backward, setup_context=_dorn_setup)


@torch.library.custom_op("rdm::ordinal_loss", mutates_args=())
def ordinal_loss(ord_: Tensor, target: Tensor) -> Tensor:
    """loss.py:17-59: ord (N,K,H,W) f64, target (N,1,H,W) integer SID labels -> 0-d f32 loss."""
    _need_cuda("ordinal_loss", ord_, target)
    if ord_.dim() != 4 or ord_.dtype != torch.float64 or target.numel() != ord_.shape[0] * ord_.shape[2] * ord_.shape[3]:
        raise RuntimeError("rdm::ordinal_loss: expected ord (N,K,H,W) f64 and target (N,1,H,W)")
    N, K, H, W = ord_.shape
    lib = load()
    ws = torch.empty((lib.rdm_ordinal_loss_ws_doubles(),), dtype=torch.float64, device=ord_.device)
    loss = torch.empty((), dtype=torch.float32, device=ord_.device)
    with torch.cuda.device(ord_.device):
        check(lib.rdm_ordinal_loss_f64(_p(ord_.contiguous()), _p(target.to(torch.int32).contiguous()), N, K, H * W, _p(ws), _p(loss), _stream()),
              "rdm_ordinal_loss_f64")
    return loss


@ordinal_loss.register_fake
def _(ord_, target):
    return ord_.new_empty((), dtype=torch.float32)


@torch.library.custom_op("rdm::ordinal_loss_bwd", mutates_args=())
def ordinal_loss_bwd(ord_: Tensor, target: Tensor, grad_loss: Tensor) -> Tensor:
    N, K, H, W = ord_.shape
    g = torch.empty_like(ord_, memory_format=torch.contiguous_format)
    with torch.cuda.device(ord_.device):
        check(load().rdm_ordinal_loss_bwd(_p(ord_.contiguous()), _p(target.to(torch.int32).contiguous()), _p(grad_loss.float().contiguous()),
                                          N, K, H * W, _p(g), _stream()), "rdm_ordinal_loss_bwd")
    return g


@ordinal_loss_bwd.register_fake
def _(ord_, target, grad_loss):
    return torch.empty_like(ord_)


def _oloss_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _oloss_backward(ctx, g):
    ord_, target = ctx.saved_tensors
    return torch.ops.rdm.ordinal_loss_bwd(ord_, target, g), None


ordinal_loss.register_autograd(_oloss_backward, setup_context=_oloss_setup)


@torch.library.custom_op("rdm::depth2label_sid", mutates_args=())
def depth2label_sid(depth: Tensor, K: float, alpha: float, beta: float) -> Tensor:
    """utils.py:195-211: depth (any shape) f32|f64 -> SID labels int32 of the same shape."""
    _need_cuda("depth2label_sid", depth)
    if depth.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("rdm::depth2label_sid: expected an f32 or f64 tensor")
    d = depth.contiguous()
    out = torch.empty(d.shape, dtype=torch.int32, device=d.device)
    k, a, lr = _sid_constants(K, alpha, beta)
    with torch.cuda.device(d.device):
        check(load().rdm_depth2label_sid(_p(d), 1 if d.dtype == torch.float64 else 0, d.numel(), k, a, lr, _p(out), _stream()),
              "rdm_depth2label_sid")
    return out


@depth2label_sid.register_fake
def _(depth, K, alpha, beta):
    return depth.new_empty(depth.shape, dtype=torch.int32)
