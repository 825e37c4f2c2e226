"""Vectorised CPU restatement of the MD_RDM fusion path (the parity oracle).

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.  All citations are into the
reference tree (az16/MD_RDM): RN = network/RDM_Net.py, CP = network/computations.py,
MOD = network/module.py.

Parity status: PINNED against the reference code itself (imported unmodified in
the build container by tools/make_golden.py) -- bins and raw pair matrices
bit-equal, ALS maps bit-equal on the generating machine, see tests/golden/.

Everything here runs on CPU tensors and keeps the reference's dtypes:
f32 for the 8x8 pair matrix and for ALS, f64 for page pair matrices, resize,
decomposition, log matrices and the recombined map.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Sequence, Tuple

import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "md_rdm_b200", "data",
                     "depth_ratio_codebooks.json")

ALS_LAMBDA = 0.05            # CP:175 regularization_term
LIMIT_8 = 30                 # CP:38 default limit, called without limit at RN:364
LIMIT_PAGE = 100             # RN:378, RN:392


# --------------------------------------------------------------------------- codebooks
def load_codebooks(path: str = _DATA) -> Dict[int, Tuple[torch.Tensor, torch.Tensor]]:
    """{scale: (thresholds[40] f64, levels[41] f64)}; RN:397-418 (`Quantization`).

    Scale 8 is the derived stand-in (see tools/import_codebooks.py)."""
    with open(path) as f:
        raw = json.load(f)
    out = {}
    for key, tab in raw["tables"].items():
        q = torch.tensor([float.fromhex(h) for h in tab["thresholds"]], dtype=torch.float64)
        lv = torch.tensor([float.fromhex(h) for h in tab["levels"]], dtype=torch.float64)
        out[int(key)] = (q, lv)
    return out


def scale_of_id(layer_id: int) -> int:
    """RN:432-442 `get_size_id`: id 3..7 -> 8..128."""
    return 1 << layer_id


# --------------------------------------------------------------------------- stage 1: pair build
def pair_v1(d3: torch.Tensor) -> torch.Tensor:
    """RN:244-252.  R[b,i,j] = fl32(d_i * fl32(1/d_j)) over the row-major 8x8 map."""
    B = d3.shape[0]
    flat = d3.reshape(B, -1)
    inv = torch.pow(flat, -1)                      # RN:248 (bit-equal to 1/x)
    return flat.unsqueeze(2) * inv.unsqueeze(1)    # RN:252 K=1 matmul == elementwise product


_BICUBIC_HALF = (-0.09375, 0.59375, 0.59375, -0.09375)   # A=-0.75 cubic at t=0.5


def resize_half_explicit(x: torch.Tensor) -> torch.Tensor:
    """The arithmetic CP:308-311 performs for newsize == N/2, written out.

    `.double()` + bicubic(align_corners=False, A=-0.75) for an exact halving is a
    separable stride-2 4-tap filter: the source coordinate of output i is 2i+0.5,
    so taps are clamp(2i-1+a, 0, N-1), a=0..3, with weights (-3, 19, 19, -3)/32.
    ATen sums the horizontal taps first, then the vertical ones.

    Exactness: for f32-valued input (the only case that feeds the bit-exact Lloyd
    bins, RN:373/RN:386) every product and partial sum is exactly representable
    in f64, so this is BIT-EQUAL to ATen; for full-mantissa f64 input (deeper
    pyramid levels) it agrees to <= 4 ulp (rounding order inside ATen's compiled
    kernel is not recoverable from the binary).  tests/test_oracle_golden.py
    checks both statements.
    """
    x = x.double()
    N = x.shape[-1]
    M = N // 2
    idx = [torch.clamp(2 * torch.arange(M) - 1 + a, 0, N - 1) for a in range(4)]
    out = None
    for a in range(4):                         # vertical tap (outer)
        rows = x.index_select(-2, idx[a])
        inner = None
        for b in range(4):                     # horizontal tap (inner)
            t = rows.index_select(-1, idx[b]) * _BICUBIC_HALF[b]
            inner = t if inner is None else inner + t
        t = inner * _BICUBIC_HALF[a]
        out = t if out is None else out + t
    return out


def resize_half(x: torch.Tensor) -> torch.Tensor:
    """CP:308-311 with newsize == N/2, evaluated by the same ATen call the
    reference makes so the oracle stays bit-identical to it at every level."""
    return resize(x, x.shape[-1] // 2)


def resize(x: torch.Tensor, newsize: int) -> torch.Tensor:
    """CP:308-311, any size (used for the 226->128 GT resize, MOD:68)."""
    return torch.nn.functional.interpolate(x.double(), size=newsize, mode="bicubic", align_corners=False)


def window_mask(parent: int = 8, side: int = 16) -> torch.Tensor:
    """RN:266-273 + CP:269-295: bool (side*side, parent*parent); True where the
    3x3 window anchored at (min(r//2, parent-3), min(c//2, parent-3)) covers the
    parent pixel.  Row index r*side+c, column index rr*parent+cc."""
    m = torch.zeros(side * side, parent * parent, dtype=torch.bool)
    for r in range(side):
        r0 = min(r // 2, parent - 3)
        for c in range(side):
            c0 = min(c // 2, parent - 3)
            for dr in range(3):
                for dc in range(3):
                    m[r * side + c, (r0 + dr) * parent + (c0 + dc)] = True
    return m


_MASK_16_8 = None


def pair_id(dn: torch.Tensor, dn_1: torch.Tensor) -> torch.Tensor:
    """RN:259-280: raw (B,256,64) f64 pair matrix of one 16x16 page `dn` (f32)
    against its 8x8 parent page `dn_1` (f64).

    Row r*16+c equals f64(dn[r,c]) * fl64(1/area) where area is 1 everywhere
    except the 3x3 window copied from dn_1 (CP:284-287)."""
    global _MASK_16_8
    if _MASK_16_8 is None:
        _MASK_16_8 = window_mask(8, 16)
    B = dn.shape[0]
    d = dn.reshape(B, 256, 1).double()
    area = torch.where(_MASK_16_8.unsqueeze(0), dn_1.reshape(B, 1, 64).double(),
                       torch.ones((), dtype=torch.float64))
    return d * torch.pow(area, -1)


def split_pages(dn: torch.Tensor, dn_1: torch.Tensor):
    """CP:201-216: row-major list of (16x16 page, 8x8 parent page) views."""
    ratio = dn.shape[2] // 16
    pages = []
    for i in range(ratio):
        for j in range(ratio):
            pages.append((dn[:, :, 16 * i:16 * i + 16, 16 * j:16 * j + 16],
                          dn_1[:, :, 8 * i:8 * i + 8, 8 * j:8 * j + 8]))
    return pages


# --------------------------------------------------------------------------- stage 2: Lloyd
def lloyd(x: torch.Tensor, thresholds: torch.Tensor, levels: torch.Tensor):
    """RN:286-311.  bin = #{i: x >= q_i} with the compare done in x's dtype (the
    f64 threshold is rounded to f32 first when x is f32 -- python-scalar
    promotion), value = levels[bin] rounded to x's dtype.  Returns (values, bins u8)."""
    q = thresholds.to(x.dtype)
    bins = (x.unsqueeze(-1) >= q).sum(-1)
    values = levels.to(x.dtype)[bins]
    return values, bins.to(torch.uint8)


# --------------------------------------------------------------------------- stage 3: ALS
def _ridge_step(ratings: torch.Tensor, fixed: torch.Tensor) -> torch.Tensor:
    """CP:175-193 `als_step`: (ratings @ f) @ inverse(f^T f + 0.05 I_1)."""
    B, n, _ = fixed.shape
    A = torch.matmul(fixed.view(B, 1, n), fixed) + torch.eye(1) * ALS_LAMBDA
    return (ratings @ fixed) @ torch.inverse(A)


# "paper-correct" knobs, OFF by default; same bits as rdm_als_scale_t.flags in include/rdm_b200.h
FLAG_TRUE_TRANSPOSE, FLAG_TRUE_GM, FLAG_CORRECT_TILING = 2, 4, 8


def als_rank1(Rq: torch.Tensor, limit: int, force_k=None, flags: int = 0):
    """CP:38-85 (H=W=64) and CP:95-155 (H=256, W=64) in one routine.

    Returns (map (B,1,sqrt(H),sqrt(H)) f32, rmse record list[limit+1] of python
    floats, kstar).  Reproduces: the reshape-not-transpose `view(B,W,H)` (CP:64,
    CP:133), the batch-wide rmse (CP:172-173) with first-arg-min selection
    (CP:74, CP:143) and the `quick_gm(p, H)` normaliser whose exponent is
    1/H**2 (CP:244-255)."""
    B, H, W = Rq.shape
    R = Rq.float()
    p = torch.ones(B, H, 1)
    q = torch.ones(B, W, 1)
    record, vecs = [], []

    def rmse():
        return torch.mean((torch.matmul(p, q.view(B, 1, W)) - R) ** 2) ** 0.5

    record.append(rmse())
    vecs.append(p)
    # CP:64 / CP:133 pass R.view(B,W,H) - a reshape; FLAG_TRUE_TRANSPOSE uses the transpose the ALS update calls for
    Rv = R.transpose(1, 2).contiguous() if flags & FLAG_TRUE_TRANSPOSE else R.view(B, W, H)
    for _ in range(limit):
        p = _ridge_step(R, q)
        record.append(rmse())
        vecs.append(p)
        q = _ridge_step(Rv, p)
    kstar = record.index(min(record))
    # force_k (tests only): emit the iterate of a given index instead of the arg-min.  When the record
    # plateaus (smooth maps: values equal to 1 f32 ulp) the arg-min is decided by summation-order noise
    # and no two implementations - or BLAS builds - agree on it; parity is then stated on the record and
    # on the iterate at a common index.
    p = vecs[kstar if force_k is None else force_k]
    # CP:248-253 with rc=H: exponent 1/H^2; FLAG_TRUE_GM: the geometric mean (exponent 1/H)
    gm = torch.prod(torch.pow(p, 1 / H if flags & FLAG_TRUE_GM else 1 / (H * H)), dim=1)
    p = torch.div(p, gm.expand(B, H).view(B, H, 1))
    side = int(round(math.sqrt(H)))
    return p.view(B, 1, side, side), [float(r) for r in record], kstar


def retile_pages(pages: Sequence[torch.Tensor], flags: int = 0) -> torch.Tensor:
    """CP:218-238 `reconstruct`, bug included: every block-column repeats the
    vertical stack of pages[0:ratio]; pages >= ratio never reach the output.
    FLAG_CORRECT_TILING: page i*ratio+j goes to block (i, j), the inverse of split_matrix (CP:201-216)."""
    ratio = int(len(pages) ** 0.5)
    if flags & FLAG_CORRECT_TILING:
        return torch.cat([torch.cat(list(pages[i * ratio:(i + 1) * ratio]), 3) for i in range(ratio)], 2)
    col = torch.cat(list(pages[0:ratio]), 2)
    return torch.cat([col] * ratio, dim=3)


def relative_decoder_tail(x: torch.Tensor, books, want_intermediates: bool = False, force_k=None, flags: int = 0):
    """RN:358-396: non-DORN `Ordinal_Layer.forward` for a (B,1,s,s) f32 decoder
    map, s in {8,16,32,64,128}.  Returns the filled relative map (B,1,s,s) f32
    and, optionally, per-page intermediates (raw, bins, kstar, record)."""
    s = x.shape[2]
    q, lv = books[s]
    inter = []
    if s == 8:
        raw = pair_v1(x)
        vals, bins = lloyd(raw, q, lv)
        out, rec, k = als_rank1(vals, LIMIT_8, None if force_k is None else force_k[0], flags)
        inter.append(dict(raw=raw, bins=bins, kstar=k, record=rec, page=out))
    else:
        dn_1 = resize_half(x)
        outs = []
        for pi, (page, parent) in enumerate(split_pages(x, dn_1)):
            raw = pair_id(page, parent)
            vals, bins = lloyd(raw, q, lv)
            o, rec, k = als_rank1(vals, LIMIT_PAGE, None if force_k is None else force_k[pi], flags)
            outs.append(o)
            inter.append(dict(raw=raw, bins=bins, kstar=k, record=rec, page=o))
        out = outs[0] if s == 16 else retile_pages(outs, flags)
    return (out, inter) if want_intermediates else out


# --------------------------------------------------------------------------- stage 4: decomposition
def quick_gm(t: torch.Tensor, rc: int) -> torch.Tensor:
    """CP:244-255: prod_i t_i ** (1/rc**2) over dim 1."""
    return torch.prod(torch.pow(t, 1 / (rc * rc)), dim=1)


def gm_normalize(x: torch.Tensor) -> torch.Tensor:
    """RN:117 / MOD:145-149: x / quick_gm(x.view(B,HW,1), H)."""
    B, _, H, W = x.shape
    return torch.div(x, quick_gm(x.view(B, H * W, 1), H).expand(B, H * W).view(B, 1, H, W))


def upsample2(x: torch.Tensor) -> torch.Tensor:
    """CP:357-360: `.double()` + nearest x2."""
    return x.double().repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)


def decompose(dn: torch.Tensor, n: int, relative_map: bool = False) -> List[torch.Tensor]:
    """CP:368-392 followed by the callers' `[::-1]` (RN:117-122, MOD:123):
    returns [D_0 (unless relative_map), F_1, ..., F_n], F_k of side 2**k, f64."""
    comps = []
    cur = dn
    for k in range(n, 0, -1):
        nxt = resize_half(cur)
        comps.append(torch.div(cur, upsample2(nxt)))
        cur = nxt
    if not relative_map:
        comps.append(cur)
    return comps[::-1]


# --------------------------------------------------------------------------- stage 5: weighted reconstruction
def fine_detail_matrices(rows: Sequence[Sequence[torch.Tensor]]) -> List[torch.Tensor]:
    """CP:423-484: bucket by side length, log, stack in decoder order -> list of
    (B,K,M) f64, empty slots dropped."""
    slots: List[List[torch.Tensor]] = [[] for _ in range(8)]
    for row in rows:
        for comp in row:
            slots[int(math.log2(comp.shape[2]))].append(comp)
    out = []
    for cands in slots:
        if cands:
            B = cands[0].shape[0]
            out.append(torch.cat([torch.log(c).reshape(B, 1, -1) for c in cands], dim=1))
    return out


def make_pred(weights: Sequence[torch.Tensor], A: Sequence[torch.Tensor], relative_only: bool = False):
    """CP:512-528: per slot, per image `A[b].T.float() @ w.float()` -> (B,1,side,side) f32."""
    w = list(weights[1:]) if relative_only else list(weights)
    out = []
    for i, a in enumerate(A):
        B, _, M = a.shape
        tmp = torch.zeros(B, M, 1)
        for b in range(B):
            tmp[b] = torch.matmul(a[b].T.float(), w[i].float())
        side = int(math.sqrt(M))
        out.append(tmp.view(B, 1, side, side))
    return out


def recombination(comps: Sequence[torch.Tensor], n: int = 7) -> torch.Tensor:
    """CP:394-421: sum of nearest-upsampled components -> (B,1,2**n,2**n) f64."""
    def up(x, times):
        for _ in range(times):
            x = upsample2(x)
        return x.double() if times == 0 else x
    comps = list(comps)
    d0 = None
    if comps[0].shape[2] == 1:
        d0 = up(comps.pop(0), n)
    result = up(comps.pop(0), n - 1)
    for i, c in enumerate(comps):
        result = result + up(c, n - (i + 2))
    return result if d0 is None else d0 + result


# --------------------------------------------------------------------------- whole path
def fusion_forward(x_d1: torch.Tensor, rel_maps: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                   books=None, want_intermediates: bool = False, force_k=None, flags: int = 0):
    """The path RN:103-133 executes with decoder 1 plus relative decoders at the
    scales of `rel_maps` (8, 16, 32, ... in that order), followed by
    `recombination` (MOD:132).

    x_d1: (B,1,8,8) int64 DORN counts; rel_maps[k]: (B,1,s_k,s_k) f32 decoder
    outputs; weights: list of (K_i,1) f32, one per non-empty slot.
    Returns dict(rel=[filled maps], comps=[[...] per decoder], y_hat=[...], depth=(B,1,128,128) f64)."""
    books = books or load_codebooks()
    B = x_d1.shape[0]
    filled, inter = [], []
    for i, x in enumerate(rel_maps):
        o, it = relative_decoder_tail(x, books, want_intermediates=True, force_k=None if force_k is None else force_k[i], flags=flags)
        filled.append(o)
        inter.append(it)
    rows = [decompose(gm_normalize(x_d1), 3)]                       # RN:117
    for f in filled:
        rows.append(decompose(f, int(math.log2(f.shape[2])), relative_map=True))   # RN:119-122
    A = fine_detail_matrices(rows)                                  # RN:125
    y_hat = make_pred(weights, A)                                   # RN:133
    depth = recombination(y_hat)                                    # MOD:132
    out = dict(rel=filled, comps=rows, A=A, y_hat=y_hat, depth=depth)
    if want_intermediates:
        out["inter"] = inter
    return out


def slot_sizes(scales: Sequence[int]) -> List[int]:
    """Candidates per slot (K_i) for decoder 1 + relative decoders at `scales`
    (what `Weights(vector_sizes=...)` must be sized to, RN:63)."""
    K = [1, 1, 1, 1, 0, 0, 0, 0]
    for s in scales:
        for k in range(1, int(math.log2(s)) + 1):
            K[k] += 1
    return K


# --------------------------------------------------------------------------- training-side pieces
def mask_target(y: torch.Tensor) -> torch.Tensor:
    """MOD:74-78: +1e-4 on every pixel, invalid (<=0) pixels become 1.0001."""
    return (y * (y > 0)) + ((y <= 0) + 1e-4)


def depth2label_sid(depth: torch.Tensor, K: float = 90.0, alpha: float = 0.02, beta: float = 10.0):
    """utils.py:195-211."""
    a, b, k = torch.tensor(alpha), torch.tensor(beta), torch.tensor(K)
    label = k * torch.log(depth / a) / torch.log(b / a)
    return torch.max(label, torch.zeros(label.shape)).int()


def gt_components(target: torch.Tensor) -> List[torch.Tensor]:
    """MOD:123: decompose(normalize(target), 7)[::-1] on the masked 128x128 f64 GT."""
    return decompose(gm_normalize(target), 7)


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_batch(B: int, scales: Sequence[int], seed: int):
    """Seeded synthetic decoder outputs: x_d1 = randint(1,90), rel maps
    exp(0.3*randn), weights abs(randn(K,1)) (RN:449-465)."""
    g = torch.Generator().manual_seed(seed)
    x_d1 = torch.randint(1, 90, (B, 1, 8, 8), generator=g, dtype=torch.int64)
    rel = [torch.exp(0.3 * torch.randn(B, 1, s, s, generator=g)) for s in scales]
    Ks = [k for k in slot_sizes(scales) if k > 0]
    weights = [torch.abs(torch.randn(k, 1, generator=g)) for k in Ks]
    return x_d1, rel, weights


# --------------------------------------------------------------------------- training step (BASELINE config 3)
def training_targets(y128: torch.Tensor) -> List[torch.Tensor]:
    """MOD:119-127 `compute_final_depth` targets for the masked 128x128 f64 ground truth: components of
    the normalised target (n=7) with slot 0 replaced by D_0 of the SID-labelled 8x8 target."""
    comps = decompose(gm_normalize(y128), 7)
    ord8 = depth2label_sid(resize(y128, 8))                         # int32 labels, utils.py:195-211
    ord_comps = decompose(gm_normalize(ord8), 3)
    comps[0] = ord_comps[0]
    return comps


def training_loss(y_raw: torch.Tensor, y_hat: Sequence[torch.Tensor]):
    """MOD:64-92 without the ordinal loss: resize to 128, mask, per-scale MSE (detached sum, CP:499-510),
    recombination, MSE against the masked target.  Returns (loss_all, mse, fine_detail_loss, final_depth)."""
    y = mask_target(resize(y_raw, 128))                              # MOD:68, MOD:74-78
    targets = training_targets(y)
    fine = torch.sum(torch.as_tensor([torch.nn.MSELoss()(a, b) for a, b in zip(y_hat, targets)]))   # CP:499-510
    final = recombination(list(y_hat))                               # MOD:132
    mse = torch.nn.MSELoss()(final, y)                               # MOD:89
    return mse + fine, mse, fine, final


# --------------------------------------------------------------------------- SURVEY 8f "next": DORN head + ordinal loss
def dorn_regression(x: torch.Tensor):
    """RN:313-345.  x (N,2K,H,W) -> (decode (N,1,H,W) int64, ord (N,K,H,W) f64)."""
    N, C, H, W = x.size()
    K = C // 2
    A = x[:, ::2, :, :].clone().view(N, 1, K * H * W)
    B = x[:, 1::2, :, :].clone().view(N, 1, K * H * W)
    Cc = torch.clamp(torch.cat((A, B), dim=1), min=1e-8, max=1e4).double()
    ord_c1 = torch.nn.functional.softmax(Cc, dim=1)[:, 1, :].clone().view(-1, K, H, W)
    return torch.sum((ord_c1 > 0.5), dim=1).view(-1, 1, H, W), ord_c1


def ordinal_loss(ord_labels: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """loss.py:17-59 with the K-loop written as an arange (same index tensor)."""
    N, C, H, W = ord_labels.size()
    Kidx = torch.arange(C, dtype=torch.int).view(1, C, 1, 1).expand(N, C, H, W)
    mask_0 = (Kidx <= target).detach()
    mask_1 = (Kidx > target).detach()
    one = torch.ones(ord_labels[mask_1].size())
    loss = torch.sum(torch.log(torch.clamp(ord_labels[mask_0], min=1e-8, max=1e8).float())) \
        + torch.sum(torch.log(torch.clamp(one - ord_labels[mask_1], min=1e-8, max=1e8).float()))
    return loss / (-(N * H * W))
