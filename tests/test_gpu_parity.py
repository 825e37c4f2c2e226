"""GPU parity: every kernel through the C ABI (torch.ops.rdm.* / FusionPlan) against the CPU
oracle on the same seeded inputs, against the golden vectors generated from the unmodified
reference, and through size-independent properties.

Bars: Lloyd bins and raw pair matrices BIT-EXACT; k* equal; ALS maps / components
relative <= 1e-5; final log-depth max |ours - ref| <= 1e-4 * max(1, |ref|); Weights gradients
relative <= 1e-4.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import fusion_ref as fr

pytestmark = pytest.mark.gpu

REL_MAP = 1e-5
DEPTH_TOL = 1e-4


def _dev_books(books, dev):
    return {s: (q.to(dev), lv.to(dev)) for s, (q, lv) in books.items()}


def _eq_nan(a, b):
    return torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def _rel_err(a, b):
    return ((a - b).abs() / b.abs().clamp_min(1e-30)).max().item()


def _depth_ok(ours, ref):
    # NaN-aware: a DORN map with violent 1 <-> 89 jumps can make the (-3,19,19,-3)/32 bicubic go
    # negative, and log() of that is NaN in the reference too
    both_nan = torch.isnan(ours) & torch.isnan(ref)
    return bool((both_nan | ((ours - ref).abs() <= DEPTH_TOL * ref.abs().clamp_min(1.0))).all())


R = None


@pytest.fixture(scope="module", autouse=True)
def _ops(dev):
    global R
    import md_rdm_b200.ops  # noqa: F401  registers torch.ops.rdm
    R = torch.ops.rdm


# ============================================================================ stage 1
def test_pair_v1_bitexact(dev):
    g = torch.Generator().manual_seed(1)
    d = torch.exp(0.5 * torch.randn(16, 1, 8, 8, generator=g))
    out = R.pair_v1(d.to(dev)).cpu()
    assert torch.equal(out, fr.pair_v1(d))


@pytest.mark.parametrize("side", [2, 4, 8, 16, 32, 64, 128])
def test_resize_half(dev, side):
    g = torch.Generator().manual_seed(side)
    x = torch.rand(3, 1, side, side, generator=g) + 0.5
    assert torch.equal(R.resize_half(x.to(dev)).cpu(), fr.resize_half(x))            # f32-valued input: bit-equal
    xd = torch.rand(3, 1, side, side, generator=g, dtype=torch.float64) + 0.5
    assert _rel_err(R.resize_half(xd.to(dev)).cpu(), fr.resize_half(xd)) < 1e-15


def test_resize_bicubic_general(dev):
    g = torch.Generator().manual_seed(11)
    y = 0.5 + 9.5 * torch.rand(2, 1, 226, 226, generator=g, dtype=torch.float64)
    assert _rel_err(R.resize_bicubic(y.to(dev), 128, 128).cpu(), fr.resize(y, 128)) < 1e-13    # network/module.py:68
    z = fr.resize(y, 128)
    assert _rel_err(R.resize_bicubic(z.to(dev), 8, 8).cpu(), fr.resize(z, 8)) < 1e-13           # network/module.py:126


@pytest.mark.parametrize("side", [16, 32, 64])
def test_pair_id_bitexact(dev, side):
    g = torch.Generator().manual_seed(100 + side)
    x = torch.exp(0.3 * torch.randn(4, 1, side, side, generator=g))
    raw, parent = R.pair_id(x.to(dev))
    dn_1 = fr.resize_half(x)
    assert torch.equal(parent.cpu(), dn_1)
    for pi, (page, par) in enumerate(fr.split_pages(x, dn_1)):
        assert torch.equal(raw[:, pi].cpu(), fr.pair_id(page, par)), pi
        lit = R.pair_pages(page.contiguous().to(dev), par.contiguous().to(dev)).cpu()   # literal sparse_comparison_id signature
        assert torch.equal(lit, fr.pair_id(page, par))


def test_upsample_nearest(dev):
    x = torch.arange(32, dtype=torch.float32).view(2, 1, 4, 4)
    up = R.upsample_nearest(x.to(dev), 2).cpu()
    ref = fr.upsample2(fr.upsample2(x))
    assert up.dtype == torch.float64 and torch.equal(up, ref)


# ============================================================================ stage 2
def test_lloyd_golden_edges(dev, books):
    g = load_golden("lloyd_edges.npz")
    db = _dev_books(books, dev)
    for s in (8, 16, 32):
        for name in ("f32", "f64"):
            x = torch.from_numpy(g[f"x_{s}_{name}"])
            v, b = R.lloyd_quantize(x.to(dev), *db[s])
            assert torch.equal(b.cpu(), torch.from_numpy(g[f"bins_{s}_{name}"])), (s, name)
            assert _eq_nan(v.cpu(), torch.from_numpy(g[f"values_{s}_{name}"])), (s, name)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_lloyd_random_bitexact(dev, books, dtype):
    g = torch.Generator().manual_seed(7)
    db = _dev_books(books, dev)
    for s in (8, 16, 32, 64, 128):
        x = torch.exp(0.8 * torch.randn(1 << 20, generator=g)).to(dtype)
        q, lv = books[s]
        # plant values exactly on and one ulp around every threshold in this dtype
        qd = q.to(dtype)
        x[:40], x[40:80], x[80:120] = qd, torch.nextafter(qd, qd * 0), torch.nextafter(qd, qd * 2)
        v, b = R.lloyd_quantize(x.to(dev), *db[s])
        rv, rb = fr.lloyd(x, q, lv)
        assert torch.equal(b.cpu(), rb) and torch.equal(v.cpu(), rv), s
        # idempotence: every level lies inside its own bin
        v2, b2 = R.lloyd_quantize(v, *db[s])
        assert torch.equal(b2, b) and torch.equal(v2, v)


def test_lloyd_ragged_and_empty(dev, books):
    db = _dev_books(books, dev)
    for n in (0, 1, 3, 5, 1023):
        x = torch.exp(torch.randn(n, dtype=torch.float64))
        v, b = R.lloyd_quantize(x.to(dev), *db[16])
        rv, rb = fr.lloyd(x, *books[16])
        assert torch.equal(b.cpu(), rb) and torch.equal(v.cpu(), rv)


# ============================================================================ stage 3
def test_als_golden(dev):
    from md_rdm_b200 import _cabi
    g = load_golden("als.npz")
    for name, rows, side, lim in (("page", 256, 16, 100), ("sq", 64, 8, 30)):
        Rq = torch.from_numpy(g[f"Rq_{name}"])
        m, pages, rec, k, _, _ = R.als_rank1(Rq.to(dev), _cabi.SRC_VAL_F32, rows, side, lim, Rq.shape[0], None, None, False, False)
        assert int(k.item()) == int(g[f"kstar_{name}"])
        assert _rel_err(m.cpu(), torch.from_numpy(g[f"map_{name}"])) < REL_MAP
        assert np.allclose(rec.cpu().numpy().reshape(-1), g[f"record_{name}"], rtol=1e-5, atol=1e-7)
        ones = torch.ones(2, rows, 64)
        m, _, rec, k, _, _ = R.als_rank1(ones.to(dev), _cabi.SRC_VAL_F32, rows, side, lim, 2, None, None, False, False)
        assert int(k.item()) == 0 and float(rec.reshape(-1)[0]) == 0.0       # constant map: p = q = 1 is exact
        assert torch.equal(m.cpu(), torch.from_numpy(g[f"const_map_{name}"]))


def test_als_replay_path_matches_oracle(dev):
    """k* >= 2 (near-symmetric 64x64 matrices, where R.view(B,W,H) == R makes the iteration a
    power iteration that keeps improving) so phase 1 replays k* iterations from the source."""
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(5)
    found = False
    for trial in range(4):
        u = torch.exp(0.8 * torch.randn(4, 64, 1, generator=g))
        Rq = (u * u.transpose(1, 2) * torch.exp(0.02 * torch.randn(4, 64, 64, generator=g))).float()
        m_ref, rec_ref, k_ref = fr.als_rank1(Rq, 30)
        m, _, rec, k, _, _ = R.als_rank1(Rq.to(dev), _cabi.SRC_VAL_F32, 64, 8, 30, 4, None, None, False, False)
        assert np.allclose(rec.cpu().numpy().reshape(-1), np.array(rec_ref, dtype=np.float32), rtol=2e-5, atol=1e-7)
        srt = sorted(rec_ref)
        found |= int(k.item()) >= 2
        if srt[1] > srt[0] * (1 + 5e-5):          # arg-min not a near tie -> must agree
            assert int(k.item()) == k_ref
            assert _rel_err(m.cpu(), m_ref) < REL_MAP
        else:                                      # near tie between converged iterates: maps still agree
            assert _rel_err(m.cpu(), m_ref) < 1e-3
    assert found, "no trial exercised k* >= 2"


def test_als_f64_input_and_step(dev):
    import md_rdm_b200.computations as cp
    g = torch.Generator().manual_seed(8)
    Rq = torch.exp(0.2 * torch.randn(3, 256, 64, generator=g, dtype=torch.float64))
    ours = cp.alternating_least_squares(Rq.to(dev), 4, True, limit=100).cpu()
    ref, _, _ = fr.als_rank1(Rq, 100)
    assert ours.shape == (3, 1, 16, 16) and _rel_err(ours, ref) < REL_MAP
    q = torch.rand(3, 64, 1, generator=g) + 0.5
    step = cp.als_step(Rq.float().to(dev), q.to(dev), True).cpu()
    assert _rel_err(step, fr._ridge_step(Rq.float(), q)) < 1e-5


def test_als_compact_pages_vs_dense_and_oracle(dev, books):
    """Page pair matrices take the compact-form kernels (rdm_als_sparse.cu); anything else takes the dense
    kernel.  Same matrices through every entry: raw f64 (quantised on the fly, bins + values emitted),
    quantised f64 (structure found by the bit-wise check), quantised f32 (always dense), a group in which
    ONE matrix has one entry outside its window changed (that unit alone goes dense), all against the
    oracle."""
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(77)
    x = torch.exp(0.3 * torch.randn(3, 1, 32, 32, generator=g))
    q, lv = books[32]
    thr, lvl = _dev_books(books, dev)[32]
    dn1 = fr.resize_half(x)
    raw = torch.stack([fr.pair_id(pg, par) for pg, par in fr.split_pages(x, dn1)], 1)          # (3,4,256,64) f64
    vals, bins = fr.lloyd(raw, q, lv)
    ref = [fr.als_rank1(vals[:, pi], 100) for pi in range(4)]

    def check(out, tag):
        m, pages, rec, k, b, v = out
        for pi in range(4):
            assert int(k.reshape(-1)[pi]) == ref[pi][2], (tag, pi)
            assert np.allclose(rec[0, pi].cpu().numpy(), np.array(ref[pi][1], dtype=np.float32), rtol=1e-5, atol=1e-7), (tag, pi)
            assert _rel_err(pages[:, pi].cpu().view(-1), ref[pi][0].reshape(-1)) < REL_MAP, (tag, pi)
        return b, v

    b, v = check(R.als_rank1(raw.to(dev), _cabi.SRC_RAW_F64, 256, 32, 100, 3, thr, lvl, True, True), "raw_f64")
    assert torch.equal(b.cpu(), bins)
    assert torch.equal(v.cpu(), vals.float())
    check(R.als_rank1(vals.to(dev), _cabi.SRC_VAL_F64, 256, 32, 100, 3, None, None, False, False), "val_f64")
    check(R.als_rank1(vals.float().to(dev), _cabi.SRC_VAL_F32, 256, 32, 100, 3, None, None, False, False), "val_f32 (dense)")
    b, _ = check(R.als_rank1(x.to(dev), _cabi.SRC_MAP_F32, 256, 32, 100, 3, thr, lvl, True, False), "map")
    assert torch.equal(b.cpu(), bins)

    # one unit of the group loses the structure: column 0 of row 200 is outside that row's window
    mixed = vals.clone()
    assert not fr.window_mask(8, 16)[200, 0]
    mixed[1, 2, 200, 0] = float(lv[3])
    ref_mixed = fr.als_rank1(mixed[:, 2], 100)
    m, pages, rec, k, _, _ = R.als_rank1(mixed.to(dev), _cabi.SRC_VAL_F64, 256, 32, 100, 3, None, None, False, False)
    assert int(k.reshape(-1)[2]) == ref_mixed[2]
    assert np.allclose(rec[0, 2].cpu().numpy(), np.array(ref_mixed[1], dtype=np.float32), rtol=1e-5, atol=1e-7)
    assert _rel_err(pages[:, 2].cpu().view(-1), ref_mixed[0].reshape(-1)) < REL_MAP
    # ... and a window entry may hold anything (here: another level) without leaving the compact path
    assert fr.window_mask(8, 16)[0, 0]
    wdw = vals.clone()
    wdw[0, 0, 0, 0] = float(lv[30])
    ref_w = fr.als_rank1(wdw[:, 0], 100)
    _, pages, rec, k, _, _ = R.als_rank1(wdw.to(dev), _cabi.SRC_VAL_F64, 256, 32, 100, 3, None, None, False, False)
    assert int(k.reshape(-1)[0]) == ref_w[2]
    assert _rel_err(pages[:, 0].cpu().view(-1), ref_w[0].reshape(-1)) < REL_MAP


@pytest.mark.parametrize("s", [8, 16, 32])
def test_relative_tail_golden(dev, books, s):
    """Fused pair build + Lloyd + ALS (MAP source) against the reference's Ordinal_Layer.forward."""
    from md_rdm_b200 import _cabi
    g = load_golden("relative_tails_b2.npz")
    x = torch.from_numpy(g[f"x_{s}"])
    thr, lvl = _dev_books(books, dev)[s]
    rows, lim = (64, 30) if s == 8 else (256, 100)
    m, pages, rec, k, bins, vals = R.als_rank1(x.to(dev), _cabi.SRC_MAP_F32, rows, s, lim, x.shape[0], thr, lvl, True, True)
    P = 1 if s == 8 else (s // 16) ** 2
    for pi in range(P):
        assert torch.equal(bins[:, pi].cpu(), torch.from_numpy(g[f"bins_{s}_p{pi}"])), pi
        assert int(k.reshape(-1)[pi]) == int(g[f"kstar_{s}_p{pi}"])
        assert np.allclose(rec[0, pi].cpu().numpy(), g[f"record_{s}_p{pi}"], rtol=1e-5, atol=1e-7)
        assert _rel_err(pages[:, pi].cpu().view(-1), torch.from_numpy(g[f"page_{s}_p{pi}"]).view(-1)) < REL_MAP
    assert _rel_err(m.cpu(), torch.from_numpy(g[f"map_{s}"])) < REL_MAP
    lv32 = books[s][1].float()
    assert torch.equal(vals.cpu(), lv32[bins.cpu().long()])


def test_relative_tail_64_vs_oracle(dev, books):
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(64)
    x = torch.exp(0.3 * torch.randn(2, 1, 64, 64, generator=g))
    thr, lvl = _dev_books(books, dev)[64]
    m, pages, rec, k, bins, _ = R.als_rank1(x.to(dev), _cabi.SRC_MAP_F32, 256, 64, 100, 2, thr, lvl, True, False)
    ref, inter = fr.relative_decoder_tail(x, books, want_intermediates=True)
    for pi, it in enumerate(inter):
        assert torch.equal(bins[:, pi].cpu(), it["bins"])
        assert int(k.reshape(-1)[pi]) == it["kstar"]
    assert _rel_err(m.cpu(), ref) < REL_MAP


def test_group_argmin_is_batch_wide(dev, books):
    """CP:172-173: one k* per reference call.  Two groups of 2 images give the same result as two
    separate calls, and the record is the rmse over the group."""
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(21)
    x = torch.exp(0.3 * torch.randn(4, 1, 16, 16, generator=g))
    x[2:] = 1.0                                            # second group: constant maps -> k* = 0
    thr, lvl = _dev_books(books, dev)[16]
    m, _, rec, k, _, _ = R.als_rank1(x.to(dev), _cabi.SRC_MAP_F32, 256, 16, 100, 2, thr, lvl, False, False)
    assert k.cpu().view(-1).tolist() == [1, 0]
    for gi in range(2):
        ref, inter = fr.relative_decoder_tail(x[2 * gi:2 * gi + 2], books, want_intermediates=True)
        assert _rel_err(m[2 * gi:2 * gi + 2].cpu(), ref) < REL_MAP
        # Constant maps leave a residual of ~8e-4 on values of 1: there one f32 ulp of |p|^2 (summation
        # order inside the reference's own matmul) moves the rmse by ~1e-3 relative, so only rec[0] == 0
        # exactly (hence k* = 0) and rough agreement are meaningful for that group.
        rtol = 1e-5 if gi == 0 else 5e-3
        ours_rec = rec[gi, 0].cpu().numpy()
        assert np.allclose(ours_rec, np.array(inter[0]["record"], dtype=np.float32), rtol=rtol, atol=1e-7)
        if gi == 1:
            assert ours_rec[0] == 0.0 and (ours_rec[1:] > 0).all()


@pytest.mark.parametrize("s", [8, 16, 32])
def test_roughness_sweep_bins_and_kstar(dev, books, s):
    """Map roughness sigma from constant to violent (k* = 0 below the cross-over band
    1e-3 < sigma < 1e-2, 1 above it): bins bit-exact everywhere, k* equal - or, if it ever differs, a tie of
    the oracle's own record - and the per-page maps within 1e-5 where k* agrees.  (Inside the band the record
    itself is ill-conditioned: the residual is ~1e-5 of the energy, so one f32 ulp in an iterate moves the rmse
    by 1e-5..1e-4 relative; tools/parity_sweep.py prints the figures.)"""
    from md_rdm_b200 import _cabi
    thr, lvl = _dev_books(books, dev)[s]
    rows, lim = (64, 30) if s == 8 else (256, 100)
    for sigma in (0.0, 1e-3, 3e-3, 1e-2, 0.1, 1.0):
        g = torch.Generator().manual_seed(4242 + s + int(sigma * 1e6))
        x = torch.exp(sigma * torch.randn(4, 1, s, s, generator=g))
        _, pages, rec, k, bins, _ = R.als_rank1(x.to(dev), _cabi.SRC_MAP_F32, rows, s, lim, 4, thr, lvl, True, False)
        _, inter = fr.relative_decoder_tail(x, books, want_intermediates=True)
        for pi, it in enumerate(inter):
            assert torch.equal(bins[:, pi].cpu(), it["bins"]), (sigma, pi)
            ko, rr = int(k.reshape(-1)[pi]), np.array(it["record"], dtype=np.float64)
            if ko == it["kstar"]:
                assert _rel_err(pages[:, pi].cpu().view(-1), it["page"].reshape(-1)) < REL_MAP, (sigma, pi)
            else:
                assert rr[ko] <= rr.min() * (1 + 3e-6) + 1e-12, (sigma, pi, ko, it["kstar"])


# ============================================================================ stage 4
def test_quick_gm_and_normalize(dev):
    g = torch.Generator().manual_seed(2)
    xi = torch.randint(1, 90, (5, 1, 8, 8), generator=g, dtype=torch.int64)
    ours = R.gm_normalize(xi.to(dev)).cpu()
    ref = fr.gm_normalize(xi)
    assert ours.dtype == ref.dtype == torch.float32 and _rel_err(ours, ref) < 5e-6
    gm = R.quick_gm(xi.view(5, 64, 1).to(dev), 8).cpu()
    assert _rel_err(gm, fr.quick_gm(xi.view(5, 64, 1), 8)) < 5e-6
    y = fr.mask_target(0.5 + 9.5 * torch.rand(2, 1, 128, 128, generator=g, dtype=torch.float64))
    assert _rel_err(R.gm_normalize(y.to(dev)).cpu(), fr.gm_normalize(y)) < 1e-12


@pytest.mark.parametrize("side,rel", [(8, False), (8, True), (16, True), (32, True), (64, True), (128, False)])
def test_decompose_vs_oracle(dev, side, rel):
    from md_rdm_b200.ops import unpack_pyramid
    g = torch.Generator().manual_seed(side)
    x = torch.exp(0.3 * torch.randn(3, 1, side, side, generator=g))
    if side == 128:
        x = x.double()
    packed = R.decompose(x.to(dev), rel)
    ours = unpack_pyramid(packed, 3, side, rel)
    ref = fr.decompose(x, int(math.log2(side)), relative_map=rel)
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert a.shape == b.shape and a.dtype == torch.float64
        assert _rel_err(a.cpu(), b) < 1e-13


def test_decompose_128_many_images_no_interband_race(dev):
    """Regression: the 128x128 decomposition bands its top level over many CTAs; a second banded launch used to
    read D_6 from the slot its own CTAs were overwriting with F_6 (halo rows of a band belong to the neighbour's
    CTA), which showed on some boxes as a 4e-5 error in D_0.  Many images keep every SM busy with bands."""
    from md_rdm_b200.ops import unpack_pyramid
    g = torch.Generator().manual_seed(128128)
    x = torch.exp(0.3 * torch.randn(96, 1, 128, 128, generator=g, dtype=torch.float64))
    ref = fr.decompose(x, 7, relative_map=False)
    for _ in range(3):
        ours = unpack_pyramid(R.decompose(x.to(dev), False), 96, 128, False)
        for a, b in zip(ours, ref):
            assert _rel_err(a.cpu(), b) < 1e-13


def test_gt_decompose_golden(dev):
    import md_rdm_b200.computations as cp
    g = load_golden("gt_decompose_b2.npz")
    y = torch.from_numpy(g["y_masked"]).to(dev)
    B = y.shape[0]
    norm = torch.div(y, cp.quick_gm(y.view(B, 128 * 128, 1), 128).expand(B, 128 * 128).view(B, 1, 128, 128))   # network/module.py:145-149
    comps = cp.decompose_depth_map([], norm, 7)[::-1]                                                                # network/module.py:123
    assert len(comps) == 8
    for i, c in enumerate(comps):
        assert _rel_err(c.cpu(), torch.from_numpy(g[f"comp_{i}"])) < 1e-11, i


def test_decompose_recombination_roundtrip_fullsize(dev):
    """recombination(log(decompose(d, 7))) == log d, batch 16, 128x128."""
    import md_rdm_b200.computations as cp
    g = torch.Generator().manual_seed(4)
    d = torch.exp(0.4 * torch.randn(16, 1, 128, 128, generator=g, dtype=torch.float64)).to(dev)
    comps = cp.decompose_depth_map([], d, 7)[::-1]
    logs = [torch.log(c) for c in comps]
    rt = cp.recombination(logs)
    assert logs == [] or len(logs) == 6            # the reference pops d_0 and f_1 from the caller's list
    assert (rt - torch.log(d)).abs().max().item() < 1e-12


# ============================================================================ stage 5
def test_log_stack_make_pred_recombination(dev):
    import md_rdm_b200.computations as cp
    g = torch.Generator().manual_seed(6)
    B = 4
    cands = [torch.exp(0.2 * torch.randn(B, 1, 8, 8, generator=g, dtype=torch.float64)) for _ in range(3)]
    A = cp.make_matrix([c.to(dev) for c in cands], True)
    A_ref = fr.fine_detail_matrices([cands])[0]
    assert torch.allclose(A.cpu(), A_ref, rtol=0, atol=1e-15)
    w = torch.abs(torch.randn(3, 1, generator=g))
    pred = cp.make_pred([w.to(dev)], [A.clone()], True, False)[0].cpu()
    ref = fr.make_pred([w], [A_ref])[0]
    assert pred.shape == (B, 1, 8, 8) and pred.dtype == torch.float32
    assert (pred - ref).abs().max().item() < 1e-6
    comps = [torch.randn(B, 1, 2 ** k, 2 ** k, generator=g) for k in range(0, 6)]
    for lst in (comps, comps[1:]):                      # with and without the 1x1 component
        ours = cp.recombination([c.to(dev) for c in lst]).cpu()
        assert torch.equal(ours, fr.recombination(lst))         # exact: same f64 adds in the same order
    c64 = [c.double() for c in comps]
    assert torch.equal(cp.recombination([c.to(dev) for c in c64]).cpu(), fr.recombination(c64))


def test_backward_make_pred_and_recombination(dev):
    import md_rdm_b200.computations as cp
    g = torch.Generator().manual_seed(12)
    B = 3
    A_list = [torch.randn(B, K, 4 ** k, generator=g, dtype=torch.float64) for k, K in enumerate((1, 3, 3, 2))]
    ws = [torch.abs(torch.randn(K, 1, generator=g)) for K in (1, 3, 3, 2)]
    # reference autograd on CPU
    w_ref = [w.clone().requires_grad_(True) for w in ws]
    A_ref = [a.clone().requires_grad_(True) for a in A_list]
    loss_ref = (fr.recombination(fr.make_pred(w_ref, A_ref)) ** 2).mean()
    loss_ref.backward()
    w_gpu = [w.clone().to(dev).requires_grad_(True) for w in ws]
    A_gpu = [a.clone().to(dev).requires_grad_(True) for a in A_list]
    loss = (cp.recombination(cp.make_pred(w_gpu, list(A_gpu), True, False)) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item())
    for a, b in zip(w_gpu, w_ref):
        assert _rel_err(a.grad.cpu(), b.grad) < 1e-4
    for a, b in zip(A_gpu, A_ref):
        assert a.grad.dtype == torch.float64
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-9)


# ============================================================================ whole path
def _run_plan(dev, x_d1, rel, weights, source, **kw):
    from md_rdm_b200.fusion import FusionPlan
    scales = tuple(r.shape[2] for r in rel)
    plan = FusionPlan(x_d1.shape[0], scales, source, device=dev, **kw)
    if source == "map":
        srcs = rel
    else:
        srcs = [R.pair_v1(r.to(dev)) if r.shape[2] == 8 else R.pair_id(r.to(dev))[0] for r in rel]
    plan.load_inputs(x_d1.to(dev), [s.to(dev) for s in srcs], torch.cat([w.reshape(-1) for w in weights]).to(dev))
    plan.run()
    torch.cuda.synchronize()
    return plan


def test_full_path_golden(dev, books):
    g = load_golden("full_path_b2.npz")
    scales = (8, 16, 32)
    x_d1 = torch.from_numpy(g["x_d1"])
    rel = [torch.from_numpy(g[f"rel_in_{s}"]) for s in scales]
    weights = [torch.from_numpy(g[f"w_{i}"]) for i in range(6)]
    for source in ("map", "raw"):
        plan = _run_plan(dev, x_d1, rel, weights, source, want_A=True)
        for s in scales:
            assert _rel_err(plan.rel[s].cpu(), torch.from_numpy(g[f"rel_out_{s}"])) < REL_MAP, (source, s)
        for i, a in enumerate(plan.A):
            assert torch.allclose(a.cpu(), torch.from_numpy(g[f"A_{i}"]), rtol=0, atol=2e-5), (source, i)
        for i, y in enumerate(plan.yhat_list()):
            assert torch.allclose(y.cpu(), torch.from_numpy(g[f"yhat_{i}"]), rtol=0, atol=5e-5), (source, i)
        assert _depth_ok(plan.depth.cpu(), torch.from_numpy(g["depth"])), source


def test_full_path_b16_vs_oracle_and_sources_agree(dev, books):
    """BASELINE config 2 at full batch: 16 images, scales 8/16/32, both input forms."""
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(16, scales, seed=1234)
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True)
    pm = _run_plan(dev, x_d1, rel, weights, "map")
    pr = _run_plan(dev, x_d1, rel, weights, "raw")
    for si, s in enumerate(scales):
        for pi, it in enumerate(ref["inter"][si]):
            assert torch.equal(pm.bins[s][:, pi].cpu(), it["bins"]), (s, pi)          # bit-exact bins
            assert int(pm.kstar[s].view(-1)[pi]) == it["kstar"]
        assert torch.equal(pm.bins[s], pr.bins[s])
        assert torch.equal(pm.rel[s], pr.rel[s])                                      # same arithmetic either way
        assert _rel_err(pm.rel[s].cpu(), ref["rel"][si]) < REL_MAP
    assert torch.equal(pm.depth, pr.depth)
    assert _depth_ok(pm.depth.cpu(), ref["depth"])
    for y, yr in zip(pm.yhat_list(), ref["y_hat"]):
        assert torch.allclose(y.cpu(), yr, rtol=0, atol=5e-5)


def test_full_path_with_64_scale(dev, books):
    scales = (8, 16, 32, 64)                      # the configuration RN:96-97 names (decoders 1,6,7,8,9)
    x_d1, rel, weights = fr.synthetic_batch(2, scales, seed=99)
    ref = fr.fusion_forward(x_d1, rel, weights, books)
    plan = _run_plan(dev, x_d1, rel, weights, "map")
    assert _depth_ok(plan.depth.cpu(), ref["depth"])
    assert plan.kmax == 6 and plan.n_weights == sum(fr.slot_sizes(scales))


def test_cuda_graph_replay_equals_eager(dev, books):
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(4, scales, seed=5)
    plan = _run_plan(dev, x_d1, rel, weights, "map")
    eager = plan.depth.clone()
    plan.depth.zero_()
    plan.replay()
    torch.cuda.synchronize()
    assert torch.equal(plan.depth, eager)
    host = plan.run_host(x_d1, rel)
    assert torch.equal(host, eager.cpu())


def test_plan_overlap_and_multi_group_equal_plain_runs(dev, books):
    """FusionPlan(overlap=True) (dense 8x8 ALS forked beside the page kernels) and a plan that carries two
    reference batches per launch (group = 8 of N = 16: one arg-min per group, CP:172-173) give bit-identical
    results to plain single-batch runs, eagerly and from a CUDA graph."""
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(16, scales, seed=31)
    rel[1][8:] = 1.0                               # second group: constant 16x16 maps -> a different k*
    w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
    halves = []
    for lo in (0, 8):
        p = FusionPlan(8, scales, "map", device=dev)
        p.load_inputs(x_d1[lo:lo + 8].to(dev), [r[lo:lo + 8].to(dev) for r in rel], w)
        p.run()
        torch.cuda.synchronize()
        halves.append((p.depth.clone(), {s: p.kstar[s].clone() for s in scales}))
    assert halves[0][1][16].view(-1).tolist() != halves[1][1][16].view(-1).tolist()
    for overlap in (False, True):
        plan = FusionPlan(16, scales, "map", group=8, device=dev, overlap=overlap)
        plan.load_inputs(x_d1.to(dev), [r.to(dev) for r in rel], w)
        plan.run()
        torch.cuda.synchronize()
        for gi in range(2):
            assert torch.equal(plan.depth[8 * gi:8 * gi + 8], halves[gi][0]), (overlap, gi)
            for s in scales:
                assert torch.equal(plan.kstar[s][gi], halves[gi][1][s][0]), (overlap, gi, s)
        eager = plan.depth.clone()
        plan.depth.zero_()
        plan.replay()
        torch.cuda.synchronize()
        assert torch.equal(plan.depth, eager), overlap


def test_weights_gradient_golden(dev):
    """the only gradient the training loss needs (loss = mean(depth^2))."""
    from md_rdm_b200.ops import fuse_tail_autograd
    g = load_golden("full_path_b2.npz")
    x_d1 = torch.from_numpy(g["x_d1"]).to(dev)
    rel = [torch.from_numpy(g[f"rel_out_{s}"]).to(dev) for s in (8, 16, 32)]
    ws = [torch.from_numpy(g[f"w_{i}"]) for i in range(6)]
    flat = torch.cat([w.reshape(-1) for w in ws]).to(dev).requires_grad_(True)
    depth, yhat = fuse_tail_autograd(x_d1, rel, flat)
    loss = (depth ** 2).mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    ref = torch.cat([torch.from_numpy(g[f"grad_w_{i}"]).reshape(-1) for i in range(6)])
    assert _rel_err(flat.grad.cpu(), ref) < 1e-4


def test_dropin_composition_matches_oracle(dev, books):
    """The literal call sequence of RN:103-133 + network/module.py:132 written against the drop-in
    names, including zero (not None) gradients upstream of Lloyd and real gradients to Weights."""
    import md_rdm_b200.computations as cp
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization, Weights
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(3, scales, seed=77)
    ref = fr.fusion_forward(x_d1, rel, weights, books)
    quant = Quantization()
    layers = [Ordinal_Layer(int(math.log2(s)) + 3, False, quant) for s in scales]
    xd = x_d1.to(dev)
    rel_in = [r.to(dev).requires_grad_(True) for r in rel]
    x_rel = [layer(r) for layer, r in zip(layers, rel_in)]
    B, C, H, W = xd.size()
    f_d1 = cp.decompose_depth_map([], torch.div(xd, cp.quick_gm(xd.view(B, H * W, 1), H).expand(B, H * W).view(B, 1, H, W)), 3)[::-1]
    rows = [f_d1] + [cp.decompose_depth_map([], r, int(math.log2(r.shape[2])), relative_map=True)[::-1] for r in x_rel]
    y_hat = cp.relative_fine_detail_matrix(rows, True)
    wl = Weights(vector_sizes=fr.slot_sizes(scales), use_cuda=True, relative_only=False)
    with torch.no_grad():
        for p_, w_ in zip(wl.weight_list, weights):
            p_.copy_(w_)
    y_hat = wl(y_hat)
    keep = [y.detach().cpu() for y in y_hat]
    depth = cp.recombination(y_hat)
    for r, rr in zip(x_rel, ref["rel"]):
        assert _rel_err(r.detach().cpu(), rr) < REL_MAP
    for a, b in zip(keep, ref["y_hat"]):
        assert torch.allclose(a, b, rtol=0, atol=5e-5)
    assert _depth_ok(depth.detach().cpu(), ref["depth"])
    (depth ** 2).mean().backward()
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in wl.weight_list if p.numel())
    # the Ordinal_Layer kernel chain returns zero gradients; the decompose op itself has no
    # backward registered, so nothing reaches rel_in: autograd reports None or zeros there
    for r in rel_in:
        assert r.grad is None or float(r.grad.abs().sum()) == 0.0


def test_ordinal_layer_methods(dev, books):
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization
    quant = Quantization()
    g = torch.Generator().manual_seed(31)
    d3 = torch.exp(0.3 * torch.randn(2, 1, 8, 8, generator=g))
    l6 = Ordinal_Layer(6, False, quant)
    q8 = l6.sparse_comparison_v1(d3.to(dev)).cpu()
    assert torch.equal(q8, fr.lloyd(fr.pair_v1(d3), *books[8])[0])
    x = torch.exp(0.3 * torch.randn(2, 1, 16, 16, generator=g))
    l7 = Ordinal_Layer(7, False, quant)
    dn_1 = fr.resize_half(x)
    q16 = l7.sparse_comparison_id(x.to(dev), dn_1.to(dev)).cpu()
    assert q16.dtype == torch.float64 and torch.equal(q16, fr.lloyd(fr.pair_id(x, dn_1), *books[16])[0])
    # in-place semantics of LloydQuantization on a contiguous tensor (RN:292-297)
    raw = fr.pair_v1(d3).to(dev)
    out = l6.LloydQuantization(torch.empty(2, 64, 64, 0), raw)
    assert out.data_ptr() == raw.data_ptr() and torch.equal(raw.cpu(), q8)


def test_edge_batches(dev, books):
    from md_rdm_b200.fusion import FusionPlan
    # B = 1
    x_d1, rel, weights = fr.synthetic_batch(1, (8, 16), seed=3)
    ref = fr.fusion_forward(x_d1, rel, weights, books)
    plan = _run_plan(dev, x_d1, rel, weights, "map")
    assert _depth_ok(plan.depth.cpu(), ref["depth"])
    # no relative decoders at all: decoder 1 only (the reference's HEAD configuration, RN:63)
    plan = FusionPlan(2, (), "map", device=dev)
    x = torch.randint(1, 90, (2, 1, 8, 8), dtype=torch.int64, generator=torch.Generator().manual_seed(17))
    w = torch.tensor([0.7, 1.1, 0.4, 0.9])
    plan.load_inputs(x.to(dev), [], w.to(dev))
    plan.run()
    rows = [fr.decompose(fr.gm_normalize(x), 3)]
    refd = fr.recombination(fr.make_pred([w[i:i + 1].view(1, 1) for i in range(4)], fr.fine_detail_matrices(rows)))
    assert _depth_ok(plan.depth.cpu(), refd)


@pytest.mark.parametrize("side,dtype", [(16, torch.float64), (32, torch.float32), (128, torch.float64)])
def test_backward_decompose_chain(dev, side, dtype):
    """gm_normalize -> decompose -> log_stack backward kernels against torch autograd on the oracle."""
    import md_rdm_b200.computations as cp
    g = torch.Generator().manual_seed(side)
    n = int(math.log2(side))
    x = torch.exp(0.3 * torch.randn(2, 1, side, side, generator=g, dtype=torch.float64)).to(dtype)
    coef = [torch.randn(2, 1, 2 ** k, 2 ** k, generator=g, dtype=torch.float64) for k in range(n + 1)]
    xr = x.clone().requires_grad_(True)
    comps = fr.decompose(fr.gm_normalize(xr), n)
    loss_ref = sum((torch.log(c) * w).sum() for c, w in zip(comps, coef))
    loss_ref.backward()
    xg = x.clone().to(dev).requires_grad_(True)
    comps_g = cp.decompose_depth_map([], R.gm_normalize(xg), n)[::-1]
    loss = sum((cp.make_matrix([c], True).view(c.shape) * w.to(dev)).sum() for c, w in zip(comps_g, coef))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-6 * max(1.0, abs(loss_ref.item()))
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    assert torch.allclose(xg.grad.cpu().double(), xr.grad.double(), rtol=tol, atol=tol * xr.grad.abs().max().item())


def test_training_step_config3(dev, books):
    """BASELINE config 3: ground-truth preparation + decomposition (MOD:68-78, 119-133, 145-149), fusion
    forward, the module's loss, backward to the Weights parameters - written against the drop-in names."""
    import md_rdm_b200.computations as cp
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization, Weights
    scales, B = (8, 16, 32), 4
    x_d1, rel, weights = fr.synthetic_batch(B, scales, seed=303)
    g = torch.Generator().manual_seed(304)
    y_raw = 0.5 + 9.5 * torch.rand(B, 1, 226, 226, generator=g, dtype=torch.float64)
    y_raw = y_raw * (torch.rand(B, 1, 226, 226, generator=g) > 0.05)         # 5 % invalid pixels
    # ---- oracle (CPU, torch autograd)
    w_ref = [w.clone().requires_grad_(True) for w in weights]
    fwd = fr.fusion_forward(x_d1, rel, w_ref, books)
    loss_ref, mse_ref, fine_ref, final_ref = fr.training_loss(y_raw, fwd["y_hat"])
    loss_ref.backward()
    # ---- product (GPU), the reference's own call sequence
    quant = Quantization()
    xd = x_d1.to(dev)
    x_rel = [Ordinal_Layer(int(math.log2(s)) + 3, False, quant)(r.to(dev)) for s, r in zip(scales, rel)]
    Bn, C, H, W = xd.size()
    f_d1 = cp.decompose_depth_map([], torch.div(xd, cp.quick_gm(xd.view(Bn, H * W, 1), H).expand(Bn, H * W).view(Bn, 1, H, W)), 3)[::-1]
    rows = [f_d1] + [cp.decompose_depth_map([], r, int(math.log2(r.shape[2])), relative_map=True)[::-1] for r in x_rel]
    wl = Weights(vector_sizes=fr.slot_sizes(scales), use_cuda=True, relative_only=False)
    with torch.no_grad():
        for p_, w_ in zip(wl.weight_list, weights):
            p_.copy_(w_)
    y_hat = wl(cp.relative_fine_detail_matrix(rows, True))

    def normalize(batch):                                            # MOD:145-149
        b, c, h, w = batch.size()
        return torch.div(batch, cp.quick_gm(batch.view(b, h * w, 1), h).expand(b, h * w).view(b, 1, h, w))

    def depth2label_sid(depth, K=90.0, alpha=0.02, beta=10.0):       # utils.py:195-211 (caller glue, out of the path)
        a, b_, k = torch.tensor(alpha, device=dev), torch.tensor(beta, device=dev), torch.tensor(K, device=dev)
        label = k * torch.log(depth / a) / torch.log(b_ / a)
        return torch.max(label, torch.zeros(label.shape, device=dev)).int()

    y = cp.resize(y_raw.to(dev), 128)                                # MOD:68
    y = (y * (y > 0)) + ((y <= 0) + 1e-4)                            # MOD:74-78
    component_target = cp.decompose_depth_map([], normalize(y), 7)[::-1]                                  # MOD:123
    ord_components = cp.decompose_depth_map([], normalize(depth2label_sid(cp.resize(y, 8))), 3)[::-1]    # MOD:126
    component_target[0] = ord_components[0]
    components, fine = cp.optimize_components(y_hat, component_target, True)                               # MOD:130
    final = cp.recombination(components)                                                                   # MOD:132
    mse = torch.nn.MSELoss()(final, y)                                                                     # MOD:89
    loss = mse + fine
    loss.backward()
    assert _depth_ok(final.detach().cpu(), final_ref.detach())
    assert abs(fine.item() - fine_ref.item()) <= 1e-5 * abs(fine_ref.item())
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    for p_, r_ in zip(wl.weight_list, w_ref):
        if p_.numel():
            assert _rel_err(p_.grad.cpu(), r_.grad) < 1e-4


def test_kitti_shaped_tiles_config5(dev, books):
    """BASELINE config 5: a 228x912 KITTI input gives 8x29 coarse maps; the reference path is
    square / power-of-two only, so the stress is defined as 4 square tiles per image, tile-major, ONE
    arg-min group per call (B*4 images).  Batch 16 -> 64 tiles, scales 8/16/32."""
    scales, B, T = (8, 16, 32), 16, 4
    x_d1, rel, weights = fr.synthetic_batch(B * T, scales, seed=555)
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True)
    plan = _run_plan(dev, x_d1, rel, weights, "map")
    for si, s in enumerate(scales):
        for pi, it in enumerate(ref["inter"][si]):
            assert torch.equal(plan.bins[s][:, pi].cpu(), it["bins"])
            assert int(plan.kstar[s].view(-1)[pi]) == it["kstar"]
    assert _depth_ok(plan.depth.cpu(), ref["depth"])
    # tiles of one image side by side: the 128x512 log-depth panorama of the stress definition
    pano = plan.depth.view(B, T, 128, 128).permute(0, 2, 1, 3).reshape(B, 128, T * 128)
    assert pano.shape == (16, 128, 512)
    assert _eq_nan(pano[:, :, 128:256], plan.depth.view(B, T, 128, 128)[:, 1])


def test_full_model_config1_golden(dev, books):
    """BASELINE config 1: what the reference's full model (CNN included, decoders 1,6,7,8,9, batch 1) feeds
    into and gets out of the fusion path, against the CUDA path (fused plan and drop-in names).

    Real decoder outputs are smooth, so the ALS record PLATEAUS: from iteration ~3 on its f32 values differ
    by one ulp, and which of them is the first minimum is decided by summation-order noise (the reference
    itself would pick another index with another BLAS).  Parity is therefore stated as: bins bit-exact;
    record equal to 1e-5; our k* is a tie of the reference's own record (within 3e-6 relative of its
    minimum, the noise level of an f32 record); and maps, y_hat, log-depth match the reference algorithm evaluated at that k*."""
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization
    g = load_golden("full_model_b1.npz")
    scales = (8, 16, 32, 64)
    x_d1 = torch.from_numpy(g["x_d1"])
    rel = [torch.from_numpy(g[f"rel_in_{s}"]) for s in scales]          # real decoder outputs: partly negative
    weights = [torch.from_numpy(g[f"w_{i}"]) for i in range(7)]
    plan = _run_plan(dev, x_d1, rel, weights, "map")
    # the oracle at the reference's own k* (stored with the golden) reproduces the reference's outputs
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True, force_k=[g[f"kstar_{s}"].tolist() for s in scales])
    assert (ref["depth"] - torch.from_numpy(g["depth"])).abs().max().item() <= 1e-5   # CPU-to-CPU f32 noise
    ks = []
    for si, s in enumerate(scales):
        ours_k = plan.kstar[s].view(-1).tolist()
        ks.append(ours_k)
        for pi, it in enumerate(ref["inter"][si]):
            assert torch.equal(plan.bins[s][:, pi].cpu(), it["bins"]), (s, pi)
            rec_ref = g[f"record_{s}"][pi]                       # the reference's record
            assert np.allclose(plan.record[s][0, pi].cpu().numpy(), rec_ref, rtol=1e-5, atol=1e-8), (s, pi)
            assert rec_ref[ours_k[pi]] <= rec_ref.min() * (1 + 1e-6), (s, pi, ours_k[pi], int(g[f"kstar_{s}"][pi]))
    # what the plateau costs in DIRECT terms (no force_k): printed (pytest -s / tools/kstar_report.py keeps the log under
    # profiles/), and bounded: the quirky gm^(1/H) normaliser (CP:244-255) makes the emitted map scale dependent, so
    # a different tie of the arg-min moves the log-depth by a few 1e-3 - the reference itself has that freedom
    direct = (plan.depth.cpu() - torch.from_numpy(g["depth"])).abs().max().item()
    print(f"\nconfig-1 golden: GPU k* per scale {ks}, reference k* {[g[f'kstar_{s}'].tolist() for s in scales]}, "
          f"direct max |log-depth - golden| = {direct:.3e}")
    assert direct <= 2e-2
    forced = fr.fusion_forward(x_d1, rel, weights, books, force_k=ks)
    for si, s in enumerate(scales):
        assert _rel_err(plan.rel[s].cpu(), forced["rel"][si]) < REL_MAP, s
    for y, yr in zip(plan.yhat_list(), forced["y_hat"]):
        assert torch.allclose(y.cpu(), yr, rtol=0, atol=5e-5)
    assert _depth_ok(plan.depth.cpu(), forced["depth"])
    # scale 8 (k* = 1, no plateau) also matches the golden directly, through the drop-in class
    out = Ordinal_Layer(6, False, Quantization())(rel[0].to(dev))
    assert _rel_err(out.cpu(), torch.from_numpy(g["rel_out_8"])) < REL_MAP


def test_dorn_regression_and_ordinal_loss_golden(dev):
    """SURVEY 8f "next" row: DORN head (RN:313-345) + Ordinal_Loss (loss.py:17-59), forward and backward,
    against the reference's own outputs and autograd gradient."""
    from md_rdm_b200.loss import Ordinal_Loss, depth2label_sid
    from md_rdm_b200.rdm_net import Ordinal_Layer
    g = load_golden("dorn_loss.npz")
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    decode, ord_ = Ordinal_Layer(1, True, None)(x)
    assert decode.dtype == torch.int64 and torch.equal(decode.cpu(), torch.from_numpy(g["decode"]))   # integer output: exact
    assert ord_.dtype == torch.float64 and torch.allclose(ord_.detach().cpu(), torch.from_numpy(g["ord"]), rtol=1e-14, atol=1e-300)
    target = torch.from_numpy(g["target"]).to(dev)
    loss = Ordinal_Loss().calc(ord_, target, cuda=True)
    assert loss.dtype == torch.float32 and abs(loss.item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    loss.backward()
    assert torch.allclose(x.grad.cpu(), torch.from_numpy(g["grad_x"]), rtol=1e-4, atol=1e-9)
    # SID labels of a depth map (utils.py:195-211) on the device
    gen = torch.Generator().manual_seed(808)
    torch.randn(3, 180, 8, 8, generator=gen)
    depth = 0.5 + 9.5 * torch.rand(3, 1, 8, 8, generator=gen, dtype=torch.float64)
    assert torch.equal(depth2label_sid(depth.to(dev)).cpu(), torch.from_numpy(g["target"]))


def test_relative_tail_128_and_fuse_maps(dev, books):
    """Decoder 10 (128x128, 64 pages, RN:383-396) through the drop-in class, and the functional fast path."""
    from md_rdm_b200.fusion import fuse_maps
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization
    g = torch.Generator().manual_seed(128)
    x = torch.exp(0.3 * torch.randn(1, 1, 128, 128, generator=g))
    out = Ordinal_Layer(10, False, Quantization())(x.to(dev))
    ref = fr.relative_decoder_tail(x, books)
    assert out.shape == (1, 1, 128, 128) and _rel_err(out.cpu(), ref) < REL_MAP
    x_d1, rel, weights = fr.synthetic_batch(3, (8, 16), seed=42)
    depth, y_hat, filled = fuse_maps(x_d1.to(dev), [r.to(dev) for r in rel], [w.to(dev) for w in weights])
    o = fr.fusion_forward(x_d1, rel, weights, books)
    assert _depth_ok(depth.cpu(), o["depth"]) and len(y_hat) == 5
    for a, b in zip(filled, o["rel"]):
        assert _rel_err(a.cpu(), b) < REL_MAP


def test_codebook_fallbacks(dev, books):
    """Arbitrary caller codebooks: unsorted thresholds (the reference's 40 compares do not care about order)
    take the literal-compare path; thresholds closer than a lookup cell take the binary search.  Both must
    still give the reference's bins, stand-alone and inside the fused ALS kernel."""
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(77)
    q, lv = books[16]
    x = torch.exp(0.5 * torch.randn(4, 256, 64, generator=g, dtype=torch.float64))
    perm = torch.randperm(40, generator=g)
    tight = q.clone()
    tight[21] = tight[20] * (1 + 1e-6)          # two thresholds inside one 2^-10 cell
    tight, _ = torch.sort(tight)
    for thr in (q[perm], tight):
        rv, rb = fr.lloyd(x, thr, lv)
        v, b = R.lloyd_quantize(x.to(dev), thr.to(dev), lv.to(dev))
        assert torch.equal(b.cpu(), rb) and torch.equal(v.cpu(), rv)
        xf = x.float()
        rvf, rbf = fr.lloyd(xf, thr, lv)
        vf, bf = R.lloyd_quantize(xf.to(dev), thr.to(dev), lv.to(dev))
        assert torch.equal(bf.cpu(), rbf) and torch.equal(vf.cpu(), rvf)
        m, _, _, k, bins, _ = R.als_rank1(x.to(dev), _cabi.SRC_RAW_F64, 256, 16, 100, 4, thr.to(dev), lv.to(dev), True, False)
        assert torch.equal(bins.view(4, 256, 64).cpu(), rb)
        ref_map, _, k_ref = fr.als_rank1(rv, 100)
        assert int(k.item()) == k_ref and _rel_err(m.cpu(), ref_map) < REL_MAP


def test_sparsify_threshold_search_adversarial(dev, books):
    """Pair-structured matrices whose fill and window values sit ON the Lloyd thresholds, one f64 ulp either side, on
    their f32 roundings and one ulp either side of those (plus zero and values beyond the f32 range) must give the
    reference's bins bit for bit through the sparsify kernel - with the shipped table, with two thresholds closer than
    an f32 ulp and with an unsorted table (the literal 40 compares).  (Guards any cheaper search: a two-stage f32-then-
    f64 search was measured and dropped - same speed - but this is the test it has to pass.)"""
    from md_rdm_b200 import _cabi
    g = torch.Generator().manual_seed(4242)
    q, lv = books[16]
    mask = fr.window_mask(8, 16)                                   # (256, 64) bool: window columns of every row
    inf = torch.tensor(float("inf"), dtype=torch.float64)
    qf = q.float().double()
    cands = torch.cat([q, torch.nextafter(q, inf), torch.nextafter(q, -inf), qf, torch.nextafter(qf, inf), torch.nextafter(qf, -inf),
                       torch.exp(0.5 * torch.randn(64, generator=g, dtype=torch.float64)),
                       torch.tensor([1e-300, 1e300, 0.0, 3.5e38, 3.4028234663852886e38], dtype=torch.float64)])
    B = 4
    x = torch.empty(B, 256, 64, dtype=torch.float64)
    for b in range(B):
        fill = cands[torch.randint(0, cands.numel(), (256,), generator=g)]
        x[b] = fill[:, None].expand(256, 64)
        x[b][mask] = cands[torch.randint(0, cands.numel(), (int(mask.sum()),), generator=g)]
    close = q.clone()
    close[21] = close[20] * (1 + 1e-9)                             # equal after rounding to f32
    close, _ = torch.sort(close)
    assert close.float()[20] == close.float()[21] and close[20] < close[21]
    for thr in (q, close, q[torch.randperm(40, generator=g)]):
        rv, rb = fr.lloyd(x, thr, lv)
        _, _, _, k, bins, vals = R.als_rank1(x.to(dev), _cabi.SRC_RAW_F64, 256, 16, 100, B, thr.to(dev), lv.to(dev), True, True)
        assert torch.equal(bins.view(B, 256, 64).cpu(), rb)
        assert torch.equal(vals.view(B, 256, 64).cpu(), rv.float())


# ============================================================================ repeat-run stress (race hunting; compute-sanitizer is closed on the pool)
def test_stress_pages_argmin_protocol_repeatable(dev, books):
    """The grouped page kernel takes the batch-wide arg-min with a lagged verdict ring between 16 warps (named
    barriers + shared-memory flags).  An ordering bug there shows up as a run-to-run difference, so: 30 reruns of
    the batch-16 path (map and raw sources) must reproduce k*, the record and every map bit for bit, and k* must be
    the oracle's.  Also smooth maps (record plateau: a new minimum nearly every iteration, i.e. the copy path of
    the ring is exercised all the time) and groups of 2 / 5 / 16 / 40 images (single- and multi-round)."""
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(16, scales, seed=4321)
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True)
    w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
    for source in ("map", "raw"):
        plan = FusionPlan(16, scales, source, device=dev)
        srcs = rel if source == "map" else [R.pair_v1(r.to(dev)) if r.shape[2] == 8 else R.pair_id(r.to(dev))[0] for r in rel]
        plan.load_inputs(x_d1.to(dev), [t.to(dev) for t in srcs], w)
        first = None
        for rep in range(30):
            for s in scales:
                plan.rel[s].fill_(-1.0)
                plan.kstar[s].fill_(-1)
            plan.run()
            torch.cuda.synchronize()
            snap = ([plan.kstar[s].clone() for s in scales], [plan.rel[s].clone() for s in scales], [plan.record[s].clone() for s in scales])
            if first is None:
                first = snap
                for si, s in enumerate(scales):
                    for pi, it in enumerate(ref["inter"][si]):
                        assert int(plan.kstar[s].view(-1)[pi]) == it["kstar"], (source, s, pi)
                    assert _rel_err(plan.rel[s].cpu(), ref["rel"][si]) < REL_MAP
            else:
                for a, b in zip(first, snap):
                    for u, v in zip(a, b):
                        assert torch.equal(u, v), (source, rep)
    # smooth maps, several group sizes
    g = torch.Generator().manual_seed(99)
    for group in (2, 5, 16, 40):
        base = torch.exp(0.2 * torch.randn(group, 1, 4, 4, generator=g))
        x = torch.nn.functional.interpolate(base, size=(32, 32), mode="bicubic", align_corners=False).clamp_min(0.05)
        x = (x * torch.exp(0.002 * torch.randn(group, 1, 32, 32, generator=g))).float()
        thr, lvl = _dev_books(books, dev)[32]
        outs = []
        for rep in range(6):
            m, pages, rec, k, _, _ = R.als_rank1(x.to(dev), 4, 256, 32, 100, group, thr, lvl, False, False)
            torch.cuda.synchronize()
            outs.append((m.clone(), rec.clone(), k.clone()))
        for o in outs[1:]:
            assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2]), group
        # against the oracle at our k* (plateau ties are decided by summation order, see test_full_model_config1_golden)
        ks = outs[0][2].view(-1).tolist()
        ref_map = fr.relative_decoder_tail(x, books, force_k=ks)
        assert _rel_err(outs[0][0].cpu(), ref_map) < REL_MAP, group
        ref_free = fr.relative_decoder_tail(x, books, want_intermediates=True)
        for pi, it in enumerate(ref_free[1]):
            rr = np.array(it["record"], dtype=np.float32)
            assert np.allclose(outs[0][1][0, pi].cpu().numpy(), rr, rtol=2e-5, atol=1e-8), (group, pi)
            assert rr[ks[pi]] <= rr.min() * (1 + 3e-6), (group, pi, ks[pi], int(rr.argmin()))


def test_pages_kernel_forms_agree(dev, books):
    """The page ALS exists as one CTA of 16 warps per (batch, page) and as a cluster of 4 CTAs of 4 warps exchanging
    residuals through distributed shared memory (st.async + mbarrier).  The launch size picks one; both flags force
    one.  Same k*, record and maps bit for bit, for one batch per launch and for many, run after run."""
    from md_rdm_b200 import _cabi
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32)

    def same_bits(u, v):   # NaNs included (random weights can make the weighted sum negative, as in the reference)
        as_int = {torch.float32: torch.int32, torch.float64: torch.int64}
        return torch.equal(u.view(as_int.get(u.dtype, u.dtype)), v.view(as_int.get(v.dtype, v.dtype)))

    for G in (1, 3, 20):
        x_d1, rel, weights = fr.synthetic_batch(16 * G, scales, seed=777 + G)
        w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
        snaps = {}
        for name, flags in (("auto", 0), ("one_cta", _cabi.ALS_PAGES_ONE_CTA), ("cluster", _cabi.ALS_PAGES_CLUSTER)):
            plan = FusionPlan(16 * G, scales, "map", group=16, device=dev, flags=flags)
            plan.load_inputs(x_d1.to(dev), [t.to(dev) for t in rel], w)
            for rep in range(5 if name == "cluster" else 1):
                for s in scales:
                    plan.rel[s].fill_(-1.0)
                    plan.kstar[s].fill_(-1)
                out = plan.run()
                torch.cuda.synchronize()
                snap = [plan.kstar[s].clone() for s in scales] + [plan.rel[s].clone() for s in scales] + [plan.record[s].clone() for s in scales] + [out.clone()]
                if name in snaps:
                    for u, v in zip(snaps[name], snap):
                        assert same_bits(u, v), (G, name, rep)
                snaps[name] = snap
        for name in ("one_cta", "cluster"):
            for u, v in zip(snaps["auto"], snaps[name]):
                assert same_bits(u, v), (G, name)


def test_skip_unused_pages_flag_changes_nothing_downstream(dev, books):
    """CP:218-238 as written copies only pages 0 .. side/16 - 1 of an image into the re-tiled map; the reference computes
    the other pages (2 of 4 at 32x32, 12 of 16 at 64x64) and drops them.  The opt-in RDM_ALS_SKIP_UNUSED_PAGES leaves
    them out: the filled maps, y_hat and the final depth must not move by a bit, the live pages keep their k*, record and
    page vectors, for both sources and both forms of the page kernel; and with RDM_ALS_CORRECT_TILING (every page is
    used) the flag is ignored."""
    from md_rdm_b200 import _cabi
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32, 64)
    for G, source in ((1, "map"), (1, "raw"), (3, "raw")):
        N = 16 * G
        x_d1, rel, weights = fr.synthetic_batch(N, scales, seed=5150 + G)
        w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
        srcs = rel if source == "map" else [R.pair_v1(r.to(dev)) if r.shape[2] == 8 else R.pair_id(r.to(dev))[0] for r in rel]
        for extra in (0, _cabi.ALS_PAGES_ONE_CTA, _cabi.ALS_CORRECT_TILING):
            outs = []
            for skip in (0, _cabi.ALS_SKIP_UNUSED_PAGES):
                plan = FusionPlan(N, scales, source, group=16, device=dev, flags=extra | skip)
                plan.load_inputs(x_d1.to(dev), [t.to(dev) for t in srcs], w)
                for s_ in scales:
                    plan.kstar[s_].fill_(-7)
                out = plan.run()
                torch.cuda.synchronize()
                outs.append((out.clone(), plan.yhat.clone(), {s_: plan.rel[s_].clone() for s_ in scales},
                             {s_: plan.kstar[s_].clone() for s_ in scales}, {s_: plan.record[s_].clone() for s_ in scales},
                             {s_: plan.pages[s_].clone() for s_ in scales}))
            full, skipped = outs
            assert _eq_nan(full[0], skipped[0]) and torch.equal(full[1].view(torch.int32), skipped[1].view(torch.int32)), (G, source, extra)
            for s_ in scales:
                assert torch.equal(full[2][s_], skipped[2][s_]), (G, source, extra, s_)
                ratio = max(s_ // 16, 1)
                live = 1 if s_ == 8 else ((s_ // 16) ** 2 if (extra & _cabi.ALS_CORRECT_TILING) else ratio)
                assert torch.equal(full[3][s_][:, :live], skipped[3][s_][:, :live])
                assert torch.equal(full[4][s_][:, :live], skipped[4][s_][:, :live])
                assert torch.equal(full[5][s_][:, :live], skipped[5][s_][:, :live])
                if live < full[3][s_].shape[1]:      # the skipped pages were really skipped
                    assert bool((skipped[3][s_][:, live:] == -7).all())


# ============================================================================ the literal call sequence of the reference
def test_literal_call_sequence_rn_383_396(dev, books):
    """RN:383-396 written against the drop-in NAMES exactly as the reference calls them - cp.resize,
    cp.split_matrix, Ordinal_Layer.sparse_comparison_id, cp.alternating_least_squares, cp.reconstruct (and
    sparse_comparison_v1 + cp.quadratic_als for the 8x8 decoder, RN:359-368; cp.multi_upsample /
    cp.get_resized_area used directly) - against the fused Ordinal_Layer.forward and the oracle."""
    import md_rdm_b200.computations as cp
    from md_rdm_b200.rdm_net import Ordinal_Layer, Quantization
    quant = Quantization()
    g = torch.Generator().manual_seed(383)
    for s in (8, 16, 32, 64):
        x = torch.exp(0.3 * torch.randn(3, 1, s, s, generator=g))
        layer = Ordinal_Layer(int(math.log2(s)) + 3, False, quant)
        xd = x.to(dev)
        if s == 8:                                                      # RN:359-368
            literal = cp.quadratic_als(layer.sparse_comparison_v1(xd), True, n=3, limit=30)
        elif s == 16:                                                   # RN:370-381
            dn_1 = cp.resize(xd, 8)
            literal = cp.alternating_least_squares(layer.sparse_comparison_id(xd, dn_1), 4, True, limit=100)
        else:                                                           # RN:383-396
            dn_1 = cp.resize(xd, s // 2)
            pages, parents = cp.split_matrix(xd, dn_1)
            assert len(pages) == (s // 16) ** 2 and pages[1].shape == (3, 1, 16, 16) and parents[1].shape == (3, 1, 8, 8)
            outs = [cp.alternating_least_squares(layer.sparse_comparison_id(pg, par), 4, True, limit=100) for pg, par in zip(pages, parents)]
            literal = cp.reconstruct(outs)
        fused = layer(xd)
        ref = fr.relative_decoder_tail(x, books)
        assert literal.shape == (3, 1, s, s) and literal.dtype == torch.float32
        assert _rel_err(literal.cpu(), ref) < REL_MAP, s
        assert _rel_err(fused.cpu(), ref) < REL_MAP, s
        assert _rel_err(literal.cpu(), fused.cpu()) < REL_MAP, s
    # cp.multi_upsample (CP:362-366) and cp.get_resized_area (CP:269-295) as free functions
    y = torch.randn(2, 1, 4, 4, generator=g)
    up = cp.multi_upsample(y.to(dev), 3)
    assert up.dtype == torch.float64 and torch.equal(up.cpu(), y.double().repeat_interleave(8, 2).repeat_interleave(8, 3))
    assert cp.multi_upsample(y.to(dev), 0).dtype == torch.float32
    par = torch.rand(2, 1, 8, 8, generator=g, dtype=torch.float64) + 0.5
    area = cp.get_resized_area(2, 4, 3, 6, par.to(dev)).cpu()
    want = torch.ones_like(par)
    want[:, :, 2:5, 3:6] = par[:, :, 2:5, 3:6]
    assert area.shape == (2, 1, 64) and torch.equal(area, want.view(2, 1, 64))


def test_training_step_config3_batch16_fast_path(dev, books):
    """BASELINE config 3 at its full batch (16) through md_rdm_b200.training.TrainingStep (what bench.py --config
    train times): loss terms, final depth and the gradients of Weights and of the DORN logits against torch
    autograd on the CPU oracle."""
    import bench
    from md_rdm_b200.training import TrainingStep
    scales, B = (8, 16, 32), 16
    _, rel, weights = fr.synthetic_batch(B, scales, seed=1603)
    y_raw, logits = bench.synthetic_gt(B, 1604)
    w_flat = torch.cat([w.reshape(-1) for w in weights])
    ts = TrainingStep(B, scales, device=dev)
    ts.load(rel, y_raw, logits, w_flat)
    out = ts.step()
    torch.cuda.synchronize()
    # ---- oracle
    lg = logits.clone().requires_grad_(True)
    decode, ord_ = fr.dorn_regression(lg)
    w_ref = [w.clone().requires_grad_(True) for w in weights]
    fwd = fr.fusion_forward(decode, rel, w_ref, books)
    loss_ref, mse_ref, fine_ref, final_ref = fr.training_loss(y_raw, fwd["y_hat"])
    y128 = fr.mask_target(fr.resize(y_raw, 128))
    ord_ref = fr.ordinal_loss(ord_, fr.depth2label_sid(fr.resize(y128, 8)))
    (loss_ref + ord_ref).backward()
    assert _depth_ok(out["final"].cpu(), final_ref.detach())
    assert abs(out["mse"].item() - mse_ref.item()) <= 1e-6 * abs(mse_ref.item())
    assert abs(out["fine"].item() - fine_ref.item()) <= 1e-5 * abs(fine_ref.item())
    assert abs(out["ord"].item() - ord_ref.item()) <= 1e-5 * abs(ord_ref.item())
    assert abs(out["loss"].item() - (loss_ref + ord_ref).item()) <= 1e-5 * abs((loss_ref + ord_ref).item())
    g_ref = torch.cat([w.grad.reshape(-1) for w in w_ref])
    assert _rel_err(ts.weights.grad.cpu(), g_ref) < 1e-4
    assert torch.allclose(ts.logits.grad.cpu(), lg.grad, rtol=1e-4, atol=1e-9)
    # a second step on the same inputs reproduces the first bit for bit (no state leaks between steps)
    g_first = ts.weights.grad.clone()
    out2 = ts.step()
    assert torch.equal(out2["loss"], out["loss"])
    # the CUDA graph of the whole step (two parallel branches: ground truth | network side; fork/join of the dense
    # ALS) replays to the same bits, twice
    ts.capture()
    for _ in range(2):
        out3 = ts.replay()
        torch.cuda.synchronize()
        assert torch.equal(out3["loss"], out["loss"]) and torch.equal(ts.weights.grad, g_first)
    # the stand-alone ground-truth ops + per-scale torch MSE instead of rdm::gt_prepare + rdm::component_loss
    ts2 = TrainingStep(B, scales, device=dev, fused_gt=False)
    ts2.load(rel, y_raw, logits, w_flat)
    out4 = ts2.step()
    assert abs(out4["loss"].item() - out["loss"].item()) <= 1e-12 * abs(out["loss"].item())
    assert torch.equal(ts2.weights.grad, g_first)


def test_fuse_maps_autograd_and_cache(dev, books):
    """fuse_maps presented as the replacement of RN:103-133 + MOD:132 must carry gradients to
    Weights when autograd records, and its plan cache must not serve another Quantization's codebooks."""
    from md_rdm_b200 import fusion
    from md_rdm_b200.rdm_net import Quantization, Weights
    scales = (8, 16)
    x_d1, rel, weights = fr.synthetic_batch(2, scales, seed=2024)
    wl = Weights(vector_sizes=fr.slot_sizes(scales), use_cuda=True, relative_only=False)
    with torch.no_grad():
        for p_, w_ in zip(wl.weight_list, weights):
            p_.copy_(w_)
    depth, y_hat, filled = fusion.fuse_maps(x_d1.to(dev), [r.to(dev) for r in rel], wl.weight_list)
    assert depth.requires_grad and y_hat[0].requires_grad
    (depth ** 2).mean().backward()
    w_ref = [w.clone().requires_grad_(True) for w in weights]
    o = fr.fusion_forward(x_d1, rel, w_ref, books)
    (fr.recombination(list(o["y_hat"])) ** 2).mean().backward()
    for p_, r_ in zip(wl.weight_list, w_ref):
        if p_.numel():
            assert _rel_err(p_.grad.cpu(), r_.grad) < 1e-4
    with torch.no_grad():
        d2, _, _ = fusion.fuse_maps(x_d1.to(dev), [r.to(dev) for r in rel], wl.weight_list)
    assert not d2.requires_grad and _depth_ok(d2.cpu(), depth.detach().cpu())
    # another codebook object -> another plan (keyed on the object, not on a recyclable id)
    q2 = Quantization()
    thr, lvl = q2.get_with_id(4)
    q2.depth_ratio_016_016_quant = thr * 1.01
    if hasattr(q2, "_dev"):
        q2._dev.clear()
    n0 = len(fusion._plans)
    with torch.no_grad():
        d3, _, f3 = fusion.fuse_maps(x_d1.to(dev), [r.to(dev) for r in rel], wl.weight_list, quant=q2)
    assert len(fusion._plans) == n0 + 1
    fusion.clear_plans()
    assert len(fusion._plans) == 0


def test_stress_tail_bands_and_gm_cluster(dev, books):
    """fuse_tail_kernel exchanges fine-detail logs between the CTAs of a cluster through distributed shared memory
    (1, 2, 4 or 8 bands per image depending on the batch), gm_cluster_kernel its partial products: 20 reruns on
    >= 256 images must be bit-identical, and every band count must give the same values for the same image."""
    g = torch.Generator().manual_seed(808)
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(256, scales, seed=909)
    w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
    xd = x_d1.to(dev)
    filled = [torch.exp(0.2 * torch.randn(256, 1, s, s, generator=g)).to(dev) for s in scales]
    ref_depth = None
    for n in (256, 16, 8, 4, 1):             # bands per image: 1, 1, 2, 4, 8 (the launch wants >= 16 CTAs in total)
        d0, y0, _ = R.fuse_tail(xd[:n], [f[:n] for f in filled], w, False)
        for rep in range(20 if n == 256 else 5):
            d, y, _ = R.fuse_tail(xd[:n], [f[:n] for f in filled], w, False)
            assert _eq_nan(d, d0) and _eq_nan(y, y0), (n, rep)
        if ref_depth is None:
            ref_depth = d0
        else:
            assert _eq_nan(d0, ref_depth[:n]), n
        if n == 16:   # the caller's choice of bands (rdm_fuse_tail_bands): same bits
            for bands in (1, 2, 4, 8):
                db, yb, _ = R.fuse_tail(xd[:n], [f[:n] for f in filled], w, False, bands)
                assert _eq_nan(db, d0) and _eq_nan(yb, y0), bands
    y = (0.5 + 9.5 * torch.rand(256, 1, 128, 128, generator=g, dtype=torch.float64)).to(dev)
    n0 = R.gm_normalize(y)
    p0 = R.decompose(n0, False)
    for rep in range(20):
        assert torch.equal(R.gm_normalize(y), n0), rep
        assert torch.equal(R.decompose(n0, False), p0), rep
    assert torch.equal(R.gm_normalize(y[:3]), n0[:3])


def test_gt_prepare_one_launch(dev):
    """resize 226->128 + mask + gm-normalise + decompose n=7 + ordinal target in ONE launch
    (rdm::gt_prepare) against the oracle's restatement of MOD:68-78, 119-127 (<= 1e-11; the oracle itself is pinned to the
    reference by gt_decompose_b2.npz, whose stored map is already masked and so cannot be fed through the mask again) and,
    bit for bit, against the stand-alone ops it replaces."""
    import bench
    import md_rdm_b200.computations as cp
    from md_rdm_b200.loss import depth2label_sid
    from md_rdm_b200.ops import unpack_pyramid
    for seed, smooth in ((11, True), (12, False)):
        if smooth:
            y_raw, _ = bench.synthetic_gt(5, seed)
        else:
            gen = torch.Generator().manual_seed(seed)
            y_raw = 0.5 + 9.5 * torch.rand(5, 1, 226, 226, generator=gen, dtype=torch.float64)
            y_raw = y_raw * (torch.rand(5, 1, 226, 226, generator=gen) > 0.05)
        for dt in (torch.float64, torch.float32):
            yr = y_raw.to(dt)
            y, pyr, ord_t = R.gt_prepare(yr.to(dev))
            comps = unpack_pyramid(pyr, 5, 128, False)
            # oracle (CPU): MOD:68, 74-78, 119-127
            y_ref = fr.mask_target(fr.resize(yr, 128))
            tg = fr.training_targets(y_ref)
            lab_ref = fr.depth2label_sid(fr.resize(y_ref, 8))
            # (masked pixels sit at 1e-4 + a bicubic residue: the residue's rounding shows at ~1e-12 relative)
            assert torch.allclose(y.cpu(), y_ref, rtol=1e-12, atol=1e-13)
            for i in range(1, 8):
                assert torch.allclose(comps[i].cpu(), tg[i], rtol=1e-11, atol=1e-12), (seed, i)
            # stand-alone ops on the same device (the previous four-launch path): labels and ordinal D_0 identical
            y2 = cp.resize(yr.to(dev), 128)
            y2 = (y2 * (y2 > 0)) + ((y2 <= 0) + 1e-4)
            assert torch.equal(y2, y)
            lab2 = depth2label_sid(cp.resize(y2, 8))
            assert torch.equal(ord_t, lab2)
            d0 = cp.decompose_depth_map([], R.gm_normalize(lab2.long()), 3)[::-1][0]
            assert _eq_nan(comps[0], d0)
            if smooth:      # white-noise ground truth drives the 8x8 bicubic below zero -> NaN labels in the reference
                assert torch.equal(ord_t.cpu(), lab_ref)
                assert _rel_err(comps[0].cpu(), tg[0]) < 1e-6


def test_training_fused_tail_backward_and_component_loss(dev):
    """The training step's two fused pieces against the per-slot ops they replace: rdm::fuse_tail_bwd (pooled
    gradients + every weight's reduction, two launches) is bit-identical to recombination_bwd + make_pred_bwd slot by
    slot - also through autograd, where y_hat's absent gradient selects it - and rdm::component_loss (CP:499-510 in one
    launch) equals the per-scale torch MSE sum to f64 rounding."""
    from md_rdm_b200.ops import fuse_tail_autograd, split_yhat, tail_layout, unpack_pyramid
    for scales in ((8, 16, 32), (8, 16, 32, 64), (16,)):
        B = 5
        x_d1, rel, weights = fr.synthetic_batch(B, scales, seed=31 + len(scales))
        w = torch.cat([t.reshape(-1) for t in weights]).to(dev).requires_grad_(True)
        rel_d = [r.to(dev) for r in rel]
        K, off, kmax, nw = tail_layout(list(scales))
        depth, yhat, A = R.fuse_tail(x_d1.to(dev), rel_d, w.detach(), True)
        g = torch.randn(B, 1, 128, 128, dtype=torch.float64, generator=torch.Generator().manual_seed(5)).to(dev)
        gw = R.fuse_tail_bwd(g, list(A))
        gs = R.recombination_bwd(g, [2 ** k for k in range(kmax + 1)], False, 7)
        ref = torch.zeros(nw, dtype=torch.float32, device=dev)
        for k in range(kmax + 1):
            _, gk = R.make_pred_bwd(A[k], w.detach()[off[k]:off[k] + K[k]], gs[k].reshape(B, -1))
            ref[off[k]:off[k] + K[k]] = gk
        assert torch.equal(gw, ref), scales
        # through autograd: depth only (fused backward) vs depth + 0 * y_hat (slot-by-slot chain)
        final, yh = fuse_tail_autograd(x_d1.to(dev), rel_d, w)
        (g1,) = torch.autograd.grad((final * g).sum(), w)
        final, yh = fuse_tail_autograd(x_d1.to(dev), rel_d, w)
        (g2,) = torch.autograd.grad((final * g).sum() + 0.0 * yh.sum(), w)
        assert torch.equal(g1, g2), scales
        # component loss against the per-scale MSE sum
        y_raw = 0.5 + 9.5 * torch.rand(B, 1, 226, 226, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
        _, pyr, _ = R.gt_prepare(y_raw.to(dev))
        comps = unpack_pyramid(pyr, B, 128, False)
        want = torch.stack([torch.nn.functional.mse_loss(a.double(), b) for a, b in zip(split_yhat(yhat, kmax), comps)]).sum()
        got = R.component_loss(yhat, pyr, kmax)
        assert got.dtype == torch.float64 and abs(float(got) - float(want)) <= 1e-12 * abs(float(want)), (float(got), float(want))


# ============================================================================ SURVEY 8f rank 4: "paper-correct" flags
@pytest.mark.parametrize("flag_name", ["true_gm", "correct_tiling", "true_transpose", "all"])
def test_paper_correct_flags_vs_oracle(dev, books, flag_name):
    """The three knobs the reference's authors evidently intended (true transpose in the ALS q-update, CP:64/133; the
    geometric mean instead of quick_gm's squared rc, CP:244-255; page (i,j) to block (i,j) instead of CP:218-238's
    tiling) are OFF by default and mirrored in the oracle: each flag, and all together, against the oracle with the
    same flag; and the default path is bit-identical to a plan built without the argument."""
    from md_rdm_b200 import _cabi
    from md_rdm_b200.fusion import FusionPlan
    flags = {"true_gm": _cabi.ALS_TRUE_GM, "correct_tiling": _cabi.ALS_CORRECT_TILING, "true_transpose": _cabi.ALS_TRUE_TRANSPOSE,
             "all": _cabi.ALS_TRUE_GM | _cabi.ALS_CORRECT_TILING | _cabi.ALS_TRUE_TRANSPOSE}[flag_name]
    assert (fr.FLAG_TRUE_TRANSPOSE, fr.FLAG_TRUE_GM, fr.FLAG_CORRECT_TILING) == (_cabi.ALS_TRUE_TRANSPOSE, _cabi.ALS_TRUE_GM, _cabi.ALS_CORRECT_TILING)
    scales = (8, 16, 32, 64)
    x_d1, rel, weights = fr.synthetic_batch(3, scales, seed=64)
    w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True, flags=flags)
    ref0 = fr.fusion_forward(x_d1, rel, weights, books)
    changed = any(not torch.equal(a, b) for a, b in zip(ref["rel"], ref0["rel"]))
    assert changed, "the flag must change the oracle's result, or the test proves nothing"
    for source in ("map", "raw"):
        plan = FusionPlan(3, scales, source, device=dev, flags=flags)
        srcs = rel if source == "map" else [R.pair_v1(r.to(dev)) if r.shape[2] == 8 else R.pair_id(r.to(dev))[0] for r in rel]
        plan.load_inputs(x_d1.to(dev), [t.to(dev) for t in srcs], w)
        plan.run()
        torch.cuda.synchronize()
        # with the true transpose the iteration converges and the record flattens out: our k* may be another tie of the
        # oracle's record (as on real decoder outputs, test_full_model_config1_golden); maps are compared at a common k*
        ks = [plan.kstar[s].view(-1).tolist() for s in scales]
        for si, s in enumerate(scales):
            for pi, it in enumerate(ref["inter"][si]):
                rr = np.array(it["record"], dtype=np.float32)
                assert rr[ks[si][pi]] <= rr.min() * (1 + 3e-6), (source, s, pi, ks[si][pi], it["kstar"])
                if not flags & _cabi.ALS_TRUE_TRANSPOSE:
                    assert ks[si][pi] == it["kstar"], (source, s, pi)
        forced = fr.fusion_forward(x_d1, rel, weights, books, force_k=ks, flags=flags)
        for si, s in enumerate(scales):
            assert _rel_err(plan.rel[s].cpu(), forced["rel"][si]) < 2e-5, (source, s)
        assert _depth_ok(plan.depth.cpu(), forced["depth"])
    # default: flags=0 is today's path, bit for bit
    a = FusionPlan(3, scales, "map", device=dev)
    b = FusionPlan(3, scales, "map", device=dev, flags=0)
    for p_ in (a, b):
        p_.load_inputs(x_d1.to(dev), [t.to(dev) for t in rel], w)
        p_.run()
    torch.cuda.synchronize()
    assert _eq_nan(a.depth, b.depth) and all(torch.equal(a.rel[s], b.rel[s]) for s in scales)
    assert _depth_ok(a.depth.cpu(), ref0["depth"])


def test_conv_head_fused_with_pair_build(dev, books):
    """the decoders' 1x1 conv heads (RN:146, RN:157) fused with the pair build.  The conv map agrees
    with torch's own f32 conv to rounding; everything downstream is EXACT with respect to the map the kernel produced
    (bins bit-equal to the oracle fed with that map, k* equal, depth within tolerance); for s >= 16 the map is only
    written to HBM on request."""
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32, 64)
    chans = {8: 2208, 16: 1664, 32: 832, 64: 416}          # _wsm_output_planes(6..9), RN:555-565
    B = 3
    g = torch.Generator().manual_seed(146)
    x_d1, _, weights = fr.synthetic_batch(B, scales, seed=147)
    feats, cw, cb, ref_maps = {}, {}, {}, {}
    for s in scales:
        C = chans[s]
        f = torch.randn(B, C, s, s, generator=g) * 0.5
        w = torch.randn(1, C, 1, 1, generator=g) / math.sqrt(C) * 0.6
        b = torch.tensor([1.5])                                # keeps the map positive like a trained decoder's
        ref_maps[s] = torch.nn.functional.conv2d(f.double(), w.double(), b.double())     # exact reference for the map
        feats[s], cw[s], cb[s] = f.to(dev), w.to(dev), b.to(dev)
    plan = FusionPlan(B, scales, "map", device=dev)
    plan.load_inputs(x_d1.to(dev), [torch.ones(B, 1, s, s) for s in scales], torch.cat([t.reshape(-1) for t in weights]).to(dev))
    for s in scales:
        plan.src[s].fill_(-5.0)
    plan.run_from_features(feats, cw, cb)
    torch.cuda.synchronize()
    assert float(plan.src[16].max()) == -5.0 and float(plan.src[64].max()) == -5.0     # maps of s >= 16 never written
    first = {s: plan.rel[s].clone() for s in scales}
    depth1 = plan.depth.clone()
    plan.run_from_features(feats, cw, cb, write_maps=True)
    torch.cuda.synchronize()
    assert _eq_nan(plan.depth, depth1) and all(torch.equal(plan.rel[s], first[s]) for s in scales)
    maps = [plan.src[s].cpu().clone() for s in scales]
    for s, m in zip(scales, maps):
        assert m.min() > 0
        assert torch.allclose(m.double(), ref_maps[s], rtol=2e-6, atol=2e-6), s          # f32 summation of up to 2208 terms
        f32 = torch.nn.functional.conv2d(feats[s].cpu(), cw[s].cpu(), cb[s].cpu())
        assert (m - f32).abs().max() <= 4e-6                                               # torch's own f32 conv: same ball park
    ref = fr.fusion_forward(x_d1, maps, weights, books, want_intermediates=True)
    for si, s in enumerate(scales):
        for pi, it in enumerate(ref["inter"][si]):
            assert torch.equal(plan.bins[s][:, pi].cpu(), it["bins"]), (s, pi)
            assert int(plan.kstar[s].view(-1)[pi]) == it["kstar"], (s, pi)
        assert _rel_err(plan.rel[s].cpu(), ref["rel"][si]) < REL_MAP
    assert _depth_ok(plan.depth.cpu(), ref["depth"])


def test_compact_result_is_the_map(dev, books):
    """Opt-in compact result (e2e scaling): no decoder is finer than 2^kmax, so the log-depth map is
    constant on 2^(7-kmax) blocks; depth_compact holds its distinct values and expands to the full map bit for bit, on
    the device path and through the pinned host call."""
    from md_rdm_b200.fusion import FusionPlan
    for scales in ((8, 16, 32), (8, 16, 32, 64), ()):
        x_d1, rel, weights = fr.synthetic_batch(4, scales, seed=7 + len(scales))
        w = torch.cat([t.reshape(-1) for t in weights]).to(dev)
        plan = FusionPlan(4, scales, "map", device=dev, compact_result=True)
        plan.load_inputs(x_d1.to(dev), [r.to(dev) for r in rel], w)
        plan.run()
        torch.cuda.synchronize()
        assert plan.depth_compact.shape == (4, 1, 1 << plan.kmax, 1 << plan.kmax)
        assert _eq_nan(plan.expand_compact(plan.depth_compact), plan.depth)
        host = plan.run_host(x_d1, rel)
        assert host.shape == plan.depth_compact.shape and _eq_nan(plan.expand_compact(host), plan.depth.cpu())
        assert plan.d2h_bytes() == 4 * 8 * 4 ** plan.kmax


def test_dorn_clamp_is_f32_and_nan_propagates(dev):
    """RN:334 clamps the f32 logits (f32 bounds) before widening, and torch.clamp / softmax propagate NaN."""
    from md_rdm_b200.rdm_net import Ordinal_Layer
    x = torch.zeros(1, 6, 8, 8)
    x[0, 0, 0, 0], x[0, 1, 0, 0] = 1e-9, 5e-9          # both below the lower bound: clamp to float32(1e-8) exactly
    x[0, 2, 0, 1] = float("nan")
    x[0, 4, 0, 2], x[0, 5, 0, 2] = 2e4, -3.0            # above / below the bounds
    xd = x.to(dev).requires_grad_(True)
    decode, ord_ = Ordinal_Layer(1, True, None)(xd)
    dref, oref = fr.dorn_regression(x)
    assert torch.equal(torch.isnan(ord_.cpu()), torch.isnan(oref))
    assert torch.allclose(torch.nan_to_num(ord_.detach().cpu(), nan=0.0), torch.nan_to_num(oref, nan=0.0), rtol=1e-14, atol=0)
    assert torch.equal(decode.cpu(), dref)
    ord_.nan_to_num(nan=0.0).sum().backward()
    assert float(xd.grad[0, 4, 0, 2]) == 0.0 and float(xd.grad[0, 5, 0, 2]) == 0.0      # clamped: no gradient


def test_depth2label_sid_kernel(dev):
    """utils.py:195-211 as one kernel: f64 maps bit-equal to the reference formula evaluated by torch
    on the CPU (the oracle), f32 maps equal except where the f32 log lands within an ulp of an integer."""
    from md_rdm_b200.loss import depth2label_sid
    g = torch.Generator().manual_seed(195)
    d64 = 0.02 + 12.0 * torch.rand(7, 1, 8, 8, generator=g, dtype=torch.float64)
    d64[0, 0, 0, 0] = 0.01                                # below alpha: negative label -> 0
    assert torch.equal(depth2label_sid(d64.to(dev)).cpu(), fr.depth2label_sid(d64))
    d32 = d64.float()
    ours, ref = depth2label_sid(d32.to(dev)).cpu(), fr.depth2label_sid(d32)
    assert ours.dtype == torch.int32 and (ours - ref).abs().max() <= 1 and (ours != ref).float().mean() < 0.01
    gold = load_golden("dorn_loss.npz")
    gen = torch.Generator().manual_seed(808)
    torch.randn(3, 180, 8, 8, generator=gen)
    depth = 0.5 + 9.5 * torch.rand(3, 1, 8, 8, generator=gen, dtype=torch.float64)
    assert torch.equal(depth2label_sid(depth.to(dev)).cpu(), torch.from_numpy(gold["target"]))


def test_fusion_plan_with_decoder_10(dev, books):
    """VERDICT r1 missing #5: the 128x128 relative decoder (decoder 10, RN:61, 64 pages per image) inside FusionPlan and
    fuse_maps: its pair build / Lloyd / ALS share the launches of the other scales, the tail is composed from the
    stand-alone kernels.  All five relative decoders against the oracle."""
    from md_rdm_b200.fusion import FusionPlan
    scales = (8, 16, 32, 64, 128)
    x_d1, rel, weights = fr.synthetic_batch(2, scales, seed=128)
    ref = fr.fusion_forward(x_d1, rel, weights, books, want_intermediates=True)
    for source in ("map", "raw"):
        plan = _run_plan(dev, x_d1, rel, weights, source, want_A=True)
        assert plan.kmax == 7 and plan.composed_tail
        for si, s in enumerate(scales):
            for pi, it in enumerate(ref["inter"][si]):
                assert torch.equal(plan.bins[s][:, pi].cpu(), it["bins"]), (source, s, pi)
                assert int(plan.kstar[s].view(-1)[pi]) == it["kstar"], (source, s, pi)
            assert _rel_err(plan.rel[s].cpu(), ref["rel"][si]) < REL_MAP, (source, s)
        for y, yr in zip(plan.yhat_list(), ref["y_hat"]):
            assert torch.allclose(y.cpu(), yr, rtol=0, atol=5e-5)
        assert _depth_ok(plan.depth.cpu(), ref["depth"]), source
        eager = plan.depth.clone()
        plan.depth.zero_()
        plan.replay()                                    # the composed tail is CUDA-graph capturable too
        torch.cuda.synchronize()
        assert _eq_nan(plan.depth, eager)
