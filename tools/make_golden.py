#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (az16/MD_RDM)
on seeded synthetic inputs, and pin oracle/fusion_ref.py against it.

Build-container only: needs /root/reference (read-only).  The reference loads its
codebooks by cwd-relative path (network/RDM_Net.py:403-407) and the 008 table is
missing from its tree, so we run it from a temp dir holding the four shipped .mat
files plus the derived 008 stand-in (tools/import_codebooks.py).

    python tools/make_golden.py            # writes tests/golden/, prints oracle-vs-reference diffs
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np
import scipy.io
import torch

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import fusion_ref as fr  # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    books = fr.load_codebooks()
    tmp = tempfile.mkdtemp(prefix="rdm_ref_cwd_")
    for s in ("016", "032", "064", "128"):
        shutil.copy(os.path.join(REF, f"depth_ratio_{s}_{s}_quant.mat"), tmp)
    q8, l8 = books[8]
    scipy.io.savemat(os.path.join(tmp, "depth_ratio_008_008_quant.mat"),
                     {"depth_ratio_008_008_quant": q8.numpy().reshape(40, 1),
                      "depth_ratio_008_008_quant_inv": l8.numpy().reshape(41, 1)})
    os.chdir(tmp)
    import network.RDM_Net as rn
    import network.computations as cp
    rn.use_cuda = False
    quant = rn.Quantization()
    # the shipped tables must equal what tools/import_codebooks.py stored
    for s in (16, 32, 64, 128):
        qq, ll = quant.get_with_id(int(np.log2(s)))
        assert np.array_equal(qq[:, 0], books[s][0].numpy()) and np.array_equal(ll[:, 0], books[s][1].numpy())

    gold = os.path.join(REPO, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    report = {}

    # ---------------------------------------------------------------- A. Lloyd edge cases (RN:286-311)
    layer = {i: rn.Ordinal_Layer(i + 3, False, quant) for i in (3, 4, 5, 6, 7)}
    out = {}
    for sid in (3, 4, 5):
        s = 1 << sid
        q, lv = books[s]
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            qd = q.to(dt)
            vals = [qd, torch.nextafter(qd, torch.tensor(0.0, dtype=dt)), torch.nextafter(qd, torch.tensor(9.0, dtype=dt)),
                    q.to(dt) * 0.999, q.to(dt) * 1.001,
                    torch.tensor([0.0, -1.0, float("nan"), float("inf"), 1e-30, 1e30, 1.0, 0.05, 50.0], dtype=dt)]
            g = torch.Generator().manual_seed(100 + sid)
            vals.append(torch.exp(0.6 * torch.randn(256 - sum(v.numel() for v in vals) % 256 + 256, generator=g)).to(dt))
            x = torch.cat(vals)
            n = x.numel() // 64 * 64
            x = x[:n].view(1, n // 64, 64).clone()
            ref = layer[sid].LloydQuantization(torch.zeros(1, n // 64, 64, 40), x.clone(), id=sid)
            v, b = fr.lloyd(x, q, lv)
            same = torch.equal(torch.nan_to_num(ref, nan=-7.0), torch.nan_to_num(v, nan=-7.0))
            report[f"lloyd_{s}_{name}_values_bitequal"] = same
            assert same and ref.dtype == dt
            out[f"x_{s}_{name}"] = x.numpy()
            out[f"values_{s}_{name}"] = ref.numpy()
            out[f"bins_{s}_{name}"] = b.numpy()
    np.savez_compressed(os.path.join(gold, "lloyd_edges.npz"), **out)

    # ---------------------------------------------------------------- B. relative decoder tails (RN:358-396)
    out = {}
    B = 2
    g = torch.Generator().manual_seed(2024)
    for sid in (3, 4, 5):
        s = 1 << sid
        x = torch.exp(0.3 * torch.randn(B, 1, s, s, generator=g))
        ref_map = layer[sid](x.clone())
        mine, inter = fr.relative_decoder_tail(x, books, want_intermediates=True)
        # intermediates from the reference's own pair builders (Lloyd applied inside)
        if sid == 3:
            ref_q = [layer[sid].sparse_comparison_v1(x.clone())]
        else:
            dn_1 = cp.resize(x, s // 2)
            assert torch.equal(dn_1, fr.resize_half(x)), "resize_half not bit-equal"
            if sid == 4:
                ref_q = [layer[sid].sparse_comparison_id(x, dn_1)]
            else:
                a, b_ = cp.split_matrix(x, dn_1)
                ref_q = [layer[sid].sparse_comparison_id(p0, p1) for p0, p1 in zip(a, b_)]
        for pi, (rq, it) in enumerate(zip(ref_q, inter)):
            q, lv = books[s]
            v, bn = fr.lloyd(it["raw"], q, lv)
            assert torch.equal(rq, v), f"quantized pair matrix differs s={s} page={pi}"
            out[f"bins_{s}_p{pi}"] = bn.numpy()
            out[f"raw_sha_{s}_p{pi}"] = np.array(sha(it["raw"]))
            out[f"kstar_{s}_p{pi}"] = np.array(it["kstar"])
            out[f"record_{s}_p{pi}"] = np.array(it["record"], dtype=np.float32)
            out[f"page_{s}_p{pi}"] = it["page"].numpy()
        d = (ref_map - mine).abs().max().item()
        report[f"tail_{s}_maxdiff"] = d
        assert d == 0.0, d
        out[f"x_{s}"] = x.numpy()
        out[f"map_{s}"] = ref_map.numpy()
    np.savez_compressed(os.path.join(gold, "relative_tails_b2.npz"), **out)

    # ---------------------------------------------------------------- C. ALS alone (CP:38-155), incl. constant map (k*=0)
    out = {}
    g = torch.Generator().manual_seed(77)
    lv16 = books[16][1]
    for name, (H, W, n, lim) in {"page": (256, 64, 4, 100), "sq": (64, 64, 3, 30)}.items():
        idx = torch.randint(12, 29, (3, H, W), generator=g)
        Rq = lv16[idx].float() * torch.exp(0.2 * torch.randn(3, H, 1, generator=g)).float()
        if name == "page":
            ref = cp.alternating_least_squares(Rq.clone(), n, False, limit=lim)
        else:
            ref = cp.quadratic_als(Rq.clone(), False, n=n, limit=lim)
        mine, rec, k = fr.als_rank1(Rq, lim)
        d = (ref - mine).abs().max().item()
        report[f"als_{name}_maxdiff"] = d
        assert d == 0.0
        out[f"Rq_{name}"] = Rq.numpy()
        out[f"map_{name}"] = ref.numpy()
        out[f"kstar_{name}"] = np.array(k)
        out[f"record_{name}"] = np.array(rec, dtype=np.float32)
        ones = torch.ones(2, H, W)
        refc = cp.alternating_least_squares(ones.clone(), n, False, limit=lim) if name == "page" \
            else cp.quadratic_als(ones.clone(), False, n=n, limit=lim)
        minec, recc, kc = fr.als_rank1(ones, lim)
        assert torch.equal(refc, minec)
        out[f"const_map_{name}"] = refc.numpy()
        out[f"const_kstar_{name}"] = np.array(kc)
    np.savez_compressed(os.path.join(gold, "als.npz"), **out)

    # ---------------------------------------------------------------- D/E. decomposition, weights, recombination, full path
    out = {}
    scales = (8, 16, 32)
    x_d1, rel, weights = fr.synthetic_batch(2, scales, seed=1234)
    o = fr.fusion_forward(x_d1, rel, weights, books)
    # reference, composed exactly as RN:103-133 + MOD:132 with decoders 1,6,7,8
    ref_rel = [layer[int(np.log2(s))](x.clone()) for s, x in zip(scales, rel)]
    Bn, _, H, W = x_d1.size()
    f_d1 = cp.decompose_depth_map([], torch.div(x_d1, cp.quick_gm(x_d1.view(Bn, H * W, 1), H).expand(Bn, H * W).view(Bn, 1, H, W)), 3)[::-1]
    rows = [f_d1] + [cp.decompose_depth_map([], r, int(np.log2(r.shape[2])), relative_map=True)[::-1] for r in ref_rel]
    y = cp.relative_fine_detail_matrix(rows, False)
    K = fr.slot_sizes(scales)
    wl = rn.Weights(vector_sizes=K, use_cuda=False, relative_only=False)
    with torch.no_grad():
        for p_, w_ in zip(wl.weight_list, weights):
            p_.copy_(w_)
    A_ref = [a.clone() for a in y]
    y_hat = wl(y)
    y_hat_keep = [t.detach().clone() for t in y_hat]
    depth = cp.recombination(y_hat).detach()
    for i, (a, b_) in enumerate(zip(A_ref, o["A"])):
        assert torch.equal(a, b_), f"fine-detail matrix {i}"
    for i, (a, b_) in enumerate(zip(y_hat_keep, o["y_hat"])):
        assert torch.equal(a, b_), f"y_hat {i}"
    d = (depth - o["depth"]).abs().max().item()
    report["full_depth_maxdiff"] = d
    assert d == 0.0
    out["x_d1"] = x_d1.numpy()
    for s, x, r in zip(scales, rel, ref_rel):
        out[f"rel_in_{s}"] = x.numpy()
        out[f"rel_out_{s}"] = r.numpy()
    for i, w_ in enumerate(weights):
        out[f"w_{i}"] = w_.numpy()
    for i, a in enumerate(A_ref):
        out[f"A_{i}"] = a.numpy()
    for i, a in enumerate(y_hat_keep):
        out[f"yhat_{i}"] = a.numpy()
    out["depth"] = depth.numpy()
    # gradient of the Weights parameters for loss = mean(depth**2)
    y2 = cp.relative_fine_detail_matrix(rows, False)
    loss = (cp.recombination(wl(y2)) ** 2).mean()
    loss.backward()
    for i, p_ in enumerate(wl.weight_list):
        if p_.numel():
            out[f"grad_w_{i}"] = p_.grad.numpy()
    out["loss"] = np.array(loss.item())
    np.savez_compressed(os.path.join(gold, "full_path_b2.npz"), **out)

    # ---------------------------------------------------------------- F. GT decomposition (MOD:74-78,119-123,145-149)
    out = {}
    g = torch.Generator().manual_seed(5)
    y = 0.5 + 9.5 * torch.rand(2, 1, 128, 128, generator=g, dtype=torch.float64)
    y = y * (torch.rand(2, 1, 128, 128, generator=g) > 0.05)
    yt = fr.mask_target(y)
    norm = torch.div(yt, cp.quick_gm(yt.view(2, 128 * 128, 1), 128).expand(2, 128 * 128).view(2, 1, 128, 128))
    ref = cp.decompose_depth_map([], norm, 7)[::-1]
    mine = fr.gt_components(yt)
    for i, (a, b_) in enumerate(zip(ref, mine)):
        assert torch.equal(a, b_), f"GT component {i}"
        out[f"comp_{i}"] = a.numpy()
    out["y_masked"] = yt.numpy()
    # round trip: recombination(log components) == log(normalised target)
    rt = cp.recombination([torch.log(c) for c in ref])
    report["gt_roundtrip_maxdiff"] = (rt - torch.log(norm)).abs().max().item()
    np.savez_compressed(os.path.join(gold, "gt_decompose_b2.npz"), **out)

    # ---------------------------------------------------------------- resize_half on every size used
    for n in (2, 4, 8, 16, 32, 64, 128):
        xx = torch.rand(2, 1, n, n, generator=g, dtype=torch.float64) + 0.5
        assert torch.equal(cp.resize(xx, n // 2), fr.resize_half(xx)), n
        rel = ((cp.resize(xx, n // 2) - fr.resize_half_explicit(xx)).abs() / cp.resize(xx, n // 2).abs()).max().item()
        assert rel < 1e-15, rel
        xf = xx.float()
        assert torch.equal(cp.resize(xf, n // 2), fr.resize_half_explicit(xf)), n
    report["resize_half_bitequal_2..128"] = True

    with open(os.path.join(gold, "README.md"), "w") as f:
        f.write("# Golden vectors\n\nGenerated by `tools/make_golden.py` from the unmodified reference "
                f"(az16/MD_RDM) on CPU, torch {torch.__version__}, numpy {np.__version__}.\n"
                "The 008 codebook is the derived stand-in (016 table squared).\n\n"
                "Oracle-vs-reference report at generation time:\n\n")
        for k, v in report.items():
            f.write(f"* `{k}`: {v}\n")
    for k, v in report.items():
        print(k, v)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
