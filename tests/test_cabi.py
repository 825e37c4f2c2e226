"""CPU: the C-ABI library builds, loads, exports every symbol include/rdm_b200.h declares, and
rejects bad arguments with the documented error convention (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from md_rdm_b200 import _cabi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rdm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/rdm_b200.h but not exported"
    assert set(syms) == set(_cabi.PROTOTYPES), "ctypes prototypes and header disagree"


def test_abi_version_and_sizes(lib):
    assert lib.rdm_abi_version() == _cabi.ABI_VERSION
    assert lib.rdm_pyramid_len(128, 0) == 21845
    assert lib.rdm_pyramid_len(8, 1) == 84
    assert lib.rdm_pyramid_len(12, 0) == -1
    assert lib.rdm_als_ws_floats(256, 4, 100) == 4 * (256 * 16 + 8)   # compact page form + band flags (no iterate history)
    assert lib.rdm_als_ws_floats(256, 4, 100) <= 4 * 4300
    assert lib.rdm_als_ws_floats(64, 1, 30) == 32                     # the unit's SSE record
    sides = (ctypes.c_int32 * 3)(8, 16, 32)
    assert lib.rdm_fuse_tail_weight_count(sides, 3) == 4 + 3 + 4 + 5
    assert ctypes.sizeof(_cabi.AlsScale) == 8 + 6 * 4 + 9 * 8


def test_binding_constants_match_the_header():
    """The enum values the ctypes host uses are the header's (flags, phase masks, source kinds)."""
    text = open(os.path.join(ROOT, "include", "rdm_b200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(RDM_[A-Z0-9_]+)\s*=\s*(\d+)", text)}
    for name in ("DENSE_ONLY", "TRUE_TRANSPOSE", "TRUE_GM", "CORRECT_TILING", "PAGES_ONE_CTA", "PAGES_CLUSTER", "SKIP_UNUSED_PAGES"):
        assert enums["RDM_ALS_" + name] == getattr(_cabi, "ALS_" + name), name
    every = 0
    for k, v in enums.items():
        if k.startswith("RDM_ALS_") and not k.startswith("RDM_ALS_PHASE_") and k != "RDM_ALS_FLAGS_ALL":
            every |= v
    assert enums["RDM_ALS_FLAGS_ALL"] == every
    for name in ("SPARSIFY", "PAGES", "DENSE", "ALL"):
        assert enums["RDM_ALS_PHASE_" + name] == getattr(_cabi, "PHASE_" + name), name
    for name in ("RAW_F64", "RAW_F32", "VAL_F32", "VAL_F64", "MAP_F32"):
        assert enums["RDM_SRC_" + name] == getattr(_cabi, "SRC_" + name), name


def test_argument_errors_do_not_launch(lib):
    null = ctypes.c_void_p(0)
    rc = lib.rdm_pair_v1_f32(null, 4, null, null)
    assert rc < 0 and b"null pointer" in lib.rdm_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.rdm_pair_id_f64(one, 1, 24, one, null, null)
    assert rc < 0 and b"side must be" in lib.rdm_last_error()
    rc = lib.rdm_decompose(one, 0, 1, 100, 0, one, null)
    assert rc < 0 and b"power of two" in lib.rdm_last_error()
    sc = _cabi.AlsScale(src=16, src_kind=_cabi.SRC_VAL_F32, rows=128, pages=1, side=16, limit=10, ws=16)
    rc = lib.rdm_als_fused(sc, 1, 4, 4, null)
    assert rc < 0 and b"rows must be 64 or 256" in lib.rdm_last_error()
    sc.rows = 256
    rc = lib.rdm_als_fused(sc, 1, 6, 4, null)
    assert rc < 0 and b"multiple of group" in lib.rdm_last_error()
    sc.ws = 20
    rc = lib.rdm_als_fused(sc, 1, 4, 4, null)
    assert rc < 0 and b"ws must be 16-byte aligned" in lib.rdm_last_error()
    sc.ws = 16
    rc = lib.rdm_als_fused_phases(sc, 1, 4, 4, 8, null)
    assert rc < 0 and b"phase_mask" in lib.rdm_last_error()
    sc.flags = 128
    rc = lib.rdm_als_fused(sc, 1, 4, 4, null)
    assert rc < 0 and b"unknown flag bits" in lib.rdm_last_error()
    sc.flags = 0
    # empty batches are a no-op, not an error
    assert lib.rdm_pair_v1_f32(one, 0, one, null) == 0
    assert lib.rdm_als_fused(sc, 1, 0, 4, null) == 0


def test_ops_refuse_cpu_tensors():
    import torch
    import md_rdm_b200.computations as cp
    with pytest.raises(RuntimeError, match="no CPU path"):
        cp.quick_gm(torch.ones(2, 4, 1), 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cp.decompose_depth_map([], torch.ones(1, 1, 8, 8), 3)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.RdmError, match="no CPU or PyTorch fallback"):
        _cabi.load()


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "md_rdm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_quantization_surface_and_mat_loader(tmp_path):
    """Drop-in `Quantization` (RN:397-442): attribute names, get_with_id / get_size_id, and the .mat loader."""
    import numpy as np
    import scipy.io
    from md_rdm_b200.codebooks import Quantization
    q = Quantization()
    assert q.depth_ratio_016_016_quant.shape == (40, 1) and q.depth_ratio_016_016_quant_inv.shape == (41, 1)
    assert q.get_size_id(5) == 32 and q.get_with_id(5)[0] is q.depth_ratio_032_032_quant
    assert q.derived[8] and not q.derived[16]
    assert np.allclose(q.depth_ratio_008_008_quant, q.depth_ratio_016_016_quant ** 2, rtol=1e-15)
    scipy.io.savemat(str(tmp_path / "depth_ratio_008_008_quant.mat"),
                     {"depth_ratio_008_008_quant": np.linspace(0.5, 2.0, 40).reshape(40, 1),
                      "depth_ratio_008_008_quant_inv": np.linspace(0.45, 2.1, 41).reshape(41, 1)})
    q2 = Quantization.from_mat_dir(str(tmp_path))
    assert not q2.derived[8] and q2.depth_ratio_008_008_quant[0, 0] == 0.5
    assert np.array_equal(q2.depth_ratio_016_016_quant, q.depth_ratio_016_016_quant)


def test_codebook_tooling(tmp_path):
    """.mat wire format round trip (RN:403-418) and the generator for missing scales."""
    import numpy as np
    from md_rdm_b200.codebooks import Quantization
    q = Quantization()
    shipped16 = q.depth_ratio_016_016_quant.copy(), q.depth_ratio_016_016_quant_inv.copy()
    shipped64 = q.depth_ratio_064_064_quant.copy(), q.depth_ratio_064_064_quant_inv.copy()
    t, lv = q.derive(16, 32)                              # codebook(s) == codebook(2s)**2
    assert t.shape == (40, 1) and lv.shape == (41, 1) and q.derived[16]
    assert np.allclose(t, shipped16[0], rtol=1e-13, atol=0) and np.allclose(lv, shipped16[1], rtol=1e-13, atol=0)
    t, lv = q.derive(64, 32)                              # ... and the square root going up
    assert np.allclose(t, shipped64[0], rtol=1e-13, atol=0) and np.allclose(lv, shipped64[1], rtol=1e-13, atol=0)
    assert Quantization().derived[8] and not Quantization().derived[16]   # only the missing 008 file is derived in the package
    fresh = Quantization()
    fresh.to_mat_dir(str(tmp_path))
    back = Quantization.from_mat_dir(str(tmp_path))
    for s in (8, 16, 32, 64, 128):
        tag = f"depth_ratio_{s:03d}_{s:03d}_quant"
        assert np.array_equal(getattr(back, tag), getattr(fresh, tag)) and np.array_equal(getattr(back, tag + "_inv"), getattr(fresh, tag + "_inv"))
        assert getattr(back, tag).shape == (40, 1) and not back.derived[s]
    assert back.get_size_id(5) == 32 and back.get_with_id(5)[0] is getattr(back, "depth_ratio_032_032_quant")


def test_sparsify_geometry_tables_match_the_window_mask(lib):
    """The compact page form's geometry (which matrix columns are window columns of which row, and where they
    land in the 3x4 span) against the oracle's window mask (RN:266-273 + CP:269-295), on the CPU."""
    import numpy as np
    from oracle import fusion_ref as fr
    window_cols = np.zeros(256 * 9, dtype=np.int32)
    fill_col = np.zeros(256, dtype=np.int32)
    compact = np.zeros(256 * 16, dtype=np.uint8)
    assert lib.rdm_sparsify_geometry(window_cols.ctypes.data, fill_col.ctypes.data, compact.ctypes.data) == 0
    window_cols, compact = window_cols.reshape(256, 9), compact.reshape(256, 16)
    mask = fr.window_mask(8, 16).numpy()                       # (256 rows, 64 columns)
    for row in range(256):
        r, c = row >> 4, row & 15
        r0, c0, span0 = min(r >> 1, 5), min(c >> 1, 5), min(2 * (c >> 2), 4)
        assert sorted(window_cols[row].tolist()) == np.flatnonzero(mask[row]).tolist()
        assert window_cols[row].tolist() == [8 * (r0 + a) + c0 + b for a in range(3) for b in range(3)]
        # the fill reference column is outside the window of EVERY pixel of this pixel row (they share r0)
        assert not mask[16 * r:16 * r + 16, fill_col[row]].any()
        # compact row: entry 0 = fill, entries 1..12 = span (alpha, gamma) -> window slot or zero
        assert compact[row, 0] == 9 and all(compact[row, 13:] == 15)
        for e in range(12):
            col = 8 * (r0 + e // 4) + span0 + e % 4
            code = int(compact[row, 1 + e])
            assert (code != 15) == bool(mask[row, col]), (row, e)
            if code != 15:
                assert window_cols[row, code] == col
