"""ctypes binding of librdm_b200.so (include/rdm_b200.h).

This is the only place the shared library is touched.  There is NO fallback: if the
library is missing or a symbol is absent, importing/using the ops raises.  Nothing
here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# RDM_B200_LIB selects an experimental build of the same ABI (tools/als_variants.py); default = the product
LIB_PATH = os.environ.get("RDM_B200_LIB") or os.path.join(_HERE, "librdm_b200.so")
ABI_VERSION = 2

# rdm_als_scale_t.src_kind
SRC_RAW_F64, SRC_RAW_F32, SRC_VAL_F32, SRC_VAL_F64, SRC_MAP_F32 = range(5)
# rdm_als_scale_t.flags and rdm_als_fused_phases masks (include/rdm_b200.h)
ALS_DENSE_ONLY, ALS_TRUE_TRANSPOSE, ALS_TRUE_GM, ALS_CORRECT_TILING, ALS_PAGES_ONE_CTA, ALS_PAGES_CLUSTER, ALS_SKIP_UNUSED_PAGES = 1, 2, 4, 8, 16, 32, 64
PHASE_SPARSIFY, PHASE_PAGES, PHASE_DENSE, PHASE_ALL = 1, 2, 4, 7
# rdm_quick_gm / rdm_gm_normalize dtype
DT_F32, DT_F64, DT_I64 = range(3)


class AlsScale(Structure):
    """rdm_als_scale_t"""
    _fields_ = [
        ("src", c_void_p), ("src_kind", c_int32), ("rows", c_int32), ("pages", c_int32), ("side", c_int32),
        ("limit", c_int32), ("flags", c_int32),
        ("thresholds", c_void_p), ("levels", c_void_p), ("bins_out", c_void_p), ("values_out", c_void_p),
        ("pages_out", c_void_p), ("map_out", c_void_p), ("ws", c_void_p), ("record_out", c_void_p),
        ("kstar_out", c_void_p),
    ]


# symbol -> (restype, argtypes); every symbol include/rdm_b200.h declares
PROTOTYPES = {
    "rdm_abi_version": (c_int, []),
    "rdm_last_error": (c_char_p, []),
    "rdm_pair_v1_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "rdm_resize_half": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "rdm_pair_id_f64": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdm_pair_pages_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "rdm_resize_bicubic_f64": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_upsample_nearest_f64": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_lloyd_quantize_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_lloyd_quantize_f64": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_als_fused": (c_int, [POINTER(AlsScale), c_int32, c_int64, c_int32, c_void_p]),
    "rdm_als_fused_phases": (c_int, [POINTER(AlsScale), c_int32, c_int64, c_int32, c_int32, c_void_p]),
    "rdm_als_ws_floats": (c_int64, [c_int32, c_int32, c_int32]),
    "rdm_sparsify_geometry": (c_int, [c_void_p, c_void_p, c_void_p]),
    "rdm_conv_head_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, POINTER(AlsScale), c_void_p]),
    "rdm_als_step_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p, c_void_p]),
    "rdm_quick_gm": (c_int, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_void_p, c_void_p]),
    "rdm_gm_normalize": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "rdm_decompose": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_pyramid_len": (c_int64, [c_int32, c_int32]),
    "rdm_gt_prepare": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "rdm_decompose_bwd": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdm_gm_bwd": (c_int, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_log_stack_bwd": (c_int, [POINTER(c_void_p), c_int32, c_int64, c_int64, c_void_p, POINTER(c_void_p), c_void_p]),
    "rdm_log_stack_f64": (c_int, [POINTER(c_void_p), c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "rdm_make_pred_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p]),
    "rdm_make_pred_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "rdm_recombination_f64": (c_int, [POINTER(c_void_p), POINTER(c_int32), c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "rdm_recombination_bwd": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int32), c_int32, c_int32, c_int64, c_int32, c_void_p]),
    "rdm_fuse_tail": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int32), c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                              POINTER(c_void_p), c_void_p]),
    "rdm_fuse_tail_bands": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int32), c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                    POINTER(c_void_p), c_int32, c_void_p]),
    "rdm_fuse_tail_weight_count": (c_int64, [POINTER(c_int32), c_int32]),
    "rdm_fuse_tail_bwd": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int32), c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "rdm_component_loss": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "rdm_dorn_regression_f32": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdm_dorn_regression_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_ordinal_loss_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdm_ordinal_loss_ws_doubles": (c_int64, []),
    "rdm_depth2label_sid": (c_int, [c_void_p, c_int32, c_int64, c_double, c_double, c_double, c_void_p, c_void_p]),
    "rdm_ordinal_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
}

_lib = None


class RdmError(RuntimeError):
    pass


def load():
    """Load librdm_b200.so once; raise loudly if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RdmError(
            f"{LIB_PATH} not found: build it with `python -m md_rdm_b200.build` (needs nvcc). "
            "md_rdm_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.rdm_abi_version()
    if got != ABI_VERSION:
        raise RdmError(f"librdm_b200.so ABI version {got}, host code expects {ABI_VERSION}: rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """0 = enqueued; <0 argument error; >0 cudaError_t.  The message is thread-local in the library."""
    if rc != 0:
        msg = load().rdm_last_error()
        raise RdmError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def i32_array(vals):
    return (c_int32 * len(vals))(*vals)
