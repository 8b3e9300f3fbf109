"""ctypes binding of ``libgenhancer_b200.so`` (the C ABI declared in ``include/genhancer_b200.h``).

The product path has **no** CPU or PyTorch fallback: if the shared library is missing or a
call fails, we raise.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C genhancer_b200/csrc``) so that it travels with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GH_LIB_PATH") or os.path.join(_HERE, "libgenhancer_b200.so")   # (override: A/B builds)

GH_BF16, GH_F32 = 0, 1
ACT_NONE, ACT_GELU_TANH, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_SILU = 0, 1, 2, 3, 4


class GhError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("b", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("d", C.c_void_p), ("ldd", C.c_int64), ("d_dtype", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("bias_dtype", C.c_int32),
        ("act", C.c_int32), ("act_grad", C.c_int32),
        ("aux_in", C.c_void_p), ("ld_aux_in", C.c_int64),
        ("aux_out", C.c_void_p), ("ld_aux_out", C.c_int64),
        ("gate", C.c_void_p), ("gate_ld", C.c_int64), ("rows_per_batch", C.c_int32),
        ("residual", C.c_void_p), ("ld_res", C.c_int64), ("res_dtype", C.c_int32),
        ("a2", C.c_void_p), ("lda2", C.c_int64), ("b2", C.c_void_p), ("ldb2", C.c_int64), ("K2", C.c_int32),
        ("k_splits", C.c_int32),
        ("batch", C.c_int32), ("a_batch_rows", C.c_int64), ("b_batch_rows", C.c_int64), ("d_batch_rows", C.c_int64),
        ("dynamic_tiles", C.c_int32),
    ]


class CopyDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("src_ld", C.c_int64), ("dst_ld", C.c_int64), ("src_dtype", C.c_int32), ("dst_dtype", C.c_int32),
                ("src_row_group", C.c_int32), ("src_row_pitch", C.c_int32),
                ("src_col_group", C.c_int32), ("src_col_pitch", C.c_int32),
                ("dst_row_group", C.c_int32), ("dst_row_pitch", C.c_int32),
                ("dst_col_group", C.c_int32), ("dst_col_pitch", C.c_int32),
                ("scale", C.c_float), ("accumulate", C.c_int32)]


class RowsView(C.Structure):
    _fields_ = [("rows_per_batch", C.c_int32), ("batch_stride", C.c_int64), ("row_stride", C.c_int64)]


class AttnTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_int64), ("head_stride", C.c_int64),
                ("row_stride", C.c_int64)]


class AttnOut(C.Structure):
    _fields_ = [("seg0", C.c_void_p), ("seg0_batch_stride", C.c_int64), ("seg0_row_stride", C.c_int64),
                ("seg1", C.c_void_p), ("seg1_batch_stride", C.c_int64), ("seg1_row_stride", C.c_int64),
                ("n_split", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("y", C.c_void_p), ("y_dtype", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("Ho", C.c_int32), ("Wo", C.c_int32),
                ("bias", C.c_void_p), ("bias_dtype", C.c_int32), ("act", C.c_int32),
                ("residual", C.c_void_p), ("res_dtype", C.c_int32)]


_vp, _i64, _i32, _f32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_double
_rv, _at, _ao = C.POINTER(RowsView), C.POINTER(AttnTensor), C.POINTER(AttnOut)

# Every symbol include/genhancer_b200.h declares, with its argument types.
# tests/test_abi.py checks that the built library exports exactly these.
SIGNATURES: dict[str, list] = {
    "gh_last_error": [],
    "gh_version": [],
    "gh_init": [C.c_int],
    "gh_gemm_bf16": [C.POINTER(GemmArgs), _vp],
    "gh_debug_gemm_prof": [_vp],
    "gh_fm_interp_fwd": [_vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "gh_fm_mse_loss_fwdbwd": [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _vp],
    "gh_layernorm_fwd": [_vp, _rv, _vp, _rv, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp, _vp],
    "gh_layernorm_bwd_dx": [_vp, _rv, _vp, _rv, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _rv, _vp, _rv, _vp],
    "gh_layernorm_bwd_params": [_vp, _rv, _vp, _rv, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp],
    "gh_gate_bwd": [_vp, _rv, _vp, _rv, _i32, _i32, _vp, _i64, _vp, _rv, _vp, _i64, _vp, _vp],
    "gh_colsum": [_vp, _rv, _i32, _i32, _vp, _i64, _vp],
    "gh_rope_table": [_vp, _vp, _i64, _i32, _i32, _i32, _f64, _vp],
    "gh_qk_norm_rope_fwd": [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "gh_qk_norm_rope_bwd": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                            _i64, _vp, _vp, _vp],
    "gh_timestep_embedding": [_vp, _vp, _i32, _i32, _vp],
    "gh_act_fwd": [_vp, _vp, _i64, _i32, _vp],
    "gh_act_bwd": [_vp, _vp, _vp, _i64, _i32, _vp],
    "gh_accum_cast": [_vp, _vp, _i32, _i64, _f32, _i32, _vp],
    "gh_euler_cfg_step": [_vp, _vp, _vp, _f32, _f32, _i64, _vp],
    "gh_batched_copy": [_vp, _i32, _i32, _vp],
    "gh_dropout_fwd": [_vp, _vp, _i64, _f32, C.c_uint64, C.c_uint64, _vp, _vp],
    "gh_dropout_bwd_add": [_vp, _vp, _i64, _f32, C.c_uint64, C.c_uint64, _vp, _vp],
    "gh_lora_dropout_fwd": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _f32, _f32, C.c_uint64, C.c_uint64, _vp, _vp],
    "gh_lora_dropout_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _f32, C.c_uint64, C.c_uint64, _vp, _vp, _i32, _vp],
    "gh_flash_attn_fwd": [_at, _at, _at, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _ao, _vp, _vp],
    "gh_debug_attn_prof": [_vp],
    "gh_flash_attn_bwd_workspace_bytes": [_i32, _i32, _i32, _i32, _i32],
    "gh_flash_attn_bwd": [_at, _at, _at, _ao, _ao, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _at, _at, _at, _vp, _vp,
                          _vp],
    "gh_conv2d_nhwc": [C.POINTER(ConvArgs), _vp],
    "gh_patch_im2col": [_vp, _vp, _i32, _i32, _i32, _i64, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp],
    "gh_im2col3x3_c3": [_vp, _vp, _i32, _i32, _i32, _f32, _f32, _vp],
    "gh_patch_embed_fwd": [_vp, _i32, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, C.POINTER(C.c_float),
                           C.POINTER(C.c_float), _vp],
    "gh_patch_im2col_u8hwc": [_vp, _vp, _i32, _i32, _i32, _i64, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp],
    "gh_im2col3x3_c3_u8hwc": [_vp, _vp, _i32, _i32, _i32, _f32, _f32, _vp],
    "gh_embed_assemble": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "gh_groupnorm_ws_bytes": [_i32, _i64],
    "gh_groupnorm_swish_nhwc": [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _f32, _i32, _vp, _vp],
    "gh_upsample2x_nhwc": [_vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "gh_u8hwc_to_f32chw": [_vp, _vp, _i32, _i32, _i32, _vp],
    "gh_softmax_rows": [_vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp],
    "gh_ae_sample_patchify": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp],
    "gh_sumsq_workspace_bytes": [],
    "gh_sumsq_accum": [_vp, _i32, _i64, _vp, _vp, _vp],
    "gh_adamw_step": [_vp, _vp, _vp, _vp, _i32, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _f32, _f32, _vp, _vp],
}

_lib = None
_lock = threading.Lock()
_inited_devices: set[int] = set()


def _declare(lib):
    for name, argtypes in SIGNATURES.items():
        if os.environ.get("GH_LIB_PATH") and not hasattr(lib, name):
            continue   # an A/B build of an OLDER source (same-box comparisons): entry points added since are absent
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = (C.c_char_p if name == "gh_last_error" else C.c_int64 if name.endswith(("_ws_bytes", "_workspace_bytes")) else C.c_int)


def lib():
    """Load (once) and return the ctypes handle. Raises if the extension is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise GhError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / PyTorch fallback on the product path)")
                h = C.CDLL(LIB_PATH)
                _declare(h)
                _lib = h
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise GhError(f"genhancer_b200 error {status}: {lib().gh_last_error().decode()}")


def init(device_index: int) -> None:
    if device_index not in _inited_devices:
        check(lib().gh_init(int(device_index)))
        _inited_devices.add(device_index)
