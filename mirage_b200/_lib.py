"""ctypes binding of libmirage_b200.so (see include/mirage_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libmirage_b200.so"

# enums mirrored from include/mirage_b200.h
MB_BF16, MB_F32 = 0, 1
MB_MAJOR_K, MB_MAJOR_MN, MB_A_PATCH32 = 0, 1, 2
MB_EPI_GELU, MB_EPI_DGELU, MB_EPI_ATOMIC, MB_EPI_UNPATCH = 1, 2, 4, 8


class MirageB200Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("aux_in", C.c_void_p), ("aux_out", C.c_void_p),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("ld_res", C.c_int64), ("ld_aux", C.c_int64),
        ("res_period", C.c_int64),
        ("a_layout", C.c_int32), ("b_layout", C.c_int32),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("epilogue", C.c_int32), ("k_splits", C.c_int32), ("block_n", C.c_int32),
        ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("out_row_period", C.c_int64), ("out_row_stride", C.c_int64), ("out_row_offset", C.c_int64),
        ("up_channels", C.c_int32), ("up_ph", C.c_int32), ("up_pw", C.c_int32),
        ("up_gh", C.c_int32), ("up_gw", C.c_int32),
        ("cta_pair", C.c_int32),
        ("colsum_out", C.c_void_p),
        ("twin_out", C.c_void_p), ("ld_twin", C.c_int64), ("row_stats", C.c_void_p),
        ("ln_stats", C.c_void_p), ("ln_c1", C.c_void_p), ("ln_eps", C.c_float), ("ln_parts", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("lse", C.c_void_p),
        ("batch", C.c_int64), ("heads", C.c_int64), ("nq", C.c_int64), ("nk", C.c_int64),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
        ("head_dim", C.c_int32), ("scale", C.c_float),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("d_out", C.c_void_p), ("lse", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("workspace", C.c_void_p),
        ("batch", C.c_int64), ("heads", C.c_int64), ("nq", C.c_int64), ("nk", C.c_int64),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
        ("lddo", C.c_int64), ("lddq", C.c_int64), ("lddk", C.c_int64), ("lddv", C.c_int64),
        ("head_dim", C.c_int32), ("scale", C.c_float),
    ]


class OptimSegment(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("shadow", C.c_void_p), ("numel", C.c_int64),
                ("group", C.c_int32), ("flags", C.c_int32)]


class OptimHyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("lr_scale", C.c_float), ("weight_decay", C.c_float), ("pad_", C.c_float)]


class OptimState(C.Structure):
    _fields_ = [("step", C.c_int64), ("bias_corr1", C.c_float), ("bias_corr2_sqrt", C.c_float),
                ("clip_coef", C.c_float), ("grad_norm", C.c_float), ("skipped", C.c_int32), ("pad_", C.c_int32)]


_lib = None


def _declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.mb_last_error.restype = C.c_char_p
    lib.mb_last_error.argtypes = []
    lib.mb_version.restype = C.c_int
    lib.mb_sm_count.restype = C.c_int
    lib.mb_set_sm_reserve.restype = C.c_int
    lib.mb_set_sm_reserve.argtypes = [C.c_int]
    lib.mb_set_pdl.restype = C.c_int
    lib.mb_set_pdl.argtypes = [C.c_int]
    lib.mb_clear_tensor_map_cache.restype = None
    lib.mb_gemm.restype = C.c_int
    lib.mb_gemm.argtypes = [C.POINTER(GemmArgs), vp]
    lib.mb_attn_fwd.restype = C.c_int
    lib.mb_attn_fwd.argtypes = [C.POINTER(AttnArgs), vp]
    lib.mb_attn_bwd.restype = C.c_int
    lib.mb_attn_bwd.argtypes = [C.POINTER(AttnBwdArgs), vp]
    lib.mb_attn_bwd_workspace.restype = C.c_int64
    lib.mb_attn_bwd_workspace.argtypes = [i64, i64, i64, i64, i32]
    lib.mb_masked_loss_workspace.restype = C.c_int64
    lib.mb_masked_loss_workspace.argtypes = [i64, i64, i64]
    lib.mb_optim_blocks.restype = C.c_int64
    lib.mb_optim_blocks.argtypes = [i64]
    lib.mb_gemm_ln_parts.restype = C.c_int
    lib.mb_gemm_ln_parts.argtypes = [i64, i64]
    lib.mb_class_colsum_workspace.restype = C.c_int64
    lib.mb_class_colsum_workspace.argtypes = [i64, i64, i32]
    for name in ("mb_layernorm_bwd_workspace", "mb_colsum_workspace", "mb_ln_meanpool_workspace"):
        getattr(lib, name).restype = C.c_int64
        getattr(lib, name).argtypes = [i64, i64]
    # the remaining entry points are declared by signature table so the loader and the symbol
    # test (tests/test_abi.py) share one source of truth
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> argtypes (all return int).  Filled in as kernels are added; order matches the header.
SIGNATURES: dict[str, list] = {
    "mb_layernorm_fwd": [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp],
    "mb_layernorm_bwd": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                         _i64, _i64, _i64, _i64, _i64, _vp, _i32, _vp],
    "mb_colsum": [_vp, _i32, _vp, _i32, _vp, _i64, _i64, _i64, _vp],
    "mb_token_gather_fwd": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    "mb_token_gather_bwd": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    "mb_fill_global_rows": [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    "mb_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "mb_semseg_patches": [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "mb_class_emb_grad": [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "mb_semseg_patches_rows": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "mb_class_emb_grad_rows": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "mb_visible_rows": [_vp, _i64, _i64, _i64, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _vp],
    "mb_embed_rows_init": [_vp, _vp, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp,
                           _vp, _i64, _i64, _vp],
    "mb_gather_patches32": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "mb_class_colsum": [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp],
    "mb_dec_assemble_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp],
    "mb_dec_assemble_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64,
                            _i64, _vp],
    "mb_patchify_cast": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
    "mb_masked_mse_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "mb_masked_mse_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "mb_masked_ce_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _vp],
    "mb_masked_ce_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _vp],
    "mb_ln_meanpool_fwd": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _f32, _vp],
    "mb_ln_meanpool_bwd": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _i64, _i64,
                           _i64, _vp],
    "mb_sumsq": [_vp, _i64, _vp, _i32, _vp],
    "mb_optim_prepare": [_vp, _vp, _i32, _f32, _f32, _f32, _f32, _vp],
    "mb_adamw_step": [_vp, _vp, _i32, _i64, _vp, _vp, _f32, _f32, _f32, _i32, _vp, _vp],
    "mb_optim_finish": [_vp, _vp, _i64, _vp],
    "mb_sample_masks": [C.c_uint64, _vp, _vp, C.POINTER(C.c_int32), C.POINTER(C.c_float), _i32, _i64, _i64, _i32,
                        _vp, _vp, _vp, _vp],
    "mb_augment_image": [_vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "mb_augment_labels": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp],
}

# every symbol include/mirage_b200.h declares
EXPORTED = ["mb_last_error", "mb_version", "mb_sm_count", "mb_set_sm_reserve", "mb_set_pdl", "mb_clear_tensor_map_cache", "mb_gemm",
            "mb_attn_fwd", "mb_attn_bwd", "mb_attn_bwd_workspace", "mb_layernorm_bwd_workspace",
            "mb_colsum_workspace", "mb_masked_loss_workspace", "mb_ln_meanpool_workspace", "mb_optim_blocks",
            "mb_class_colsum_workspace", "mb_gemm_ln_parts"]


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MirageB200Error(
                f"{LIB_PATH} is missing: build it with `python -m mirage_b200.build` "
                "(there is no CPU or PyTorch fallback for the MIRAGE hot path)")
        handle = C.CDLL(str(LIB_PATH), mode=os.RTLD_NOW | os.RTLD_LOCAL)
        _declare(handle)
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().mb_last_error().decode(errors="replace")
        raise MirageB200Error(f"{what} failed (rc={rc}): {msg}")


def exported_symbols():
    return list(EXPORTED) + list(SIGNATURES.keys())
