"""ctypes binding of the C-ABI library (include/capdec_b200.h).

There is no CPU fallback: if the sm_100a library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcapdec_b200.so"

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_u32 = C.c_uint32
_f = C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "capdec_last_error": [],
    "capdec_version": [],
    "capdec_launch_count": [],
    "capdec_gemm_tf32": [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i, _p, _i, _p, _i, _i, _p, _p, _i, _i, _p],
    "capdec_gemm_tf32_ex": [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i, _p, _i, _p, _i, _i, _p, _p, _i, _i, _p, _p, _p],
    "capdec_gemm_tf32_mul": [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i, _p, _i, _p, _i, _p, _p],
    "capdec_gemm_tf32_mul_ex": [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i, _p, _i, _p, _i, _p, _i, _p],
    "capdec_gemm_debug_mn_encoding": [_i, _i, _i, _i],
    "capdec_gemm_debug_force_pair": [_i],
    "capdec_gemm_debug_trace": [_p],
    "capdec_gemm_set_row_hint": [_i],
    "capdec_gemm_set_schedule": [_i],
    "capdec_gemm_autotune": [_i],
    "capdec_gemm_plan_query": [_i, _i, _i, _i, _i, _i, _i, _i],
    "capdec_gemm_fp32_simt": [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i, _p, _i, _p, _i, _p],
    "capdec_split_tf32": [_p, _p, _p, _i64, _p],
    "capdec_noise_injection": [_p, _p, _i, _i, _f, _p, _p, _i, _i, _p, _u64, _p],
    "capdec_embed_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _u32, _p],
    "capdec_embed_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _u32, _p],
    "capdec_add_ln_fwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _p, _u32, _p, _p],
    "capdec_add_ln_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _p, _u32, _p, _p],
    "capdec_attention_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _p,
                             _f, _p, _u32, _p],
    "capdec_attention_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64,
                             _i64, _f, _i, _p, _f, _p, _u32, _p],
    "capdec_attention_tc_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _p,
                                _f, _p, _u32, _p, _p],
    "capdec_attention_tc_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64,
                                _i64, _f, _i, _p, _f, _p, _u32, _p, _p],
    "capdec_attention_tc_fwd_x3": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _f, _i, _p,
                                   _f, _p, _u32, _p, _p],
    "capdec_attention_tc_bwd_x3": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64,
                                   _i64, _f, _i, _p, _f, _p, _u32, _p, _p],
    "capdec_ce_count": [_p, _i64, _i64, _p, _p, _p],
    "capdec_ce_fwd_bwd": [_p, _i64, _p, _i, _i, _i64, _p, _f, _p, _i, _p, _p],
    "capdec_compact_targets": [_p, _i, _i, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p],
    "capdec_rows_gather_idx": [_p, _p, _p, _p, _i, _i, _p],
    "capdec_rows_scatter_idx": [_p, _p, _p, _i, _i, _p],
    "capdec_colsum_acc": [_p, _i64, _p, _i, _i, _p],
    "capdec_act_bwd": [_p, _p, _p, _p, _i, _i, _i, _p],
    "capdec_rows_gather": [_p, _p, _i, _i, _i, _i, _i, _p],
    "capdec_rows_scatter": [_p, _p, _i, _i, _i, _i, _i, _p],
    "capdec_mapper_concat_fwd": [_p, _p, _p, _i, _i, _i, _i, _p],
    "capdec_mapper_concat_bwd": [_p, _p, _p, _i, _i, _i, _i, _p],
    "capdec_beam_init": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "capdec_kv_prefill": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "capdec_decode_embed": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "capdec_decode_attention": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p],
    "capdec_row_topk": [_p, _i64, _i, _i, _f, _i, _p, _p, _p, _p],
    "capdec_beam_select": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "capdec_pack_plan": [_p, _i, _i, _i, _p, _p, _p, _p],
    "capdec_embed_fwd_packed": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _u32, _p],
    "capdec_embed_bwd_packed": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _u32, _p],
    "capdec_zero_tail_rows": [_p, _i64, _p, _p],
    "capdec_batch_gather": [_p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "capdec_zero_fill": [_p, _i64, _p],
    "capdec_step_clock": [_p, _p, _p, _p, _f, _i, _i, _p],
    "capdec_adamw_step": [_p, _p, _p, _p, _i64, _p, _p, _f, _f, _f, _f, _p, _i, _p],
    "capdec_peer_export": [_p, _p, _p],
    "capdec_peer_open": [_p, _i64, _p],
    "capdec_peer_close": [_p, _i64],
    "capdec_copy_async": [_p, _p, _i64, _p],
    "capdec_adamw_peer_step": [_p, _p, _i, _i, _i64, _i64, _p, _p, _p, _p, _f, _f, _f, _f, _p, _p],
}
_RESTYPES = {"capdec_last_error": C.c_char_p, "capdec_launch_count": C.c_int64, "capdec_gemm_debug_mn_encoding": None,
             "capdec_gemm_debug_force_pair": None, "capdec_gemm_debug_trace": None, "capdec_gemm_set_row_hint": None}


class CapdecError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libcapdec_b200.so (building it in-tree with nvcc if absent). Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing or os.environ.get("CAPDEC_NO_BUILD"):
            raise CapdecError(f"{LIB_PATH} is missing (run `python -m capdec_b200.build`); there is no CPU fallback")
        from . import build as _build
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().capdec_last_error()
        raise CapdecError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().capdec_launch_count())
