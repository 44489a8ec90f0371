"""ctypes binding of ``libovqa_b200.so`` (the C ABI declared in ``include/ovqa.h``).

There is no CPU fallback: if the library is missing, or an entry point returns a non-zero code, this module
raises.  PyTorch is only used by the callers for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libovqa_b200.so")

F32, BF16 = 0, 1
OP_N, OP_T = 0, 1
EPI_GELU, EPI_DGELU, EPI_SAVE_DGELU, EPI_MUL_AUX, EPI_RELU = 1, 2, 4, 8, 16
ERR_BAD_SHAPE = -2

_vp, _i64, _i32, _f32, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "ovqa_version": [],
    "ovqa_error_string": [_i32],
    "ovqa_last_cuda_error": [],
    "ovqa_device_sm_count": [],
    "ovqa_set_sm_reserve": [_i32],
    "ovqa_launch_count": [],
    "ovqa_gemm": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _i32,
                  _vp, _vp, _i64, _vp, _i64, _vp, _i64, _f32, _u64, _u64, _vp, _vp, _i32, _vp],
    "ovqa_gemm_bf16_simt": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _i32, _vp],
    "ovqa_cast_rows_padmask": [_vp, _vp, _vp, _i64, _i64, _i64, _f32, _vp],
    "ovqa_rows_padmask_bf16": [_vp, _vp, _i64, _i64, _i64, _f32, _vp],
    "ovqa_mha_decode_step": [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _i64, _i64,
                             _vp, _i64, _i64, _i64, _f32, _i32, _vp, _i64, _vp, _vp, _vp],
    "ovqa_mha_decode_workspace_bytes": [_i64] * 7 + [_i32],
    "ovqa_mha_block_infer_workspace_bytes": [_i64] * 7 + [_i32, _i32],
    "ovqa_mha_block_infer": [_vp] * 6 + [_i32] + [_i64] * 7 + [_vp] * 6 + [_f32, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _f32, _f32, _i32, _vp, _i64,
                             _vp, _vp, _vp],
    "ovqa_pwff_infer_workspace_bytes": [_i64, _i64, _i64, _i32],
    "ovqa_topk_rows_workspace_bytes": [_i64, _i64, _i32],
    "ovqa_topk_rows": [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp],
    "ovqa_pwff_infer": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _i64, _vp, _vp, _vp],
    "ovqa_mul_to_bf16": [_vp, _vp, _i32, _vp, _i64, _vp],
    "ovqa_sdpa_fwd_prefix_lm": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _f32, _f32, _i32,
                                _vp, _vp],
    "ovqa_sdpa_fwd": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                      _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _vp, _vp, _f32, _u64, _u64, _vp, _vp],
    "ovqa_sdpa_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64,
                      _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                      _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _u64, _u64, _vp, _vp],
    "ovqa_geo_bias_fwd": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "ovqa_geo_bias_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "ovqa_layernorm_fwd": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _vp],
    "ovqa_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _u64, _u64, _vp, _vp],
    "ovqa_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "ovqa_cast_bf16_to_f32": [_vp, _vp, _i64, _vp],
    "ovqa_colsum": [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp],
    "ovqa_add_f32": [_vp, _vp, _vp, _i64, _vp],
    "ovqa_dropout_apply": [_vp, _vp, _i64, _f32, _u64, _u64, _vp, _i32, _vp],
    "ovqa_advance_u64": [_vp, _u64, _vp],
    "ovqa_stage_many": [_vp, _i64, _vp],
    "ovqa_box_geometry_bias": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "ovqa_ptr_scores": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _i32, _i32, _vp],
    "ovqa_gemm_ln": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _u64, _u64, _vp, _vp],
    "ovqa_lstm_supported": [_i64],
    "ovqa_lstm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "ovqa_lstm_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "ovqa_attr_pool_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "ovqa_attr_pool_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _i32, _vp],
    "ovqa_ptr_scores_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _i32, _vp],
    "ovqa_prev_pred_embed": [_vp] * 12 + [_i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp],
    "ovqa_greedy_step": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp],
    "ovqa_dropout_mask": [_vp, _i64, _f32, _u64, _u64, _vp, _vp],
}
_RESTYPES = {"ovqa_error_string": C.c_char_p, "ovqa_launch_count": C.c_uint64, "ovqa_mha_decode_workspace_bytes": C.c_int64,
             "ovqa_pwff_infer_workspace_bytes": C.c_int64, "ovqa_topk_rows_workspace_bytes": C.c_int64, "ovqa_mha_block_infer_workspace_bytes": C.c_int64}

_lib = None
_lock = threading.Lock()


def launch_count() -> int:
    """Kernels launched by the library so far, counted inside the library at each launch site
    (bench.py reports the difference over the timed region as ``gpu_launches``)."""
    return int(lib().ovqa_launch_count())


class OvqaError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OvqaError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  openvivqa_b200 has no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = l
    return _lib


def declared_symbols():
    return list(_PROTOS)


def check(code: int, what: str):
    if code != 0:
        l = lib()
        msg = l.ovqa_error_string(code).decode()
        extra = f" (cudaError {l.ovqa_last_cuda_error()})" if code == -4 else ""
        raise OvqaError(f"{what}: {msg}{extra}")


def call(name: str, *args):
    check(getattr(lib(), name)(*args), name)
