"""ctypes binding of libst_b200.so (include/st_b200.h).

There is no CPU fallback: if the shared library is missing, or the current device is not sm_100,
every operator raises.  The structures below mirror the C structs field for field.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ST_B200_LIB") or os.path.join(HERE, "libst_b200.so")   # the override serves A/B runs of two builds

c_float_p = C.c_void_p  # raw device pointers travel as integers

# ST_DTYPE_* (include/st_b200.h): element type of activation tensors
DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_F32_H16 = 0, 1, 2, 3
i64 = C.c_int64
u64 = C.c_uint64


class GemmEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("aux", C.c_void_p), ("ldaux", i64), ("aux_mode", C.c_int),
                ("relu", C.c_int), ("round_tf32", C.c_int), ("k_splits", C.c_int),
                ("dropout_p", C.c_float), ("seed", u64)]


class AttnArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("dk", C.c_int),
                ("q", C.c_void_p), ("ldq", i64), ("k", C.c_void_p), ("ldk", i64), ("v", C.c_void_p), ("ldv", i64),
                ("mask", C.c_void_p), ("ms_b", i64), ("ms_q", i64), ("ms_k", i64),
                ("dropout_p", C.c_float), ("seed", u64),
                ("ctx", C.c_void_p), ("ldctx", i64), ("lse", C.c_void_p), ("attn", C.c_void_p),
                ("dtype", C.c_int), ("k_len", C.c_void_p), ("causal", C.c_int)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("f", AttnArgs), ("dctx", C.c_void_p), ("lddctx", i64), ("delta", C.c_void_p),
                ("dq", C.c_void_p), ("lddq", i64), ("dk", C.c_void_p), ("lddk", i64), ("dv", C.c_void_p), ("lddv", i64)]


class MhaArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("H", C.c_int), ("d_model", C.c_int), ("dk", C.c_int),
                ("q_in", C.c_void_p), ("k_in", C.c_void_p), ("v_in", C.c_void_p), ("residual", C.c_void_p),
                ("wq", C.c_void_p), ("bq", C.c_void_p), ("wk", C.c_void_p), ("bk", C.c_void_p),
                ("wv", C.c_void_p), ("bv", C.c_void_p), ("wo", C.c_void_p), ("bo", C.c_void_p),
                ("ln_g", C.c_void_p), ("ln_b", C.c_void_p),
                ("mask", C.c_void_p), ("ms_b", i64), ("ms_q", i64), ("ms_k", i64),
                ("eps", C.c_float), ("dropout_p", C.c_float), ("seed", u64),
                ("inputs_tf32", C.c_int), ("round_out", C.c_int),
                ("out", C.c_void_p), ("attn", C.c_void_p),
                ("saved", C.c_void_p), ("saved_floats", i64), ("ws", C.c_void_p), ("ws_floats", i64),
                ("wq_tf32", C.c_void_p), ("wk_tf32", C.c_void_p), ("wv_tf32", C.c_void_p), ("wo_tf32", C.c_void_p),
                ("dtype", C.c_int), ("k_len", C.c_void_p), ("causal", C.c_int),
                ("q_h16", C.c_void_p), ("k_h16", C.c_void_p), ("v_h16", C.c_void_p), ("out_h16", C.c_void_p)]


class MhaBwdArgs(C.Structure):
    _fields_ = [("f", MhaArgs), ("dout", C.c_void_p),
                ("dq_in", C.c_void_p), ("dk_in", C.c_void_p), ("dv_in", C.c_void_p), ("dresidual", C.c_void_p),
                ("dwq", C.c_void_p), ("dbq", C.c_void_p), ("dwk", C.c_void_p), ("dbk", C.c_void_p),
                ("dwv", C.c_void_p), ("dbv", C.c_void_p), ("dwo", C.c_void_p), ("dbo", C.c_void_p),
                ("dln_g", C.c_void_p), ("dln_b", C.c_void_p), ("grads_zeroed", C.c_int),
                ("dout_amax", C.c_void_p), ("dq_amax", C.c_void_p)]


class FfnArgs(C.Structure):
    _fields_ = [("rows", i64), ("d_model", C.c_int), ("d_ff", C.c_int),
                ("x", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("ln_g", C.c_void_p), ("ln_b", C.c_void_p),
                ("eps", C.c_float), ("dropout_p", C.c_float), ("seed", u64),
                ("x_is_tf32", C.c_int), ("round_out", C.c_int),
                ("out", C.c_void_p), ("saved", C.c_void_p), ("saved_floats", i64), ("ws", C.c_void_p), ("ws_floats", i64),
                ("w1_tf32", C.c_void_p), ("w2_tf32", C.c_void_p), ("dtype", C.c_int),
                ("x_h16", C.c_void_p), ("out_h16", C.c_void_p)]


class FfnBwdArgs(C.Structure):
    _fields_ = [("f", FfnArgs), ("dout", C.c_void_p), ("dx", C.c_void_p),
                ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p),
                ("dln_g", C.c_void_p), ("dln_b", C.c_void_p), ("grads_zeroed", C.c_int),
                ("dout_amax", C.c_void_p), ("dx_amax", C.c_void_p)]


class FrontendArgs(C.Structure):
    _fields_ = [("rows", i64), ("T", C.c_int), ("in_dim", C.c_int), ("d_model", C.c_int),
                ("x", C.c_void_p), ("w", C.c_void_p), ("b", C.c_void_p), ("ln_g", C.c_void_p), ("ln_b", C.c_void_p),
                ("pe", C.c_void_p), ("eps", C.c_float), ("dropout_p", C.c_float), ("seed", u64), ("round_out", C.c_int),
                ("out", C.c_void_p), ("saved", C.c_void_p), ("saved_floats", i64), ("ws", C.c_void_p), ("ws_floats", i64),
                ("dtype", C.c_int)]


class FrontendBwdArgs(C.Structure):
    _fields_ = [("f", FrontendArgs), ("dout", C.c_void_p), ("dx", C.c_void_p), ("dw", C.c_void_p), ("db", C.c_void_p),
                ("dln_g", C.c_void_p), ("dln_b", C.c_void_p), ("grads_zeroed", C.c_int)]


class LinearArgs(C.Structure):
    _fields_ = [("rows", i64), ("in_dim", C.c_int), ("out_dim", C.c_int), ("x", C.c_void_p), ("x_is_tf32", C.c_int),
                ("w", C.c_void_p), ("b", C.c_void_p), ("y", C.c_void_p), ("ldy", i64),
                ("saved", C.c_void_p), ("saved_floats", i64), ("ws", C.c_void_p), ("ws_floats", i64), ("dtype", C.c_int)]


class LinearBwdArgs(C.Structure):
    _fields_ = [("f", LinearArgs), ("dy", C.c_void_p), ("lddy", i64), ("dx", C.c_void_p), ("dw", C.c_void_p),
                ("db", C.c_void_p), ("grads_zeroed", C.c_int)]


class AdamArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", i64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("step", C.c_int), ("max_grad_norm", C.c_float), ("grad_scale", C.c_float),
                ("norm_ws", C.c_void_p), ("param_tf32", C.c_void_p), ("twin_dtype", C.c_int)]


# every symbol include/st_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_S = C.c_void_p  # cudaStream_t
SIGNATURES = {
    "st_version": (C.c_int, []),
    "st_last_error": (C.c_char_p, []),
    "st_device_check": (C.c_int, [C.c_int]),
    "st_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "st_launch_count": (i64, []),
    "st_profile_enable": (C.c_int, [C.c_int]),
    "st_profile_reset": (C.c_int, []),
    "st_profile_dump": (C.c_int, [C.c_char_p]),
    "st_profile_classes": (C.c_int, []),
    "st_profile_class_name": (C.c_char_p, [C.c_int]),
    "st_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]),
    "st_debug_read_trace": (C.c_int, [C.POINTER(u64), C.c_int]),
    "st_debug_read_fwd_trace": (C.c_int, [C.POINTER(u64), C.c_int]),
    "st_debug_mma_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "st_selftest_count": (C.c_int, []),
    "st_selftest": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "st_add_ln_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, i64, C.c_int, C.c_float, C.c_int, C.c_float, u64, _S]),
    "st_add_ln_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, i64, C.c_int, C.c_int, C.c_float, u64, _S]),
    "st_lsce_fwd_bwd": (C.c_int, [_P, i64, _P, _P, _P, C.c_float, i64, C.c_int, i64, C.c_int, _P, _P, _P, i64, _S]),
    "st_softce_fwd_bwd": (C.c_int, [_P, i64, _P, _P, C.c_int, i64, C.c_int, _P, _P, _P, i64, _S]),
    "st_round_tf32": (C.c_int, [_P, i64, _P, i64, i64, C.c_int, _S]),
    "st_colsum_add": (C.c_int, [_P, i64, i64, C.c_int, _P, _S]),
    "st_gemm": (C.c_int, [C.c_int, _P, i64, _P, i64, _P, i64, C.c_int, C.c_int, C.c_int, C.POINTER(GemmEpilogue), _S]),
    "st_gemm_dt": (C.c_int, [C.c_int, C.c_int, _P, i64, _P, i64, _P, i64, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.POINTER(GemmEpilogue), _S]),
    "st_cast": (C.c_int, [_P, C.c_int, i64, _P, C.c_int, i64, i64, C.c_int, C.c_float, _S]),
    "st_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), _S]),
    "st_attn_bwd": (C.c_int, [C.POINTER(AttnBwdArgs), _S]),
    "st_mha_saved_floats": (i64, [C.c_int] * 8),
    "st_mha_ws_floats": (i64, [C.c_int] * 5),
    "st_mha_saved_floats_dt": (i64, [C.c_int] * 9),
    "st_mha_ws_floats_dt": (i64, [C.c_int] * 6),
    "st_mha_fwd": (C.c_int, [C.POINTER(MhaArgs), _S]),
    "st_mha_bwd": (C.c_int, [C.POINTER(MhaBwdArgs), _S]),
    "st_ffn_saved_floats": (i64, [i64, C.c_int, C.c_int, C.c_int]),
    "st_ffn_ws_floats": (i64, [i64, C.c_int, C.c_int]),
    "st_ffn_saved_floats_dt": (i64, [C.c_int, i64, C.c_int, C.c_int, C.c_int]),
    "st_ffn_ws_floats_dt": (i64, [C.c_int, i64, C.c_int, C.c_int]),
    "st_ffn_hidden_offset_dt": (i64, [C.c_int, i64, C.c_int, C.c_int, C.c_int]),
    "st_ffn_hidden_offset": (i64, [i64, C.c_int, C.c_int, C.c_int]),
    "st_ffn_fwd": (C.c_int, [C.POINTER(FfnArgs), _S]),
    "st_ffn_bwd": (C.c_int, [C.POINTER(FfnBwdArgs), _S]),
    "st_embed_fwd": (C.c_int, [_P, _P, _P, i64, _P, i64, C.c_int, C.c_int, C.c_int, C.c_int, _S]),
    "st_embed_bwd": (C.c_int, [_P, _P, _P, i64, C.c_int, C.c_int, i64, C.c_int, C.c_int, _S]),
    "st_decode_self_attn": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _S]),
    "st_beam_step": (C.c_int, [_P, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _S]),
    "st_frontend_saved_floats": (i64, [i64, C.c_int, C.c_int]),
    "st_frontend_ws_floats": (i64, [i64, C.c_int, C.c_int]),
    "st_frontend_hidden_offset": (i64, [i64, C.c_int, C.c_int]),
    "st_frontend_fwd": (C.c_int, [C.POINTER(FrontendArgs), _S]),
    "st_frontend_bwd": (C.c_int, [C.POINTER(FrontendBwdArgs), _S]),
    "st_linear_saved_floats": (i64, [i64, C.c_int, C.c_int, C.c_int]),
    "st_linear_ws_floats": (i64, [i64, C.c_int, C.c_int]),
    "st_linear_saved_floats_dt": (i64, [C.c_int, i64, C.c_int, C.c_int, C.c_int]),
    "st_linear_ws_floats_dt": (i64, [C.c_int, i64, C.c_int, C.c_int]),
    "st_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), _S]),
    "st_linear_bwd": (C.c_int, [C.POINTER(LinearBwdArgs), _S]),
    "st_ctc_ws_floats": (i64, [C.c_int, C.c_int, C.c_int]),
    "st_ctc_fwd_bwd": (C.c_int, [_P, i64, _P, i64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, i64, _P, i64, _S]),
    "st_ctc_grad": (C.c_int, [_P, i64, _P, i64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, i64, _P, i64, _S]),
    "st_allreduce_id_bytes": (C.c_int, []),
    "st_allreduce_unique_id": (C.c_int, [_P]),
    "st_allreduce_init": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "st_allreduce_run": (C.c_int, [_P, _P, i64, _S]),
    "st_allreduce_broadcast": (C.c_int, [_P, _P, i64, C.c_int, _S]),
    "st_allreduce_destroy": (C.c_int, [_P]),
    "st_sumsq": (C.c_int, [_P, i64, _P, _S]),
    "st_adam_step": (C.c_int, [C.POINTER(AdamArgs), _S]),
}

_lib = None


class StError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libst_b200.so and bind every declared symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StError(
            f"{LIB_PATH} not found: build it with `python speech-tranformer-pytorch_b200/build.py` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if name.startswith("st_debug_") and os.environ.get("ST_B200_LIB") and not hasattr(lib, name):
            continue             # an older build loaded for an A/B run may lack a debug hook
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    # ST_OPTIONS="name=value,name=value": library options (st_set_option, see csrc/st_host.cu) applied at load time
    for kv in filter(None, os.environ.get("ST_OPTIONS", "").split(",")):
        name, _, value = kv.partition("=")
        if lib.st_set_option(name.strip().encode(), int(value)) != 0:
            raise StError(f"ST_OPTIONS: {lib.st_last_error().decode()}")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().st_last_error()
        raise StError(f"libst_b200 error {status}: {msg.decode() if msg else '?'}")
