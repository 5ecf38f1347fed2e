"""ctypes binding of the C ABI in include/scat_b200.h.

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised with the library's own error string.  Build it with ``python -m scat_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libscat_b200.so")

PREC = {"fp32": 0, "tf32": 1, "bf16": 2, "tf32x3": 3}
X2_DTYPE = {"fp32": 0, "bf16": 1}       # SCAT_DTYPE_*: storage of the seam tensors x2 / x2.grad
ABI_VERSION = 2
EPI = {"none": 0, "bias": 1, "bias_resid": 2, "bias_gelu": 3, "dgelu": 4, "resid": 5}
NUM_PARAMS = 35


class ScatHeadDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "n_tokens", "channels", "token_dim", "heads", "iteration", "pos_embed", "n_masked",
        "pl_reg", "precision", "main_feat_dim", "n_out", "x2_dtype")]


_f = C.c_void_p      # device pointers travel as integers
_i32 = C.c_int32
_i64 = C.c_int64
_sz = C.c_size_t
_fl = C.c_float
_pp = C.POINTER(C.c_void_p)
_desc = C.POINTER(ScatHeadDesc)

# name -> (restype, argtypes); must list every symbol declared in include/scat_b200.h
SIGNATURES = {
    "scat_abi_version": (_i32, []),
    "scat_last_error_string": (C.c_char_p, []),
    "scat_launch_count": (C.c_uint64, []),
    "scat_head_workspace_bytes": (_sz, [_desc]),
    "scat_head_forward": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _f, _f, _f, _sz, _f]),
    "scat_head_backward": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _pp, _f, _f, _f, _sz, _f]),
    "scat_proj_loss": (_i32, [_i32, _i32, _i32, _f, _f, _i32, _f, _fl, _fl, _fl, _f, _f, _f, _f]),
    "scat_head_train_step": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _i32, _fl, _fl, _fl, _f, _f, _f, _f, _pp,
                                    _f, _f, _f, _sz, _f]),
    "scat_head_train_step_phase": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _i32, _fl, _fl, _fl, _f, _f, _f, _f, _pp,
                                          _f, _f, _f, _sz, _f, _i32]),
    "scat_head_train_step_hooked": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _i32, _fl, _fl, _fl, _f, _f, _f, _f, _pp,
                                           _f, _f, _f, _sz, C.c_void_p, C.c_void_p, _f]),
    "scat_adam_step": (_i32, [_f, _f, _f, _f, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                             _i32, _f, _f, _f, _f]),
    "scat_eval_procrustes": (_i32, [_f, _f, _i32, _i32, _f, _f, _f]),
    "scat_eval_joint_errors": (_i32, [_f, _f, _i32, _i32, _fl, C.POINTER(C.c_double), _i32, _f, _f, _f]),
    "scat_eval_accel": (_i32, [_f, _f, _i32, _i32, _f, _f]),
    "scat_peer_signal_bytes": (_sz, []),
    "scat_peer_alloc": (_i32, [_sz, _pp]),
    "scat_peer_free": (_i32, [_f]),
    "scat_peer_export": (_i32, [_f, _f]),
    "scat_peer_open": (_i32, [_f, _pp]),
    "scat_peer_close": (_i32, [_f]),
    "scat_peer_allreduce": (_i32, [_pp, _pp, _i32, _i32, C.c_longlong, C.c_longlong, _f]),
    "scat_peer_allreduce_part": (_i32, [_pp, _pp, _i32, _i32, C.c_longlong, C.c_longlong, _i32, _f]),
    "scat_peer_error": (_i32, [_f, C.POINTER(C.c_int32)]),
    "scat_peer_error_word": (C.c_void_p, [_f]),
    "scat_tokens_forward": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _sz, _f]),
    "scat_coarse_forward": (_i32, [_desc, _pp, _f, _f, _f, _f, _f, _f, _f, _f, _f, _sz, _f]),
    "scat_gemm": (_i32, [_f, _i64, _i64, _f, _i64, _i64, _f, _i32, _i32, _i32, _i32, _i32, _f, _f, _i32, _f, _i32,
                         _i32, _i32, _f]),
    "scat_gemm_bf16": (_i32, [_f, _i64, _i64, _f, _i64, _i64, _f, _i32, _f, _i32, _i32, _i32, _i32, _i32, _f, _f, _i32,
                              _f, _i32, _i32, _f]),
    "scat_debug_gemm_timeline": (None, [_f]),
    "scat_conv_pe_mask_fwd": (_i32, [_f, _f, _f, _f, _f, _i32, _i32, _f, _f, _i32, _i32, _i32, _i32, _f]),
    "scat_conv_tc_scratch_floats": (_sz, [_i32, _i32, _i32, _i32]),
    "scat_conv_pe_mask_fwd_tc": (_i32, [_f, _i32, _f, _f, _f, _f, _i32, _i32, _f, _f, _f, _i32, _i32, _i32, _i32, _f]),
    "scat_conv_bwd_tc": (_i32, [_f, _f, _i32, _f, _f, _i32, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _f]),
    "scat_conv_bwd_scratch_floats": (_sz, [_i32, _i32, _i32, _i32]),
    "scat_conv_bwd": (_i32, [_f, _f, _f, _f, _i32, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _f]),
    "scat_layernorm_fwd": (_i32, [_f, _f, _f, _f, _f, _f, _i32, _i32, _f]),
    "scat_layernorm_bwd": (_i32, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i32, _i32, _f]),
    "scat_attention_fwd": (_i32, [_f, _f, _f, _i32, _i32, _i32, _f]),
    "scat_attention_bwd": (_i32, [_f, _f, _f, _f, _i32, _i32, _i32, _f]),
    "scat_attention_fwd_tc": (_i32, [_f, _f, _f, _i32, _i32, _i32, _f]),
    "scat_attention_bwd_tc": (_i32, [_f, _f, _f, _f, _i32, _i32, _i32, _f]),
    "scat_regressor_fwd": (_i32, [_f, _f, _f, _f, _f, _f, _f, _f, _i32, _i32, _i32, _i32, _i32, _f]),
    "scat_lbs_derived_floats": (_sz, []),
    "scat_lbs_prepare": (_i32, [_f, _f, _f, _f, _f, _f, _f]),
    "scat_lbs_fwd": (_i32, [_f, _f, _f, _f, _f, _f, _i32, _f]),
    "scat_lbs_tc_table_floats": (_sz, []),
    "scat_lbs_tc_scratch_floats": (_sz, [_i32]),
    "scat_lbs_tc_prepare": (_i32, [_f, _f, _f, _f]),
    "scat_lbs_fwd_tc": (_i32, [_f, _f, _f, _f, _f, _f, _f, _i32, _f, _sz, _f]),
    "scat_lbs_bwd": (_i32, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i32, _f]),
}

_lib = None


# int (*scat_grads_ready_fn)(void* user, int32_t part, void* side_stream)  (include/scat_b200.h)
GRADS_READY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p)


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"scat_b200: {LIB_PATH} is missing. The CUDA extension is required (there is no CPU/PyTorch "
            f"fallback); build it with `python -m scat_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.scat_abi_version() != ABI_VERSION:
        raise RuntimeError("scat_b200: ABI version mismatch between _lib.py and libscat_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().scat_last_error_string().decode(errors="replace")
        raise RuntimeError(f"scat_b200: {what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
