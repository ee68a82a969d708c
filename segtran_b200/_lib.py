"""ctypes binding of libsegtran_b200.so (the C ABI declared in include/segtran_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make``.  There is no fallback:
if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegtran_b200.so")

SX_F32, SX_BF16 = 0, 1
SX_OP_TF32, SX_OP_BF16 = 0, 1
SX_MAJOR_K, SX_MAJOR_MN = 0, 1
SX_BIAS_NONE, SX_BIAS_N, SX_BIAS_M = 0, 1, 2
SX_ACT_NONE, SX_ACT_GELU, SX_ACT_GELU_BWD = 0, 1, 2
SX_SCHED_WARMUP_LINEAR, SX_SCHED_WARMUP_CONSTANT = 0, 1


class SxError(RuntimeError):
    pass


class sx_operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("major", C.c_int32), ("_pad", C.c_int32), ("ld", C.c_int64),
                ("stride_z0", C.c_int64), ("stride_z1", C.c_int64)]


class sx_gemm_args(C.Structure):
    _fields_ = [("op_dtype", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("Z0", C.c_int32),
                ("Z1", C.c_int32), ("A", sx_operand), ("B", sx_operand), ("C", C.c_void_p), ("c_dtype", C.c_int32),
                ("round_tf32", C.c_int32), ("ldc", C.c_int64), ("c_stride_z0", C.c_int64), ("c_stride_z1", C.c_int64),
                ("alpha", C.c_float), ("bias_mode", C.c_int32), ("bias", C.c_void_p), ("bias_stride_z0", C.c_int64),
                ("bias_stride_z1", C.c_int64), ("act", C.c_int32), ("accumulate", C.c_int32), ("preact", C.c_void_p),
                ("split_k", C.c_int32), ("_pad2", C.c_int32), ("amax", C.c_void_p), ("drop_p", C.c_float),
                ("_pad3", C.c_uint32), ("drop_seed", C.c_uint64), ("drop_seed_dev", C.c_void_p), ("addend", C.c_void_p), ("colsum", C.c_void_p)]


class sx_attn_probs_args(C.Structure):
    _fields_ = [("B", C.c_int32), ("M", C.c_int32), ("U1", C.c_int32), ("U2", C.c_int32), ("d", C.c_int32),
                ("round_tf32", C.c_int32), ("Q", C.c_void_p), ("q_ld", C.c_int64), ("q_bstride", C.c_int64),
                ("K", C.c_void_p), ("k_ld", C.c_int64), ("k_bstride", C.c_int64), ("alpha", C.c_float),
                ("clip", C.c_float), ("P", C.c_void_p), ("S", C.c_void_p), ("ldp", C.c_int64), ("lse", C.c_void_p),
                ("rowmax", C.c_void_p), ("stat", C.c_void_p), ("diag", C.c_void_p), ("drop_p", C.c_float),
                ("_pad", C.c_uint32), ("drop_seed", C.c_uint64), ("drop_seed_dev", C.c_void_p),
                ("scratch", C.c_void_p), ("scratch_floats", C.c_int64)]


_P, _I, _L, _F, _U64, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_double

# name -> argtypes (every function returns int; 0 = success)
_PROTOS = {
    "sx_gemm": [C.POINTER(sx_gemm_args), _P],
    "sx_gemm_debug_set": [C.c_char_p, _L],
    "sx_attn_probs_fwd": [C.POINTER(sx_attn_probs_args), _P],
    "sx_reduce_max": [_P, _L, _P, _P],
    "sx_pos_lsinu_fwd": [_P, _P, _L, _I, _P, _P, _I, _P, _P],
    "sx_pos_lsinu_bwd": [_P, _P, _L, _I, _P, _P, _I, _P, _P, _P, _P, _P],
    "sx_prologue_fwd": [_P, _L, _I, _I, _P, _P, _P, _I, _L, _F, _P, _F, _U64, _P, _P, _I, _I, _P, _P],
    "sx_prologue_bwd": [_P, _P, _L, _I, _I, _P, _P, _P, _I, _L, _F, _P, _F, _U64, _P, _P, _P, _P, _P, _P, _P, _P],
    "sx_softmax_fwd": [_P, _L, _I, _L, _P, _F, _F, _U64, _P, _P, _I, _L, _I, _P, _P, _P],
    "sx_softmax_bwd": [_P, _L, _P, _L, _P, _L, _I, _P, _F, _F, _U64, _P, _L, _P, _I, _L, _I, _P],
    "sx_layernorm_fwd": [_P, _L, _I, _P, _P, _P, _I, _I, _P, _P],
    "sx_layernorm_bwd": [_P, _P, _L, _I, _P, _P, _P, _I, _I, _P, _P, _P],
    "sx_ln_softaggr_fwd": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _F, _U64, _P, _P, _P, _P, _P],
    "sx_ln_softaggr_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _U64, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P],
    "sx_softaggr_fwd": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "sx_softaggr_bwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "sx_gelu_bwd": [_P, _P, _I, _L, _F, _U64, _P, _P, _I, _I, _P],
    "sx_seed_derive": [_P, _U64, _P, _P],
    "sx_seed_advance": [_P, _U64, _P],
    "sx_convert": [_P, _I, _L, _P, _I, _I, _P],
    "sx_sw_accumulate": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "sx_sw_finalize": [_P, _P, _I, _L, _I, _P, _P],
    "sx_split_tf32": [_P, _L, _P, _P, _P],
    "sx_split_tf32_cat": [_P, _I, _I, _I, _I, _L, _L, _L, _L, _I, _I, _P, _P],
    "sx_colsum": [_P, _I, _L, _I, _L, _P, _P],
    "sx_transpose": [_P, _L, _I, _I, _P, _P],
    "sx_colsum_batched": [_P, _I, _L, _I, _L, _L, _I, _L, _P, _P],
    "sx_dot": [_P, _P, _L, _P, _P],
    "sx_add": [_P, _P, _L, _P, _P],
    "sx_rowsum": [_P, _L, _L, _L, _I, _P, _P],
    "sx_scale": [_P, _L, _P, _F, _P, _P],
    "sx_head_contract_fwd": [_P, _P, _P, _I, _I, _L, _I, _P, _I, _P],
    "sx_head_contract_bwd_data": [_P, _P, _I, _I, _L, _I, _P, _P],
    "sx_head_contract_bwd_weight": [_P, _P, _I, _I, _L, _I, _P, _P],
    "sx_token_scores": [_P, _P, _I, _I, _I, _I, _P, _P],
    "sx_token_scores_bwd": [_P, _P, _I, _I, _I, _I, _P, _P],
    "sx_resize_axis_fwd": [_P, _L, _I, _I, _L, _P, _I, _P],
    "sx_resize_axis_bwd": [_P, _L, _I, _I, _L, _P, _P],
    "sx_groupnorm_fwd": [_P, _I, _I, _L, _I, _P, _P, _F, _P, _P, _P, _I, _P],
    "sx_groupnorm_bwd": [_P, _P, _I, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "sx_seg_loss_fwd": [_P, _P, _I, _I, _L, _P, _P, _F, _P, _P, _P, _P],
    "sx_seg_loss_bwd": [_P, _P, _I, _I, _L, _P, _P, _F, _P, _P, _P],
    "sx_adam_step": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _D, _D, _D, _F, _F, _F, _L, _I, _P, _P, _P, _P, _P, _P],
    "sx_sgemm_small": [_P, _P, _P, _I, _I, _I, _L, _L, _L, _L, _L, _L, _I, _L, _L, _L, _F, _I, _P],
}
# every symbol include/segtran_b200.h declares (checked by tests/test_abi.py)
EXPORTS = sorted(list(_PROTOS) + ["sx_version", "sx_last_error", "sx_device_info"])

_lib = None


def lib():
    """The loaded CDLL; raises if the extension has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SxError("segtran_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make` at the repo root (there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.sx_version.restype = C.c_int
        l.sx_last_error.restype = C.c_char_p
        l.sx_device_info.argtypes = [C.POINTER(C.c_int)] * 3
        l.sx_device_info.restype = C.c_int
        for name, at in _PROTOS.items():
            f = getattr(l, name)
            f.argtypes = at
            f.restype = C.c_int
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise SxError("%s failed (rc=%d): %s" % (what, rc, lib().sx_last_error().decode("utf-8", "replace")))


# kernels launched per C-ABI call (for bench.py's gpu_launches claim); default 1
_LAUNCHES = {"sx_pos_lsinu_bwd": 2, "sx_ln_softaggr_bwd": 1, "sx_prologue_bwd": 2, "sx_layernorm_bwd": 2, "sx_gemm_debug_set": 0,
             "sx_attn_probs_fwd": 3}
launch_count = 0
_hook = None          # optional callable(name, args) -> context manager, installed by bench.py for per-kernel timing


def set_hook(h):
    global _hook
    _hook = h


def call(name, *args):
    global launch_count
    launch_count += _LAUNCHES.get(name, 1)
    if _hook is None:
        check(getattr(lib(), name)(*args), name)
    else:
        with _hook(name, args):
            check(getattr(lib(), name)(*args), name)
