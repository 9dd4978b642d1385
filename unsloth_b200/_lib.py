"""ctypes binding of libunsloth_b200.so (the C ABI declared in include/unsloth_b200.h).

There is NO fallback: if the shared library is missing or does not export a symbol the
import raises.  Every call passes raw device pointers, sizes, strides and the CURRENT CUDA
stream (never a stream cached at import -- compare unsloth/kernels/utils.py:241-258 in the
reference, which captures the default stream once).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UB200_LIB_PATH") or os.path.join(_HERE, "_C", "libunsloth_b200.so")   # override: A/B builds

F32, F16, BF16 = 0, 1, 2
ACT_SWIGLU, ACT_GEGLU_APPROX, ACT_GEGLU_EXACT = 0, 1, 2
GEMM_MAX_SEGMENTS = 8
GLU_EPI_FWD, GLU_EPI_BWD = 1, 2

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

ERRORS = {-1: "bad argument (shape / stride / alignment)", -2: "unsupported configuration",
          -3: "cuTensorMapEncodeTiled unavailable (no CUDA driver)",
          -4: "tensor-map encoding rejected the operand"}


class GemmSegment(Structure):
    _fields_ = [("a", c_void_p), ("lda", c_int64), ("b", c_void_p), ("ldb", c_int64),
                ("k", c_int64)]


class GemmProblem(Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("segs", POINTER(GemmSegment)), ("n_segs", c_int),
                ("a_mn_major", c_int), ("b_mn_major", c_int), ("C", c_void_p), ("ldc", c_int64),
                ("c_dtype", c_int), ("alpha", c_float), ("accumulate", c_int), ("split_k", c_int),
                ("workspace", c_void_p), ("block_n", c_int), ("signals", c_int),
                ("wait_problem", c_int), ("wait_segment", c_int), ("wait_all", c_int)]


GROUPED_MAX_PROBLEMS, GROUPED_MAX_SEGMENTS = 8, 4


class CastDesc(Structure):          # ub200_cast_desc
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("src_ld", c_int64), ("dst_ld", c_int64),
                ("src_dtype", c_int), ("dst_dtype", c_int), ("rows", c_int), ("cols", c_int),
                ("dst_rows", c_int), ("dst_cols", c_int), ("row_off", c_int), ("col_off", c_int),
                ("scale", c_float), ("transpose", c_int)]


class AccDesc(Structure):           # ub200_acc_desc
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("src_rs", c_int64), ("src_cs", c_int64),
                ("rows", c_int), ("cols", c_int)]


def _load():
    if not os.path.exists(LIB_PATH):
        # build in-tree when a compiler is around (developer box); never fall back to CPU code
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # pragma: no cover
            raise RuntimeError(
                "unsloth_b200: %s is missing and could not be built (%s). The CUDA library is "
                "mandatory -- there is no CPU / eager fallback." % (LIB_PATH, e)) from e
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_i, _l, _f, _p = c_int, c_int64, c_float, c_void_p
_SIGS = {
    "ub200_abi_version": ([], c_int),
    "ub200_rms_layernorm_fwd": ([_p, _l, _p, _i, _p, _l, _p, _l, _i, _f, _i, _i, _p], c_int),
    "ub200_rms_layernorm_bwd": ([_p, _l, _p, _l, _p, _i, _p, _p, _l, _l, _i, _i, _i, _p], c_int),
    "ub200_add_rms_layernorm_fwd": ([_p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _l, _i, _f, _i, _p], c_int),
    "ub200_rms_layernorm_bwd_acc": ([_p, _l, _p, _l, _p, _p, _p, _l, _l, _i, _i, _p], c_int),
    "ub200_rope_qk": ([_p, _l, _l, _l, _p, _l, _l, _l, _p, _l, _p, _l, _p, _i, _i, _i, _i, _i, _i,
                       _i, _i, _i, _p], c_int),
    "ub200_glu_fwd": ([_i, _p, _p, _p, _l, _i, _p], c_int),
    "ub200_glu_bwd": ([_i, _p, _p, _p, _l, _i, _p], c_int),
    "ub200_cross_entropy_fwd": ([_p, _l, _p, _p, _p, _l, _i, _f, _f, _i, _p], c_int),
    "ub200_cross_entropy_bwd": ([_p, _l, _p, _p, _p, _l, _l, _i, _f, _f, _i, _p], c_int),
    "ub200_dequantize_nf4": ([_p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p], c_int),
    "ub200_quantize_nf4": ([_p, _i, _p, _p, _l, _i, _p], c_int),
    "cdequantize_blockwise_fp32": ([_p, _p, _p, _p, _i, _i, _p], None),
    "cdequantize_blockwise_bf16_nf4": ([_p, _p, _p, _p, _i, _i, _p], None),
    "cdequantize_blockwise_fp16_nf4": ([_p, _p, _p, _p, _i, _i, _p], None),
    "ub200_gemv_nf4": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _f, _i, _p], c_int),
    "ub200_gemv_dense": ([_p, _p, _l, _p, _i, _i, _i, _i, _p], c_int),
    "cgemm_4bit_inference_naive_bf16": ([_i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p], None),
    "cgemm_4bit_inference_naive_fp16": ([_i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p], None),
    "ub200_gemm": ([_i, _i, POINTER(GemmSegment), _i, _i, _i, _i, _p, _l, _i, _f, _i, _i, _p, _i,
                    _i, _p], c_int),
    "ub200_gemm_workspace_bytes": ([_i, _i, _i, POINTER(c_int64)], c_int),
    "ub200_gemm_glu": ([_i, _i, _i, _i, POINTER(GemmSegment), _i, _i, _i, _i, _p, _l, _p, _p, _l, _f, _i, _i,
                        _p], c_int),
    "ub200_gemm_grouped": ([POINTER(GemmProblem), _i, _i, _p, _p], c_int),
    "ub200_gemm_grouped_scratch_ints": ([POINTER(GemmProblem), _i, POINTER(c_int)], c_int),
    "ub200_gemm_nf4": ([_i, _i, _i, _p, _l, _p, _p, _p, _p, _p, _i, _i, _p, _l, _p, _l, _i, _p, _l, _i, _p], c_int),
    "ub200_attention_fwd": ([_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _l, _l, _l, _l, _f, _i, _f, _i, _p], c_int),
    "ub200_attention_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _l, _l, _l, _l,
                             _f, _i, _f, _i, _p], c_int),
    "ub200_cast_pad_2d": ([_p, _i, _l, _i, _i, _p, _i, _l, _i, _i, _i, _i, _f, _i, _p], c_int),
    "ub200_adamw_flat": ([_p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _f, _f, _p], c_int),
    "ub200_cast_pad_multi": ([POINTER(CastDesc), _i, _p], c_int),
    "ub200_accumulate_multi": ([POINTER(AccDesc), _i, _p], c_int),
}
EXPORTED_SYMBOLS = tuple(_SIGS)
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == ABI mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _res

# number of kernel-launching C-ABI calls made through this module (bench.py: gpu_launches)
launch_count = 0


def dt(t) -> int:
    d = t if isinstance(t, torch.dtype) else t.dtype
    try:
        return _DT[d]
    except KeyError:
        raise TypeError("unsloth_b200: unsupported dtype %s" % d) from None


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "unsloth_b200: got a %s tensor; the hot path runs only on CUDA (sm_100a) -- "
                "there is no CPU fallback" % t.device)


def call(name, *args):
    """Invoke an int-returning entry point, raise on a non-zero status."""
    global launch_count
    launch_count += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        if rc > 0:
            msg = "CUDA error %d" % rc
        else:
            msg = ERRORS.get(rc, "error %d" % rc)
        raise RuntimeError("unsloth_b200.%s failed: %s" % (name, msg))
