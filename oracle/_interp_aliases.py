"""TEST INFRASTRUCTURE ONLY.  Interpreter-safe replacements for the two import-time
aliases in the reference's kernels/utils.py:61-74 (`triton_cast`, `triton_tanh`), which
capture builtins the TRITON_INTERPRET=1 interpreter cannot call."""
import triton
import triton.language as tl


@triton.jit
def interp_cast(x, dtype):
    return tl.full((), 1, dtype) * x


@triton.jit
def interp_tanh(x):
    return 2 * tl.sigmoid(2 * x) - 1
