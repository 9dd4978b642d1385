"""GEGLU elementwise kernels -- mirror of unsloth/kernels/geglu.py:56-71, 126-139, 170-185,
247-260 (exact erf form and the tanh approximation used by Gemma / Gemma-2)."""
from __future__ import annotations

from .. import _lib as L
from .swiglu import _bwd, _fwd


def geglu_exact_forward_kernel(gate, up):
    return _fwd(L.ACT_GEGLU_EXACT, gate, up)


def geglu_exact_backward_kernel(DW, e, g):
    return _bwd(L.ACT_GEGLU_EXACT, DW, e, g)


def geglu_approx_forward_kernel(gate, up):
    return _fwd(L.ACT_GEGLU_APPROX, gate, up)


def geglu_approx_backward_kernel(DW, e, g):
    return _bwd(L.ACT_GEGLU_APPROX, DW, e, g)
