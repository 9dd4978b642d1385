"""SwiGLU elementwise kernels -- mirror of unsloth/kernels/swiglu.py:50-64, 112-125."""
from __future__ import annotations

import torch

from .. import _lib as L


def _fwd(act, e, g):
    L.require_cuda(e, g)
    assert e.shape == g.shape and e.dtype == g.dtype
    e_c = e if e.is_contiguous() else e.contiguous()
    g_c = g if g.is_contiguous() else g.contiguous()
    h = torch.empty(e.shape, dtype=e.dtype, device=e.device)
    L.call("ub200_glu_fwd", act, L.ptr(e_c), L.ptr(g_c), L.ptr(h), e.numel(), L.dt(e), L.stream())
    return h


def _bwd(act, DW, e, g):
    """In place: DW <- h, e <- df, g <- de (the reference's buffer-reuse contract)."""
    L.require_cuda(DW, e, g)
    assert DW.is_contiguous() and e.is_contiguous() and g.is_contiguous()
    assert DW.numel() == e.numel() == g.numel()
    L.call("ub200_glu_bwd", act, L.ptr(DW), L.ptr(e), L.ptr(g), e.numel(), L.dt(e), L.stream())
    return DW, e, g


def swiglu_fg_kernel(e, g):
    return _fwd(L.ACT_SWIGLU, e, g)


def swiglu_DWf_DW_dfg_kernel(DW, e, g):
    return _bwd(L.ACT_SWIGLU, DW, e, g)
