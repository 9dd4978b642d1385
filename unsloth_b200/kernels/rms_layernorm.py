"""fast_rms_layernorm -- host-side mirror of unsloth/kernels/rms_layernorm.py:162-286.

Same names, argument meaning and autograd contract as the reference
(`Fast_RMS_Layernorm`, `fast_rms_layernorm`, `Unsloth_LlamaRMSNorm`, `patch_rms_layernorm`);
the device work is ub200_rms_layernorm_{fwd,bwd} (csrc/rmsnorm.cu).
"""
from __future__ import annotations

import torch

from .. import _lib as L


class Fast_RMS_Layernorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X: torch.Tensor, W: torch.Tensor, eps: float, gemma: bool = False):
        L.require_cuda(X, W)
        shape = X.shape
        dim = shape[-1]
        X2 = X.reshape(-1, dim)
        if X2.stride(-1) != 1:
            X2 = X2.contiguous()
        n_rows = X2.shape[0]
        Y = torch.empty((n_rows, dim), dtype=X.dtype, device=X.device)
        r = torch.empty(n_rows, dtype=torch.float32, device=X.device)
        Wc = W if W.is_contiguous() else W.contiguous()
        L.call("ub200_rms_layernorm_fwd", L.ptr(X2), X2.stride(0), L.ptr(Wc), L.dt(Wc), L.ptr(Y),
               Y.stride(0), L.ptr(r), n_rows, dim, float(eps), int(bool(gemma)), L.dt(X), L.stream())
        ctx.eps = eps
        ctx.GEMMA = bool(gemma)
        ctx.save_for_backward(X2, Wc, r)
        return Y.view(*shape)

    @staticmethod
    def backward(ctx, dY: torch.Tensor):
        shape = dY.shape
        dim = shape[-1]
        dY2 = dY.reshape(-1, dim)
        if dY2.stride(-1) != 1:
            dY2 = dY2.contiguous()
        X, W, r = ctx.saved_tensors
        n_rows = dY2.shape[0]
        # reference contract (rms_layernorm.py:218): in place over dY unless Gemma
        dX = torch.empty_like(dY2) if ctx.GEMMA else dY2
        L.call("ub200_rms_layernorm_bwd", L.ptr(dY2), dY2.stride(0), L.ptr(X), X.stride(0), L.ptr(W),
               L.dt(W), L.ptr(r), L.ptr(dX), dX.stride(0), n_rows, dim, int(ctx.GEMMA), L.dt(dY2),
               L.stream())
        return dX.view(*shape), None, None, None


@torch.compiler.disable
def fast_rms_layernorm(layernorm, X: torch.Tensor, gemma: bool = False):
    """unsloth/kernels/rms_layernorm.py:244-255: reads `.weight` and `.variance_epsilon`/`.eps`."""
    W = layernorm.weight
    eps = layernorm.variance_epsilon if hasattr(layernorm, "variance_epsilon") else layernorm.eps
    return Fast_RMS_Layernorm.apply(X, W, eps, gemma)


class Fast_Add_RMS_Layernorm(torch.autograd.Function):
    """(S, Y) = (A + B, RMSNorm(A + B) * W) in one pass; the backward accumulates the norm gradient
    straight into the residual-stream gradient dS (in place, like the reference's in-place dY of
    rms_layernorm.py:218) and hands that one tensor to both addends."""

    @staticmethod
    def forward(ctx, A: torch.Tensor, B: torch.Tensor, W: torch.Tensor, eps: float):
        L.require_cuda(A, B, W)
        shape = A.shape
        dim = shape[-1]
        A2, B2 = A.reshape(-1, dim), B.reshape(-1, dim)
        if A2.stride(-1) != 1:
            A2 = A2.contiguous()
        if B2.stride(-1) != 1:
            B2 = B2.contiguous()
        n_rows = A2.shape[0]
        S = torch.empty((n_rows, dim), dtype=A.dtype, device=A.device)
        Y = torch.empty_like(S)
        r = torch.empty(n_rows, dtype=torch.float32, device=A.device)
        Wc = W if W.is_contiguous() else W.contiguous()
        L.call("ub200_add_rms_layernorm_fwd", L.ptr(A2), A2.stride(0), L.ptr(B2), B2.stride(0),
               L.ptr(Wc), L.ptr(S), S.stride(0), L.ptr(Y), Y.stride(0), L.ptr(r), n_rows, dim,
               float(eps), L.dt(A), L.stream())
        ctx.save_for_backward(S, Wc, r)
        ctx.set_materialize_grads(False)
        return S.view(*shape), Y.view(*shape)

    @staticmethod
    def backward(ctx, dS, dY):
        S, W, r = ctx.saved_tensors
        if dY is None:
            return dS, dS, None, None
        shape = dY.shape
        dim = shape[-1]
        dY2 = dY.reshape(-1, dim)
        if dY2.stride(-1) != 1:
            dY2 = dY2.contiguous()
        n_rows = dY2.shape[0]
        if dS is None:          # the sum itself is unused downstream (last norm of the stack)
            L.call("ub200_rms_layernorm_bwd", L.ptr(dY2), dY2.stride(0), L.ptr(S), S.stride(0),
                   L.ptr(W), L.dt(W), L.ptr(r), L.ptr(dY2), dY2.stride(0), n_rows, dim, 0,
                   L.dt(dY2), L.stream())
            g = dY2.view(*shape)
            return g, g, None, None
        dS2 = dS.reshape(-1, dim)
        if dS2.stride(-1) != 1:
            dS2 = dS2.contiguous()
        L.call("ub200_rms_layernorm_bwd_acc", L.ptr(dY2), dY2.stride(0), L.ptr(S), S.stride(0),
               L.ptr(W), L.ptr(r), L.ptr(dS2), dS2.stride(0), n_rows, dim, L.dt(dY2), L.stream())
        g = dS2.view(*shape)
        return g, g, None, None


@torch.compiler.disable
def fast_add_rms_layernorm(layernorm, residual: torch.Tensor, X: torch.Tensor):
    """`residual + X` followed by fast_rms_layernorm (models/llama.py:838-844), fused.  Returns
    (new_residual, normed).  Falls back to the two separate ops for fp32 activations, mixed
    weight dtype, a hidden size that is not a multiple of 8 or above 8192 (16-bit packed kernel,
    at most four 16-byte vectors per thread)."""
    W = layernorm.weight
    eps = layernorm.variance_epsilon if hasattr(layernorm, "variance_epsilon") else layernorm.eps
    if (X.dtype in (torch.bfloat16, torch.float16) and W.dtype == X.dtype and residual.dtype == X.dtype
            and X.shape[-1] % 8 == 0 and X.shape[-1] <= 8192):    # wider rows: the 8-vector variant spills
        return Fast_Add_RMS_Layernorm.apply(residual, X, W, eps)
    S = residual + X
    return S, Fast_RMS_Layernorm.apply(S, W, eps, False)


def _llama_rmsnorm_cls():
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
    return LlamaRMSNorm


_UNSLOTH_CLS = None


def _unsloth_rmsnorm_cls():
    """`Unsloth_LlamaRMSNorm` (rms_layernorm.py:261-264), built lazily so that importing the
    kernels does not import transformers."""
    global _UNSLOTH_CLS
    if _UNSLOTH_CLS is None:
        base = _llama_rmsnorm_cls()

        class Unsloth_LlamaRMSNorm(base):
            def forward(self, X):
                return fast_rms_layernorm(self, X, gemma=False)

        Unsloth_LlamaRMSNorm._hf_base = base
        _UNSLOTH_CLS = Unsloth_LlamaRMSNorm
    return _UNSLOTH_CLS


def __getattr__(name):
    if name == "Unsloth_LlamaRMSNorm":
        return _unsloth_rmsnorm_cls()
    raise AttributeError(name)


def patch_rms_layernorm():
    """Class-level route (rms_layernorm.py:261-274): serve a stock HF model."""
    import transformers.models.llama.modeling_llama as m
    m.LlamaRMSNorm = _unsloth_rmsnorm_cls()


def unpatch_rms_layernorm():
    import transformers.models.llama.modeling_llama as m
    m.LlamaRMSNorm = _unsloth_rmsnorm_cls()._hf_base
