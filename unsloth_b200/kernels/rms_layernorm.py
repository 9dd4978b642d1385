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
