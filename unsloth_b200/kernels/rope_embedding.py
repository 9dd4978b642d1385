"""fast_rope_embedding -- host-side mirror of unsloth/kernels/rope_embedding.py:169-399.

Differences that are deliberate (B200-first), not semantic:
  * Q and K are rotated by ONE launch in both call forms (the reference's no-index form
    launches twice, rope_embedding.py:276-277) directly on the strided [B,H,S,D] views, so no
    `.contiguous()` / `.clone()` copies are ever made for last-dim-contiguous inputs;
  * no `stream.synchronize()` (the reference blocks the host per layer when more than one GPU
    is visible, rope_embedding.py:278-279).
In-place contract preserved: the returned tensors alias the inputs' storage.
"""
from __future__ import annotations

import torch

from .. import _lib as L


def _launch(Q, K, cos, sin, indices, backward, noindex_form):
    B, Hq, S, D = Q.shape
    Hk = K.shape[1] if K is not None else 0
    if noindex_form:
        comp = cos.dtype                      # math in the table dtype (rope_embedding.py:154)
    else:
        comp = torch.promote_types(Q.dtype, cos.dtype)
    L.call("ub200_rope_qk",
           L.ptr(Q), Q.stride(0), Q.stride(1), Q.stride(2),
           L.ptr(K), *( (K.stride(0), K.stride(1), K.stride(2)) if K is not None else (0, 0, 0)),
           L.ptr(cos), cos.stride(0), L.ptr(sin), sin.stride(0), L.ptr(indices),
           B, S, Hq, Hk, D, int(backward), L.dt(Q), L.dt(cos), L.dt(comp), L.stream())


def _prep_tables(cos, sin, S):
    cos, sin = cos.squeeze(), sin.squeeze()
    if cos.dim() != 2 or cos.stride(-1) != 1 or sin.stride(-1) != 1:
        cos, sin = cos.reshape(-1, cos.shape[-1]).contiguous(), sin.reshape(-1, sin.shape[-1]).contiguous()
    if cos.dtype != sin.dtype:
        sin = sin.to(cos.dtype)
    return cos, sin


def _inplace_ok(T):
    return T.stride(-1) == 1


class Fast_RoPE_Embedding_QK(torch.autograd.Function):
    """rope_embedding.py:283-399; also serves the no-index form (indices=None, noindex=True)."""

    @staticmethod
    def forward(ctx, Q, K, cos, sin, rope_indices, noindex_form=False):
        L.require_cuda(Q, K, cos, sin)
        B, Hq, S, D = Q.shape
        cos, sin = _prep_tables(cos, sin, S)
        if rope_indices is None:
            assert S <= cos.shape[0]
        Q_out = Q if _inplace_ok(Q) else Q.clone(memory_format=torch.contiguous_format)
        K_out = K if _inplace_ok(K) else K.clone(memory_format=torch.contiguous_format)
        idx = None
        if rope_indices is not None:
            idx = rope_indices.reshape(-1).to(dtype=torch.int32, device=Q.device).contiguous()
        _launch(Q_out, K_out, cos, sin, idx, False, noindex_form)
        ctx.cos, ctx.sin, ctx.idx, ctx.noindex = cos, sin, idx, noindex_form
        return Q_out, K_out

    @staticmethod
    def backward(ctx, dQ, dK):
        dQ_out = dQ if _inplace_ok(dQ) else dQ.clone(memory_format=torch.contiguous_format)
        dK_out = dK if _inplace_ok(dK) else dK.clone(memory_format=torch.contiguous_format)
        _launch(dQ_out, dK_out, ctx.cos, ctx.sin, ctx.idx, True, ctx.noindex)
        return dQ_out, dK_out, None, None, None, None


class Fast_RoPE_Embedding(torch.autograd.Function):
    """rope_embedding.py:169-261: single tensor [B, S, n_heads, D], tables [>=S, D]."""

    @staticmethod
    def forward(ctx, Q, cos, sin):
        L.require_cuda(Q, cos, sin)
        B, S, H, D = Q.shape
        cos, sin = _prep_tables(cos, sin, S)
        assert S <= cos.shape[0]
        Qv = Q if _inplace_ok(Q) else Q.contiguous()
        _launch(Qv.transpose(1, 2), None, cos, sin, None, False, True)
        ctx.cos, ctx.sin = cos, sin
        return Qv

    @staticmethod
    def backward(ctx, dY):
        dYv = dY if _inplace_ok(dY) else dY.contiguous()
        _launch(dYv.transpose(1, 2), None, ctx.cos, ctx.sin, None, True, True)
        return dYv, None, None


@torch.compiler.disable
def fast_rope_embedding(Q, K, cos, sin, rope_embedding_indices=None):
    """rope_embedding.py:265-280.  Q: [B, Hq, S, D], K: [B, Hk, S, D]; returns (Q, K)."""
    if rope_embedding_indices is not None:
        return Fast_RoPE_Embedding_QK.apply(Q, K, cos, sin, rope_embedding_indices, False)
    return Fast_RoPE_Embedding_QK.apply(Q, K, cos, sin, None, True)
