"""Causal attention on the tcgen05 tensor cores -- the product the reference delegates to an external
library between `fast_rope_embedding` and `apply_o` (unsloth/utils/attention_dispatch.py:298-617:
flash-attn | xformers | SDPA; sliding window models/mistral.py:112-157; Gemma-2 soft-capping + window
models/gemma2.py:139-199).

`fast_attention(Q, K, V, softmax_scale, window, softcap, seq_info)` takes the [B, S, H, D] VIEWS of the
projection buffers (GQA native, no copies), returns [B, S, Hq, D]:
  * forward: csrc/attention.cu `ub200_attention_fwd` (TMA + tcgen05 + TMEM, exp2 online softmax,
    window / softcap / packed rows) saving the row log-sum-exp;
  * backward: csrc/attention.cu `ub200_attention_bwd` (dK/dV and dQ kernels, recompute of P from the
    saved LSE, deterministic: no atomics); D = 256 (Gemma-2) runs dV and dK as two passes because
    dK + dV + the score tiles would need 768 of the 512 TMEM columns.
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from .utils import GradModeAware

SUPPORTED_HEAD_DIMS = (64, 128, 256)


def _rows(t):
    """[B, S, H, D] view -> (row stride in elements) of the underlying [B*S, H*D] token matrix."""
    B, S, H, D = t.shape
    if t.stride(3) != 1 or t.stride(2) != D or (B > 1 and t.stride(0) != S * t.stride(1)):
        return None
    return t.stride(1)


def _as_token_rows(t):
    if _rows(t) is None or (t.data_ptr() % 16) or (t.stride(1) % 8):
        t = t.contiguous()
    return t, t.stride(1)


def attention_forward(Q, K, V, softmax_scale, window_left=-1, softcap=0.0, seq_info=None, need_lse=True):
    """Returns (O [B,S,Hq,D], lse fp32 [B,Hq,S] | [Hq, tokens] for packed rows | None)."""
    L.require_cuda(Q, K, V)
    B, S, Hq, D = Q.shape
    Hk = K.shape[2]
    if D not in SUPPORTED_HEAD_DIMS:
        raise RuntimeError("unsloth_b200.attention: head_dim %d unsupported (64 / 128 / 256)" % D)
    Q, qs = _as_token_rows(Q)
    K, ks = _as_token_rows(K)
    V, vs = _as_token_rows(V)
    O = torch.empty((B, S, Hq, D), dtype=Q.dtype, device=Q.device)
    cu, n_docs, max_len = None, 0, S
    if seq_info is not None:
        _, cu, max_len = seq_info
        n_docs = cu.numel() - 1
        if cu.dtype != torch.int32:
            cu = cu.to(torch.int32)
    lse = None
    if need_lse:
        lse = torch.empty((Hq, B * S) if cu is not None else (B, Hq, S), dtype=torch.float32, device=Q.device)
    batch, seqlen = (1, B * S) if cu is not None else (B, S)
    L.call("ub200_attention_fwd", L.ptr(Q), L.ptr(K), L.ptr(V), L.ptr(O), L.ptr(lse), L.ptr(cu), int(n_docs),
           int(max_len), int(batch), int(seqlen), Hq, Hk, D, qs, ks, vs, Hq * D, float(softmax_scale),
           int(window_left), float(softcap or 0.0), L.dt(Q), L.stream())
    return O, lse


def _bwd_library(dO, Q, K, V, O, lse, softmax_scale, window_left, softcap, seq_info):
    """A/B only (UB200_ATTN_BWD=library): flash-attn 2's backward on our forward's (O, LSE)."""
    from flash_attn.flash_attn_interface import _wrapped_flash_attn_backward, _wrapped_flash_attn_varlen_backward
    B, S, Hq, D = Q.shape
    Qc, Kc, Vc = (t.contiguous() for t in (Q, K, V))
    dQ, dK, dV = torch.empty_like(Qc), torch.empty_like(Kc), torch.empty_like(Vc)
    dOc = dO.contiguous()
    if seq_info is not None:
        _, cu, max_len = seq_info
        f = lambda t: t.reshape(B * S, t.shape[2], D)
        _wrapped_flash_attn_varlen_backward(f(dOc), f(Qc), f(Kc), f(Vc), f(O), lse, f(dQ), f(dK), f(dV), cu, cu,
                                            max_len, max_len, 0.0, softmax_scale, True, window_left, 0,
                                            float(softcap or 0.0), None, False)
    else:
        _wrapped_flash_attn_backward(dOc, Qc, Kc, Vc, O, lse, dQ, dK, dV, 0.0, softmax_scale, True, window_left, 0,
                                     float(softcap or 0.0), None, False)
    return dQ, dK, dV


def attention_backward(dO, Q, K, V, O, lse, softmax_scale, window_left=-1, softcap=0.0, seq_info=None):
    B, S, Hq, D = Q.shape
    if os.environ.get("UB200_ATTN_BWD", "own") == "library":
        return _bwd_library(dO, Q, K, V, O, lse, softmax_scale, window_left, softcap, seq_info)
    Hk = K.shape[2]
    Q, qs = _as_token_rows(Q)
    K, ks = _as_token_rows(K)
    V, vs = _as_token_rows(V)
    dO, dos = _as_token_rows(dO)
    dQ = torch.empty((B, S, Hq, D), dtype=Q.dtype, device=Q.device)
    dK = torch.empty((B, S, Hk, D), dtype=Q.dtype, device=Q.device)
    dV = torch.empty((B, S, Hk, D), dtype=Q.dtype, device=Q.device)
    delta = torch.empty_like(lse)
    cu, n_docs, max_len = None, 0, S
    if seq_info is not None:
        _, cu, max_len = seq_info
        n_docs = cu.numel() - 1
        if cu.dtype != torch.int32:
            cu = cu.to(torch.int32)
    batch, seqlen = (1, B * S) if cu is not None else (B, S)
    L.call("ub200_attention_bwd", L.ptr(dO), L.ptr(Q), L.ptr(K), L.ptr(V), L.ptr(O), L.ptr(lse), L.ptr(delta),
           L.ptr(dQ), L.ptr(dK), L.ptr(dV), L.ptr(cu), int(n_docs), int(max_len), int(batch), int(seqlen), Hq, Hk, D,
           qs, ks, vs, dos, float(softmax_scale), int(window_left), float(softcap or 0.0), L.dt(Q), L.stream())
    return dQ, dK, dV


class Fast_Attention(GradModeAware, torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, K, V, softmax_scale, window_left, softcap, seq_info):
        O, lse = attention_forward(Q, K, V, softmax_scale, window_left, softcap, seq_info)
        ctx.save_for_backward(Q, K, V, O, lse)
        ctx.args = (softmax_scale, window_left, softcap, seq_info)
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, lse = ctx.saved_tensors
        scale, window_left, softcap, seq_info = ctx.args
        dQ, dK, dV = attention_backward(dO, Q, K, V, O, lse, scale, window_left, softcap, seq_info)
        return dQ, dK, dV, None, None, None, None


def fast_attention(Q, K, V, softmax_scale=None, window=(-1, -1), softcap=0.0, seq_info=None):
    """Causal attention.  Q [B,S,Hq,D], K / V [B,S,Hk,D] (views of the projection buffers are fine).
    `window` is flash-attn's window_size pair as the reference passes it (`(sw, sw)` or `(-1, -1)`):
    under causal masking only the left extent matters.  `seq_info` = (lengths, cu_seqlens int32,
    max_seqlen) for packed rows (utils/packing.py:586-606)."""
    if softmax_scale is None:
        softmax_scale = Q.shape[-1] ** -0.5
    wl = int(window[0]) if window is not None else -1
    return Fast_Attention.apply(Q, K, V, float(softmax_scale), wl, float(softcap or 0.0), seq_info)
