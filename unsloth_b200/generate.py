"""KV-cache decoding around `fast_linear_forward`: host-side mirror of the
reference's single-token inference path -- `LlamaAttention_fast_forward_inference`
(unsloth/models/llama.py:352-560), `fast_swiglu_inference` (:566-602), `fast_rms_layernorm_inference`
(:608-640) and the decode loop of `LlamaModel_fast_forward_inference` (:1233-1367).

Per generated token and layer: RMSNorm (torch ops, as the reference), q/k/v through
`fast_linear_forward` = ONE NF4 GEMV launch each with the LoRA term in its epilogue
(`ub200_gemv_nf4`), rotate-half RoPE on the single position, the new K/V row appended to a
pre-allocated cache that grows in steps of KV_CACHE_INCREMENT, attention of the one query row
against the cache (`(Q * scale) @ K^T` -> fp32 softmax -> `@ V`, the reference's bsz == 1 form),
o_proj / gate / up / down through `fast_linear_forward`, and `lm_head` as a dense GEMV.
The prompt is prefilled through the training-path forward (patch.Model_fast_forward) under no_grad.
"""
from __future__ import annotations

import math

import torch

from .kernels.utils import fast_linear_forward

KV_CACHE_INCREMENT = 512   # models/llama.py: KV_CACHE_INCREMENT


def _rms_inference(norm, X, gemma=False):
    """fast_rms_layernorm_inference (llama.py:608-640; Gemma variant gemma.py:80-95)."""
    eps = getattr(norm, "variance_epsilon", None)
    if eps is None:
        eps = norm.eps
    Xf = X.float()
    var = Xf.square().mean(-1, keepdim=True)
    Xf = Xf * torch.rsqrt(var + eps)
    if gemma:
        return (Xf * (1.0 + norm.weight.float())).to(X.dtype)
    return Xf.to(X.dtype) * norm.weight


class _LayerCache:
    def __init__(self, K, V):
        """K, V: [B, Hk, S, D] from the prefill."""
        B, Hk, S, D = K.shape
        self.len = S
        self.buf = torch.empty((KV_CACHE_INCREMENT + S + 1, 2, B, Hk, D), dtype=K.dtype, device=K.device)
        self.buf[:S, 0] = K.permute(2, 0, 1, 3)
        self.buf[:S, 1] = V.permute(2, 0, 1, 3)

    def append(self, Kn, Vn):
        if self.len + 1 >= self.buf.shape[0]:
            new = torch.empty((self.buf.shape[0] + KV_CACHE_INCREMENT,) + tuple(self.buf.shape[1:]),
                              dtype=self.buf.dtype, device=self.buf.device)
            new[:self.len] = self.buf[:self.len]
            self.buf = new
        self.buf[self.len, 0] = Kn.permute(2, 0, 1, 3)[0]
        self.buf[self.len, 1] = Vn.permute(2, 0, 1, 3)[0]
        self.len += 1
        K = self.buf[:self.len, 0].permute(1, 2, 0, 3)
        V = self.buf[:self.len, 1].permute(1, 2, 0, 3)
        return K, V


def _attention_decode(attn, X, cache, cos, sin, pos):
    """llama.py:446-560 for one new token.  X [B, 1, H]."""
    bsz = X.shape[0]
    n_heads, n_kv, hd = attn._ub_heads
    Qn = fast_linear_forward(attn.q_proj, X).view(bsz, 1, n_heads, hd).transpose(1, 2)
    Kn = fast_linear_forward(attn.k_proj, X).view(bsz, 1, n_kv, hd).transpose(1, 2)
    Vn = fast_linear_forward(attn.v_proj, X).view(bsz, 1, n_kv, hd).transpose(1, 2)
    c = cos[pos].to(Qn.dtype).view(1, 1, 1, hd)
    s = sin[pos].to(Qn.dtype).view(1, 1, 1, hd)
    h = hd // 2

    def rot(t):
        rh = torch.cat((-t[..., h:], t[..., :h]), -1)
        return t * c + rh * s
    Qn, Kn = rot(Qn), rot(Kn)
    K, V = cache.append(Kn, Vn)
    sw = attn._ub_window
    if sw is not None and K.shape[2] > sw + 1:          # the training window keeps sw + 1 keys (mistral.py:112-128)
        K, V = K[:, :, -(sw + 1):], V[:, :, -(sw + 1):]
    g = n_heads // n_kv
    if g > 1:
        K = K[:, :, None].expand(bsz, n_kv, g, K.shape[2], hd).reshape(bsz, n_heads, -1, hd)
        V = V[:, :, None].expand(bsz, n_kv, g, V.shape[2], hd).reshape(bsz, n_heads, -1, hd)
    A = torch.matmul(Qn * attn._ub_scale, K.transpose(2, 3))          # (Q * scalar) @ K^T  (llama.py:515-519)
    if attn._ub_softcap:
        A = attn._ub_softcap * torch.tanh(A / attn._ub_softcap)
    A = torch.softmax(A, dim=-1, dtype=torch.float32).to(Qn.dtype)
    A = torch.matmul(A, V).transpose(1, 2).reshape(bsz, 1, n_heads * hd)
    return fast_linear_forward(attn.o_proj, A)


def _mlp_decode(mlp, X, gemma):
    """fast_swiglu_inference / fast_geglu_inference (llama.py:566-602, gemma.py:50-75)."""
    gate = fast_linear_forward(mlp.gate_proj, X)
    up = fast_linear_forward(mlp.up_proj, X)
    act = torch.nn.functional.gelu(gate, approximate="tanh") if gemma else torch.nn.functional.silu(gate)
    return fast_linear_forward(mlp.down_proj, act * up)


@torch.inference_mode()
def generate(model, input_ids, max_new_tokens=16):
    """Greedy decoding.  input_ids [B, S] on the model's device.  Returns [B, S + max_new_tokens].
    The q_len == 1 GEMV path of `fast_linear_forward` is taken for B == 1."""
    from . import patch as P
    inner = model.model
    gemma = inner._ub_gemma
    sinks = []
    for layer in inner.layers:
        layer.self_attn._ub_kv_sink = []
        sinks.append(layer.self_attn._ub_kv_sink)
    try:
        hidden = P.Model_fast_forward(inner, input_ids)               # prefill (training-path kernels)
    finally:
        for layer in inner.layers:
            layer.self_attn._ub_kv_sink = None
    caches = [_LayerCache(k, v) for (k, v), in sinks]
    logits = torch.nn.functional.linear(hidden[:, -1:], model.lm_head.weight)
    out = [input_ids]
    pos = input_ids.shape[1]
    softcap = model._ub_final_softcap
    for step in range(max_new_tokens):
        if softcap:
            logits = softcap * torch.tanh(logits.float() / softcap)
        nxt = logits[:, -1].argmax(-1, keepdim=True)
        out.append(nxt)
        if step == max_new_tokens - 1:
            break
        cos, sin = inner._ub_rotary.get(pos + 1)
        X = inner.embed_tokens(nxt)
        if gemma and not hasattr(inner.embed_tokens, "embed_scale"):
            X = X * torch.tensor(math.sqrt(inner.config.hidden_size), dtype=X.dtype, device=X.device)
        for layer, cache in zip(inner.layers, caches):
            res = X
            h = _rms_inference(layer.input_layernorm, X, gemma)
            h = _attention_decode(layer.self_attn, h, cache, cos, sin, pos)
            if gemma:
                h = _rms_inference(layer.post_attention_layernorm, h, True)
            X = res + h
            res = X
            if gemma:
                h = _rms_inference(layer.pre_feedforward_layernorm, X, True)
                h = _mlp_decode(layer.mlp, h, True)
                h = _rms_inference(layer.post_feedforward_layernorm, h, True)
            else:
                h = _rms_inference(layer.post_attention_layernorm, X)
                h = _mlp_decode(layer.mlp, h, False)
            X = res + h
        X = _rms_inference(inner.norm, X, gemma)
        logits = fast_linear_forward(model.lm_head, X) if hasattr(model.lm_head, "base_layer") else \
            torch.nn.functional.linear(X, model.lm_head.weight)
        pos += 1
    return torch.cat(out, dim=1)
