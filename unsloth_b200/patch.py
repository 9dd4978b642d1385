"""The monkey-patch surface: host-side mirror of the patched model forward of the reference
(`unsloth/models/llama.py`: LlamaAttention_fast_forward :671-771, LlamaDecoderLayer_fast_forward
:774-863, LlamaModel_fast_forward :866-1230, CausalLM_fast_forward :1370-1590; patch_peft_model
:3599-3770; per-arch deltas `mistral.py:112-157`, `gemma2.py:104-283`).

`install(model)` rebinds, per instance, exactly what Unsloth rebinds:
    layer.mlp.forward          = MethodType(apply_lora_mlp_swiglu | apply_lora_mlp_geglu_approx, mlp)
    layer.self_attn.apply_qkv  = apply_lora_qkv
    layer.self_attn.apply_o    = apply_lora_o
and replaces the forward of the attention / decoder-layer / model / CausalLM modules of a STOCK
HuggingFace Llama / Mistral / Gemma-2 model with slim fast forwards that call the kernel API by
name (fast_rms_layernorm, fast_rope_embedding, unsloth_fused_ce_loss).  The attention product
itself is the external flash-attn library call, as in the reference
(`unsloth/utils/attention_dispatch.py:433-617`).

`build_qlora_model(...)` creates a random-init model of a named architecture directly on the GPU,
NF4-quantises the seven projections of every layer (bitsandbytes-format statistics) and wraps
them with LoRA -- what `FastLanguageModel.from_pretrained(load_in_4bit=True)` +
`get_peft_model` produce, minus the hub.
"""
from __future__ import annotations

import math
import os
import types
from types import SimpleNamespace

import torch

from . import kernels as K
from .lora import LoraLinear
from .nf4 import Linear4bit
from .packing import (get_packed_info_from_kwargs, mask_packed_boundary_labels,
                      mask_packed_sequence_boundaries, packed_position_ids)

from .model_configs import CONFIGS, hf_config  # noqa: E402,F401  (re-exported)


# ---------------------------------------------------------------------------------------------
# rotary tables (models/llama.py:1775-1914 LlamaRotaryEmbedding; gemma.py:247-319 fp32 tables)
# ---------------------------------------------------------------------------------------------
class RotaryCache:
    """cos/sin tables [max_pos, D] (both halves duplicated), fp32 inv_freq, llama3 scaling; table
    dtype = model dtype (Llama / Mistral) or fp32 (Gemma family), grown in steps of 8192."""

    def __init__(self, head_dim, base, device, dtype, rope_scaling=None):
        self.dim, self.base, self.device, self.dtype = head_dim, base, device, dtype
        self.rope_scaling = rope_scaling
        self.size = 0
        self.cos = self.sin = None

    def inv_freq(self):
        inv = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.int64).float() / self.dim))
        rs = self.rope_scaling
        kind = (rs.get("rope_type", rs.get("type")) if rs else None) or "default"
        if kind == "linear":
            # LlamaLinearScalingRotaryEmbedding (models/llama.py:1917-1960): positions / factor
            inv = inv / float(rs["factor"])
        elif kind not in ("default", "llama3"):
            raise NotImplementedError(
                "unsloth_b200: rope scaling %r is not implemented (default, linear and llama3 are); "
                "refusing to train with unscaled frequencies" % kind)
        if kind == "llama3":
            factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
            old = rs["original_max_position_embeddings"]
            low_wl, high_wl = old / lo, old / hi
            wl = 2 * math.pi / inv
            scaled = torch.where(wl > low_wl, inv / factor, inv)
            smooth = (old / wl - lo) / (hi - lo)
            smoothed = (1 - smooth) / factor * inv + smooth * inv
            mid = ~(wl < high_wl) * ~(wl > low_wl)
            inv = torch.where(mid, smoothed, scaled)
        return inv

    def get(self, seq_len):
        if seq_len > self.size:
            self.size = ((seq_len + 8191) // 8192) * 8192
            t = torch.arange(self.size, dtype=torch.float32)
            freqs = torch.outer(t, self.inv_freq())
            emb = torch.cat((freqs, freqs), dim=-1)
            self.cos = emb.cos().to(device=self.device, dtype=self.dtype)
            self.sin = emb.sin().to(device=self.device, dtype=self.dtype)
        return self.cos, self.sin


# ---------------------------------------------------------------------------------------------
# fast forwards
# ---------------------------------------------------------------------------------------------
_SDPA_CUDNN = {"ok": None}


def _attention_policy():
    """UB200_ATTENTION = own | auto | library.
      own      every attention product runs on csrc/attention.cu (tcgen05 / TMEM / TMA);
      auto     (default) csrc/attention.cu wherever the reference's dispatcher would have called
               flash-attn 2 -- sliding window, soft-capping, packed rows, head_dim 256 (1.4-2.3x faster
               than flash-attn 2 on a B200, benchmarks/attn_bench.py) -- and cuDNN's fused attention
               through torch SDPA for plain dense causal attention, where cuDNN is still ahead;
      library  round-1 behaviour (cuDNN SDPA / flash-attn 2 only)."""
    return os.environ.get("UB200_ATTENTION", "auto")


def _attention(Q, K_, V, scale, window, softcap, seq_info=None):
    """The attention product on [B, S, H, D] views of the projection buffers (GQA native, no copies);
    the reference's dispatcher is utils/attention_dispatch.py:298-617 (flash-attn | xformers | SDPA;
    packed rows :433-447)."""
    policy = _attention_policy()
    special = seq_info is not None or window != (-1, -1) or bool(softcap) or Q.shape[-1] == 256
    if policy == "own" or (policy == "auto" and special):
        return K.fast_attention(Q, K_, V, scale, window, softcap, seq_info)
    if seq_info is not None:
        from flash_attn import flash_attn_varlen_func
        bsz, q_len, n_heads, hd = Q.shape
        _, cu, max_len = seq_info
        o = flash_attn_varlen_func(Q.reshape(bsz * q_len, n_heads, hd),
                                   K_.reshape(bsz * q_len, K_.shape[2], hd),
                                   V.reshape(bsz * q_len, V.shape[2], hd), cu, cu, max_len, max_len,
                                   dropout_p=0.0, softmax_scale=scale, causal=True,
                                   window_size=window, softcap=softcap)
        return o.view(bsz, q_len, n_heads, hd)
    if window == (-1, -1) and not softcap and _SDPA_CUDNN["ok"] is not False:
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        try:
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                o = F.scaled_dot_product_attention(Q.transpose(1, 2), K_.transpose(1, 2), V.transpose(1, 2),
                                                   is_causal=True, scale=scale, enable_gqa=True)
            _SDPA_CUDNN["ok"] = True
            return o.transpose(1, 2)
        except RuntimeError:
            if _SDPA_CUDNN["ok"]:
                raise
            _SDPA_CUDNN["ok"] = False          # backend unavailable for this shape: use our kernel
            if policy != "library":
                return K.fast_attention(Q, K_, V, scale, window, softcap, seq_info)
    from flash_attn import flash_attn_func
    return flash_attn_func(Q, K_, V, dropout_p=0.0, softmax_scale=scale, causal=True,
                           window_size=window, softcap=softcap)


FUSE_ADD_NORM = True      # tests flip this to compare against the layer-by-layer form


def LlamaAttention_fast_forward(self, hidden_states, cos, sin, position_ids=None, seq_info=None):
    """models/llama.py:671-771 (training branch), mistral.py:61-157, gemma2.py:86-201."""
    bsz, q_len, _ = hidden_states.size()
    n_heads, n_kv, hd = self._ub_heads
    Q, Kt, V = self.apply_qkv(self, hidden_states)                      # llama.py:702
    Q = Q.view(bsz, q_len, n_heads, hd).transpose(1, 2)                 # views, no copies
    Kt = Kt.view(bsz, q_len, n_kv, hd).transpose(1, 2)
    Q, Kt = K.fast_rope_embedding(Q, Kt, cos, sin, position_ids)        # llama.py:730, in place
    sink = getattr(self, "_ub_kv_sink", None)
    if sink is not None:        # prefill of generate(): keep the rotated K and V for the KV cache
        sink.append((Kt.detach(), V.view(bsz, q_len, n_kv, hd).transpose(1, 2).detach()))
    window = (-1, -1)
    sw = self._ub_window
    longest = q_len if seq_info is None else seq_info[2]
    if sw is not None and longest > sw:                                 # mistral.py:112-120
        window = (sw, sw)
    A = _attention(Q.transpose(1, 2), Kt.transpose(1, 2), V.view(bsz, q_len, n_kv, hd),
                   self._ub_scale, window, self._ub_softcap, seq_info)
    return self.apply_o(self, A.reshape(bsz, q_len, n_heads * hd))      # llama.py:768


def DecoderLayer_fast_forward(self, hidden_states, cos, sin, position_ids=None, seq_info=None):
    """models/llama.py:823-844 (Llama / Mistral) and gemma2.py:258-283 (four norms)."""
    gemma = self._ub_gemma
    residual = hidden_states
    h = K.fast_rms_layernorm(self.input_layernorm, hidden_states, gemma=gemma)
    h = LlamaAttention_fast_forward(self.self_attn, h, cos, sin, position_ids, seq_info)
    if gemma:
        h = K.fast_rms_layernorm(self.post_attention_layernorm, h, gemma=True)
    hidden_states = residual + h
    residual = hidden_states
    if gemma:
        h = K.fast_rms_layernorm(self.pre_feedforward_layernorm, hidden_states, gemma=True)
        h = self.mlp(h)
        h = K.fast_rms_layernorm(self.post_feedforward_layernorm, h, gemma=True)
    else:
        h = K.fast_rms_layernorm(self.post_attention_layernorm, hidden_states)
        h = self.mlp(h)
    return residual + h


def Model_fast_forward(self, input_ids, position_ids=None, packed_seq_lengths=None):
    """models/llama.py:866-1230: embed, (Gemma: * sqrt(H) in model dtype :961-989), layers, norm."""
    h = self.embed_tokens(input_ids)
    if self._ub_gemma and not hasattr(self.embed_tokens, "embed_scale"):
        # transformers >= 4.5x wraps the Gemma embedding in a *ScaledWordEmbedding that already
        # multiplies by sqrt(H); older versions (the reference's llama.py:961-989) scale here
        h = h * torch.tensor(math.sqrt(self.config.hidden_size), dtype=h.dtype, device=h.device)
    seq_len = input_ids.shape[1]
    need = seq_len
    # packed / padding-free rows (llama.py:706, :721-723): per-document attention through
    # cu_seqlens and RoPE through the reset-style position ids (derived when the collator
    # did not send them)
    seq_info = get_packed_info_from_kwargs({"packed_seq_lengths": packed_seq_lengths}, h.device,
                                           input_ids.numel())
    derived = seq_info is not None and position_ids is None
    if derived:
        position_ids = packed_position_ids(packed_seq_lengths, input_ids.numel(), h.device)
    idx = None
    if position_ids is not None:
        idx = position_ids.reshape(-1).to(torch.int32)
        if idx.numel() != input_ids.numel():
            raise ValueError("unsloth_b200: position_ids must hold one entry per token (%d != %d)"
                             % (idx.numel(), input_ids.numel()))
        # caller-supplied positions may exceed the row length: size the table for the model's
        # position range (no device sync to read the actual maximum); packed positions are < seq_len
        if not derived:
            need = max(need, int(getattr(self.config, "max_position_embeddings", 0) or 0))
    cos, sin = self._ub_rotary.get(need)
    # optional recompute of every decoder layer in the backward (the plain-recompute part of the
    # reference's use_gradient_checkpointing, models/llama.py:1169-1192; no activation offload):
    # not needed for the BASELINE configs on a 180 GB B200, available for longer sequences
    ckpt = getattr(self, "_ub_gradient_checkpointing", False) and torch.is_grad_enabled()
    if ckpt == "unsloth" and seq_len < 512:       # _utils.py:360-386: offloading pays from ~512 tokens
        ckpt = True
    if ckpt == "unsloth":
        from .checkpoint import offloaded_checkpoint

        def checkpoint(fn, layer, hidden, *rest, **kw):              # fn = DecoderLayer_fast_forward
            return offloaded_checkpoint(lambda h, *r: fn(layer, h, *r), hidden, *rest)
    elif ckpt:
        from torch.utils.checkpoint import checkpoint
    if self._ub_gemma or not FUSE_ADD_NORM or ckpt == "unsloth":
        # (offloaded checkpoints keep exactly ONE tensor per layer: the layer-by-layer form, whose
        # only live input is the hidden state)
        for layer in self.layers:
            if ckpt:
                h = checkpoint(DecoderLayer_fast_forward, layer, h, cos, sin, idx, seq_info,
                               use_reentrant=False, preserve_rng_state=False)
            else:
                h = DecoderLayer_fast_forward(layer, h, cos, sin, idx, seq_info)
        return K.fast_rms_layernorm(self.norm, h, gemma=self._ub_gemma)
    # Llama / Mistral: every `residual + branch` is fused with the norm that consumes the sum
    # (this layer's post_attention_layernorm, then the NEXT layer's input_layernorm or the final
    # norm), so the residual stream is read and written once per norm instead of twice.
    layers = list(self.layers)
    residual = h
    normed = K.fast_rms_layernorm(layers[0].input_layernorm, h)
    def layer_step(i, residual, normed):
        layer = layers[i]
        a = LlamaAttention_fast_forward(layer.self_attn, normed, cos, sin, idx, seq_info)
        residual, normed = K.fast_add_rms_layernorm(layer.post_attention_layernorm, residual, a)
        m = layer.mlp(normed)
        nxt = layers[i + 1].input_layernorm if i + 1 < len(layers) else self.norm
        return K.fast_add_rms_layernorm(nxt, residual, m)

    for i in range(len(layers)):
        if ckpt:
            residual, normed = checkpoint(layer_step, i, residual, normed, use_reentrant=False,
                                          preserve_rng_state=False)
        else:
            residual, normed = layer_step(i, residual, normed)
    return normed


def CausalLM_fast_forward(self, input_ids=None, labels=None, position_ids=None,
                          packed_seq_lengths=None, num_items_in_batch=None, **kwargs):
    """models/llama.py:1371-1590, the `labels is not None and not UNSLOTH_RETURN_LOGITS` branch:
    boundary-mask packed labels (:1483), logits-free fused CE (:1497-1509), EMPTY logits."""
    hidden = Model_fast_forward(self.model, input_ids, position_ids, packed_seq_lengths)
    if os.environ.get("UNSLOTH_RETURN_LOGITS", "0") == "1":
        # models/llama.py:1525-1562: materialise the logits (one tcgen05 GEMM against lm_head),
        # shift the labels here, guard the packed boundaries on the SHIFTED labels, Triton-style CE.
        logits = K.LoRA_W.apply(hidden, self.lm_head.weight, None, None, None, None)
        loss = None
        if labels is not None:
            shift_labels = torch.empty_like(labels)
            shift_labels[..., :-1] = labels[..., 1:]
            shift_labels[..., -1] = -100
            mask_packed_sequence_boundaries(shift_labels, packed_seq_lengths)
            loss = K.fast_cross_entropy_loss(logits=logits, labels=shift_labels,
                                             logit_softcapping=self._ub_final_softcap,
                                             logit_scaling=0, n_items=num_items_in_batch)
        return SimpleNamespace(loss=loss, logits=logits, hidden_states=None)
    if labels is None:
        return SimpleNamespace(loss=None, logits=None, hidden_states=hidden)
    labels = mask_packed_boundary_labels(labels, packed_seq_lengths)
    loss = K.unsloth_fused_ce_loss(
        trainer=None, hidden_states=hidden, lm_head_weight=self.lm_head.weight,
        lm_head_bias=getattr(self.lm_head, "bias", None),
        labels=labels, mask=None, n_items=num_items_in_batch,
        scaling=getattr(self, "accelerator_scaler", None), target_gb=None,
        torch_compile=False, logit_softcapping=self._ub_final_softcap)
    return SimpleNamespace(loss=loss, logits=None, hidden_states=None)


# ---------------------------------------------------------------------------------------------
# install: the per-instance rebinding of patch_peft_model (models/llama.py:3599-3770)
# ---------------------------------------------------------------------------------------------
def _arch_of(model):
    mt = getattr(model.config, "model_type", "llama")
    if mt not in ("llama", "mistral", "gemma2"):
        raise NotImplementedError("unsloth_b200: model_type %r is outside the hot-path scope" % mt)
    return mt


def install(model, gradient_checkpointing=False, tiled_mlp=0):
    """Rebind a HuggingFace Llama / Mistral / Gemma-2 CausalLM (with LoRA-wrapped projections)
    onto the unsloth_b200 kernels.  Returns the model.  `gradient_checkpointing`: False (default: the
    BASELINE configs fit a B200 without), True (recompute each decoder layer in the backward) or
    "unsloth" (the same with the layer inputs offloaded to pinned host memory); `tiled_mlp` = n > 1:
    the MLP runs over n token shards and is recomputed shard-wise (checkpoint.py)."""
    arch = _arch_of(model)
    cfg = model.config
    gemma = arch == "gemma2"
    inner = model.model
    dev = inner.embed_tokens.weight.device
    dtype = inner.embed_tokens.weight.dtype
    hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    rope_theta = getattr(cfg, "rope_theta", None)
    rope_scaling = getattr(cfg, "rope_scaling", None)
    rp = getattr(cfg, "rope_parameters", None)
    if isinstance(rp, dict):
        rope_theta = rp.get("rope_theta", rope_theta)
        if rp.get("rope_type", "default") != "default":
            rope_scaling = rp
    inner._ub_rotary = RotaryCache(hd, float(rope_theta or 10000.0), dev,
                                   torch.float32 if gemma else dtype, rope_scaling)
    inner._ub_gemma = gemma
    inner._ub_gradient_checkpointing = gradient_checkpointing if gradient_checkpointing == "unsloth" \
        else bool(gradient_checkpointing)
    if gradient_checkpointing:
        # per-layer recompute owns the memory: do not keep 16-bit expansions alive across layers
        from .kernels import utils as _KU
        _KU.KEEP_DEQUANT_BLOCKED = True
    model._ub_final_softcap = float(getattr(cfg, "final_logit_softcapping", 0) or 0)
    mlp_fn = K.apply_lora_mlp_geglu_approx if gemma else K.apply_lora_mlp_swiglu   # llama.py:3618-3639
    for i, layer in enumerate(inner.layers):
        attn = layer.self_attn
        layer._ub_gemma = gemma
        attn._ub_heads = (cfg.num_attention_heads, cfg.num_key_value_heads, hd)
        if gemma:
            attn._ub_scale = float(cfg.query_pre_attn_scalar) ** -0.5
            attn._ub_softcap = float(getattr(cfg, "attn_logit_softcapping", 0) or 0)
            attn._ub_window = cfg.sliding_window if i % 2 == 0 else None             # gemma2.py:139-150
        else:
            attn._ub_scale = hd ** -0.5
            attn._ub_softcap = 0.0
            attn._ub_window = getattr(cfg, "sliding_window", None) if arch == "mistral" else None
        layer.mlp.forward = types.MethodType(mlp_fn, layer.mlp)                      # llama.py:3725
        if tiled_mlp and tiled_mlp > 1:                                              # llama.py:3719-3723
            from .checkpoint import tiled_mlp_forward
            layer.mlp._unsloth_forward = layer.mlp.forward
            layer.mlp.forward = tiled_mlp_forward(layer.mlp._unsloth_forward, int(tiled_mlp))
        attn.apply_qkv = K.apply_lora_qkv                                            # llama.py:3748
        attn.apply_o = K.apply_lora_o                                                # llama.py:3766
    model.forward = types.MethodType(CausalLM_fast_forward, model)
    return model


def lora_parameters(model):
    return [p for n, p in model.named_parameters() if ("lora_A" in n or "lora_B" in n)]


TARGET_MODULES = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def attach_qlora(model, r=16, lora_alpha=16, init_b_std=0.0, quantize=True):
    """NF4-quantise (bitsandbytes layout) and LoRA-wrap the seven projections of every layer:
    get_peft_model defaults of models/llama.py:3061-3075 (r=16, alpha=16, dropout 0, bias none)."""
    for layer in model.model.layers:
        for parent in (layer.self_attn, layer.mlp):
            for name in TARGET_MODULES:
                lin = getattr(parent, name, None)
                if lin is None or isinstance(lin, LoraLinear):
                    continue
                if quantize:
                    base = Linear4bit.from_dense(lin.weight.data)
                else:
                    base = lin
                setattr(parent, name, LoraLinear(base, r=r, lora_alpha=lora_alpha,
                                                 init_b_std=init_b_std))
                del lin
    for p in model.parameters():
        p.requires_grad_(False)
    for p in lora_parameters(model):
        p.requires_grad_(True)
    return model


def build_qlora_model(name="llama-3-8b", r=16, lora_alpha=16, device="cuda", dtype=torch.bfloat16,
                      seed=3407, init_b_std=0.0, num_hidden_layers=None, quantize=True,
                      gradient_checkpointing=False, tiled_mlp=0, **overrides):
    """Random-init model of a BASELINE.json config on `device`, NF4 + LoRA, kernels installed."""
    from transformers import AutoModelForCausalLM
    cfg = hf_config(name, num_hidden_layers, **overrides)
    torch.manual_seed(seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = AutoModelForCausalLM.from_config(cfg)
    finally:
        torch.set_default_dtype(prev)
    model.to(dtype)
    attach_qlora(model, r=r, lora_alpha=lora_alpha, init_b_std=init_b_std, quantize=quantize)
    torch.cuda.empty_cache()
    return install(model, gradient_checkpointing=gradient_checkpointing, tiled_mlp=tiled_mlp)
