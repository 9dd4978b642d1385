"""BASELINE infrastructure (never imported by the product path): the REFERENCE's GPU path on a B200.

BASELINE.md B2 (the end-to-end Unsloth step) needs `unsloth_zoo`, `peft`, `trl`, `bitsandbytes`,
none installable offline.  What CAN run here is everything the reference itself executes on the
device for this path, loaded unmodified from its install under baseline/_ref through
oracle/ref_shim.load_reference_kernels_native():

  * its Triton kernels (RMSNorm, RoPE, SwiGLU / GEGLU, cross entropy) compiled by Triton 3.6 for
    the B200,
  * its `LoRA_MLP` / `LoRA_QKV` / `LoRA_W` autograd functions with their torch.matmul / addmm_
    (cuBLAS) schedule, and its `fast_dequantize` -- whose five bitsandbytes C symbols resolve to the
    same-signature exports of libunsloth_b200.so (bitsandbytes is absent),
  * attention through flash-attn 2, which is what its dispatcher picks when flash-attn is installed
    (utils/attention_dispatch.py:298-447), or cuDNN SDPA as a best-case alternative,
  * the loss through its materialised-logits route (models/llama.py:1525-1562: lm_head matmul ->
    shift -> `fast_cross_entropy_loss`), because the default logits-free loss lives in unsloth_zoo,
  * launched eagerly (the reference does not capture CUDA graphs), AdamW via torch.optim.

The per-instance binding below is the one `patch_peft_model` performs (models/llama.py:3599-3770);
the slim fast forwards of unsloth_b200/patch.py are reused for the module glue (the reference's own
`*_fast_forward` functions live in models/llama.py, which imports the absent packages) with every
kernel name pointing at the REFERENCE implementation.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def reference_namespace(refk):
    """The names patch.py's fast forwards call, bound to the reference's implementations."""
    fl = refk.fast_lora
    return SimpleNamespace(
        fast_rms_layernorm=refk.rms_layernorm.fast_rms_layernorm,
        fast_rope_embedding=refk.rope_embedding.fast_rope_embedding,
        fast_cross_entropy_loss=refk.cross_entropy_loss.fast_cross_entropy_loss,
        apply_lora_mlp_swiglu=fl.apply_lora_mlp_swiglu,
        apply_lora_mlp_geglu_approx=fl.apply_lora_mlp_geglu_approx,
        apply_lora_qkv=fl.apply_lora_qkv, apply_lora_o=fl.apply_lora_o)


class ReferenceStep:
    """One QLoRA training step of `model` (built by unsloth_b200.patch.build_qlora_model) executed
    through the reference's kernels.  `attention`: "flash" (the reference's priority) or "sdpa"."""

    def __init__(self, model, refk, attention="flash", lr=2e-4, weight_decay=0.01):
        from unsloth_b200 import patch as P
        self.P, self.model, self.ns = P, model, reference_namespace(refk)
        self.attention = attention
        self.params = P.lora_parameters(model)
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, fused=True)
        gemma = model.model._ub_gemma
        self.mlp_fn = self.ns.apply_lora_mlp_geglu_approx if gemma else self.ns.apply_lora_mlp_swiglu

    # -- (un)binding ------------------------------------------------------------------------------
    def _bind(self):
        P = self.P
        self._saved = (P.K, P.FUSE_ADD_NORM, P._attention)
        P.K, P.FUSE_ADD_NORM = self.ns, False
        if self.attention == "flash":
            def attn(Q, K_, V, scale, window, softcap, seq_info=None):
                from flash_attn import flash_attn_func
                return flash_attn_func(Q, K_, V, dropout_p=0.0, softmax_scale=scale, causal=True,
                                       window_size=window, softcap=softcap)
            P._attention = attn
        self._saved_layers = []
        for layer in self.model.model.layers:
            a = layer.self_attn
            self._saved_layers.append((layer.mlp.forward, a.apply_qkv, a.apply_o))
            layer.mlp.forward = types.MethodType(self.mlp_fn, layer.mlp)          # llama.py:3725
            a.apply_qkv, a.apply_o = self.ns.apply_lora_qkv, self.ns.apply_lora_o  # :3748, :3766

    def _unbind(self):
        P = self.P
        P.K, P.FUSE_ADD_NORM, P._attention = self._saved
        for layer, (m, q, o) in zip(self.model.model.layers, self._saved_layers):
            layer.mlp.forward, layer.self_attn.apply_qkv, layer.self_attn.apply_o = m, q, o

    def __enter__(self):
        self._bind()
        return self

    def __exit__(self, *exc):
        self._unbind()

    # -- the step ---------------------------------------------------------------------------------
    def loss(self, input_ids, labels):
        model, P = self.model, self.P
        hidden = P.Model_fast_forward(model.model, input_ids)
        logits = model.lm_head(hidden)                                            # llama.py:1525
        shift = torch.empty_like(labels)                                          # :1545-1547
        shift[..., :-1] = labels[..., 1:]
        shift[..., -1] = -100
        return self.ns.fast_cross_entropy_loss(logits=logits, labels=shift,
                                               logit_softcapping=model._ub_final_softcap,
                                               logit_scaling=0, n_items=None)

    def step(self, input_ids, labels):
        self.opt.zero_grad(set_to_none=False)   # grads may be views of a flat bucket: keep them
        loss = self.loss(input_ids, labels)
        loss.backward()
        self.opt.step()
        return loss.detach()


def time_reference_step(model, ids, labels, steps=5, warmup=3, attention="flash"):
    """tokens/s, ms/step, peak GiB, last loss of the reference path on `model` (device-resident
    batches [n, bs, seq]).  The model's LoRA parameters are restored afterwards."""
    from oracle import ref_shim
    refk = ref_shim.load_reference_kernels_native()
    saved = [p.detach().clone() for p in model.parameters() if p.requires_grad]
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    with ReferenceStep(model, refk, attention) as rs:
        n = ids.shape[0]
        for i in range(warmup):
            rs.step(ids[i % n], labels[i % n])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            loss = rs.step(ids[(warmup + i) % n], labels[(warmup + i) % n])
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
    peak = torch.cuda.max_memory_allocated() / 2 ** 30
    with torch.no_grad():
        for p, q in zip([p for p in model.parameters() if p.requires_grad], saved):
            p.copy_(q)
    from unsloth_b200.kernels.utils import bump_param_epoch
    bump_param_epoch()
    tokens = ids.shape[1] * ids.shape[2]
    return {"tokens_per_s": tokens / ms * 1e3, "ms_per_step": ms, "peak_vram_gib": peak,
            "loss_last": float(loss.item()), "attention": attention, "steps": steps, "warmup": warmup}
