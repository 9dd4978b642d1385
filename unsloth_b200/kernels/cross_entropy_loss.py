"""Cross entropy -- host-side mirror of unsloth/kernels/cross_entropy_loss.py:288-449 and of
the external logits-free loss the reference calls at models/llama.py:1497-1509
(`unsloth_zoo.loss_utils.unsloth_fused_ce_loss`).

  * `Fast_CrossEntropyLoss` / `fast_cross_entropy_loss`: materialised logits, in-place gradient
    (the gradient overwrites the logits buffer, cross_entropy_loss.py:380-418).
  * `unsloth_fused_ce_loss`: chunked linear + CE that never materialises the full [T, V] logits:
    per row-chunk a tcgen05 GEMM produces the logits chunk, the CE kernels turn it into the
    gradient chunk in place, and a second GEMM (lm_head consumed as an MN-major operand)
    reduces it to dHidden.  Labels are shifted inside (llama.py:1479-1482).
"""
from __future__ import annotations

import torch

from .. import _lib as L
from .utils import MAX_FUSED_SIZE, GradModeAware, gemm, outer_grad_enabled  # noqa: F401


def _ce_forward(logits2, labels1, softcap, scale):
    n_rows, vocab = logits2.shape
    losses = torch.empty(n_rows, dtype=torch.float32, device=logits2.device)
    lse = torch.empty(n_rows, dtype=torch.float32, device=logits2.device)
    L.call("ub200_cross_entropy_fwd", L.ptr(logits2), logits2.stride(0), L.ptr(labels1), L.ptr(losses),
           L.ptr(lse), n_rows, vocab, float(softcap), float(scale), L.dt(logits2), L.stream())
    return losses, lse


def _ce_backward_(logits2, lse, labels1, dloss, dloss_stride, softcap, scale):
    n_rows, vocab = logits2.shape
    L.call("ub200_cross_entropy_bwd", L.ptr(logits2), logits2.stride(0), L.ptr(lse), L.ptr(labels1),
           L.ptr(dloss), int(dloss_stride), n_rows, vocab, float(softcap), float(scale),
           L.dt(logits2), L.stream())
    return logits2


class Fast_CrossEntropyLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, logit_softcapping: float = 0, logit_scaling: float = 0):
        L.require_cuda(logits)
        assert logits.dim() == 2 and logits.stride(-1) == 1
        labels = labels.to(logits.device)
        if labels.dtype != torch.int64:
            labels = labels.to(torch.int64)
        labels = labels.contiguous()
        losses, lse = _ce_forward(logits, labels, logit_softcapping or 0.0, logit_scaling or 0.0)
        ctx.save_for_backward(logits, lse, labels)
        ctx.logit_softcapping = logit_softcapping or 0.0
        ctx.logit_scaling = logit_scaling or 0.0
        return losses

    @staticmethod
    def backward(ctx, dlosses):
        logits, lse, labels = ctx.saved_tensors
        dl = dlosses if dlosses.dtype == torch.float32 else dlosses.float()
        stride = dl.stride(0) if dl.numel() > 1 else 0
        _ce_backward_(logits, lse, labels, dl, stride, ctx.logit_softcapping, ctx.logit_scaling)
        return logits, None, None, None


def fast_cross_entropy_loss(logits, labels, logit_softcapping=0, logit_scaling=0, n_items=None):
    """cross_entropy_loss.py:421-449.  logits [B,S,V], labels [B,S] (already shifted)."""
    batch, seq_len, d = logits.shape
    assert labels.shape == (batch, seq_len)
    device = logits.device
    loss = Fast_CrossEntropyLoss.apply(logits.view(batch * seq_len, d), labels.view(-1),
                                       logit_softcapping, logit_scaling)
    if n_items is None:
        n_items = torch.count_nonzero(labels != -100)
    if torch.is_tensor(n_items):
        n_items = n_items.to(device)
    return loss.sum() / n_items


class Fused_Linear_CrossEntropy(GradModeAware, torch.autograd.Function):
    """hidden [T,H] x lm_head [V,H] (+ bias [V]) -> scalar sum of per-row losses / n_items.  dHidden
    (and dW / dbias when lm_head is trainable) are computed in the same pass as the gradient of
    `loss_scale * mean loss` and rescaled by `dloss / loss_scale` in backward: with an fp16
    GradScaler the softmax-gradient chunk and dH are produced at the SCALED magnitude, as they
    would be by autograd on a scaled loss, instead of underflowing at 1/n_items."""

    @staticmethod
    def forward(ctx, hidden2, weight, bias, labels1, inv_n, softcap, scale, chunk_rows, loss_scale):
        L.require_cuda(hidden2, weight)
        T, H = hidden2.shape
        V = weight.shape[0]
        dev, dt = hidden2.device, hidden2.dtype
        W = weight if weight.dtype == dt else weight.to(dt)
        W = W if W.stride(-1) == 1 else W.contiguous()
        need_dW = bool(ctx.needs_input_grad[1]) and outer_grad_enabled()
        need_db = bias is not None and bool(ctx.needs_input_grad[2]) and outer_grad_enabled()
        need_dH = bool(ctx.needs_input_grad[0]) and outer_grad_enabled()
        losses = torch.empty(T, dtype=torch.float32, device=dev)
        dH = torch.empty((T, H), dtype=dt, device=dev) if need_dH else None
        dW = torch.zeros((V, H), dtype=torch.float32, device=dev) if need_dW else None
        db = torch.zeros(V, dtype=torch.float32, device=dev) if need_db else None
        chunk_rows = max(128, min(chunk_rows, T))
        # the chunk is a TMA operand of the dH GEMM (row stride * 2 B must be a multiple of 16):
        # pad the row STRIDE to a multiple of 8 for vocabularies such as 32001 / 50257
        Vp = (V + 7) // 8 * 8
        buf = torch.empty((chunk_rows, Vp), dtype=dt, device=dev)[:, :V]
        g_scale = inv_n if loss_scale == 1.0 else inv_n * loss_scale
        bias_c = None if bias is None else bias.to(dt)
        for r0 in range(0, T, chunk_rows):
            r1 = min(T, r0 + chunk_rows)
            n = r1 - r0
            logits = buf[:n]
            gemm(n, V, [(hidden2[r0:r1], W, H)], logits)
            if bias_c is not None:
                logits += bias_c                                  # nn.Linear: bf16 add after the GEMM
            lab = labels1[r0:r1]
            l, lse = _ce_forward(logits, lab, softcap, scale)
            losses[r0:r1] = l
            if not (need_dH or need_dW or need_db):
                continue
            _ce_backward_(logits, lse, lab, g_scale, 0, softcap, scale)      # logits <- d logits
            if need_dH:
                gemm(n, H, [(logits, W, V)], dH[r0:r1], a_mn=False, b_mn=True)  # d hidden
            if need_dW:                                            # dW += dlogits^T @ hidden (fp32)
                gemm(V, H, [(logits, hidden2[r0:r1], n)], dW, a_mn=True, b_mn=True, accumulate=True)
            if need_db:
                db += logits.float().sum(0)
        ctx.save_for_backward(dH, dW, db)
        ctx.loss_scale = loss_scale
        ctx.w_dtype, ctx.b_dtype = weight.dtype, (None if bias is None else bias.dtype)
        return losses.sum() * inv_n.squeeze()

    @staticmethod
    def backward(ctx, dloss):
        dH, dW, db = ctx.saved_tensors
        g = dloss / ctx.loss_scale
        return (None if dH is None else dH * g.to(dH.dtype),
                None if dW is None else (dW * g).to(ctx.w_dtype),
                None if db is None else (db * g).to(ctx.b_dtype), None, None, None, None, None, None)


def unsloth_fused_ce_loss(trainer=None, hidden_states=None, lm_head_weight=None, lm_head_bias=None,
                          labels=None, mask=None, n_items=None, scaling=None, target_gb=None,
                          torch_compile=False, logit_softcapping=0, chunk_rows=2048,
                          logit_scaling=0, **kwargs):
    """Drop-in for `unsloth_zoo.loss_utils.unsloth_fused_ce_loss` at the reference call sites
    models/llama.py:1497-1509 and models/mistral.py:352-364.  hidden_states [B,S,H], labels [B,S]
    UNSHIFTED; divides by n_items (count of non-ignored shifted labels when None).  Returns a
    fresh 0-d tensor (HF Trainer multiplies the loss in place, models/_utils.py:3200-3225).

    `scaling` is what the call sites pass: `getattr(self, "accelerator_scaler", None)`, the fp16
    GradScaler (or None) -- its current scale is folded into the in-pass gradient so fp16 softmax
    gradients do not underflow; a plain number is taken as the loss scale itself.  Logit scaling
    (Cohere / Granite style `logits * s`) is the separate `logit_scaling` argument.  A trainable
    lm_head and an lm_head bias are supported (gradients in fp32, one extra tcgen05 GEMM per
    chunk)."""
    B, S, H = hidden_states.shape
    dev = hidden_states.device
    shift = torch.full_like(labels, -100)
    shift[..., :-1] = labels[..., 1:]
    if mask is not None:
        shift = torch.where(mask.to(torch.bool), shift, torch.full_like(shift, -100))
    shift = shift.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
    if n_items is None:
        n_items = torch.count_nonzero(shift != -100)
    if torch.is_tensor(n_items):
        inv_n = (1.0 / n_items.to(device=dev, dtype=torch.float32)).reshape(1)
    else:
        inv_n = torch.full((1,), 1.0 / float(n_items), dtype=torch.float32, device=dev)
    loss_scale = 1.0
    if scaling is not None:
        if hasattr(scaling, "get_scale"):
            loss_scale = float(scaling.get_scale()) if getattr(scaling, "is_enabled", lambda: True)() else 1.0
        else:
            loss_scale = float(scaling) or 1.0
    h2 = hidden_states.reshape(-1, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    return Fused_Linear_CrossEntropy.apply(h2, lm_head_weight, lm_head_bias, shift, inv_n,
                                           float(logit_softcapping or 0), float(logit_scaling or 0),
                                           int(chunk_rows), loss_scale)


# ---------------------------------------------------------------------------------------------
# class-level route (cross_entropy_loss.py:459-473): serve a STOCK HuggingFace model
# ---------------------------------------------------------------------------------------------
_HF_LOSS_BACKUP = {}


def UnslothForCausalLMLoss(logits, labels, vocab_size=None, num_items_in_batch=None, ignore_index=-100,
                           shift_labels=None, **kwargs):
    """Drop-in for transformers.loss.loss_utils.ForCausalLMLoss on materialised logits: same label
    shift, then `fast_cross_entropy_loss` (no fp32 upcast copy of the logits; the gradient is
    written in place into them)."""
    if ignore_index != -100:
        raise NotImplementedError("unsloth_b200: the CE kernels use ignore_index = -100")
    if shift_labels is None:
        labels = torch.nn.functional.pad(labels, (0, 1), value=-100)
        shift_labels = labels[..., 1:].contiguous()
    if logits.dim() == 2:
        logits = logits.unsqueeze(0)
    shift_labels = shift_labels.reshape(logits.shape[0], logits.shape[1]).to(logits.device)
    return fast_cross_entropy_loss(logits, shift_labels, n_items=num_items_in_batch)


def patch_loss_functions(torch_compile=False):
    """cross_entropy_loss.py:459-473: point HF's LOSS_MAPPING["ForCausalLM"] (and the aliases that
    still hold the stock function) at the kernel-backed loss.  `torch_compile` is accepted and
    ignored: nothing here goes through a tracing compiler."""
    import transformers.loss.loss_utils as lu
    for key, fn in list(lu.LOSS_MAPPING.items()):
        if getattr(fn, "__name__", "") == "ForCausalLMLoss":
            _HF_LOSS_BACKUP.setdefault(key, fn)
            lu.LOSS_MAPPING[key] = UnslothForCausalLMLoss


def unpatch_loss_functions():
    import transformers.loss.loss_utils as lu
    for key, fn in _HF_LOSS_BACKUP.items():
        lu.LOSS_MAPPING[key] = fn
    _HF_LOSS_BACKUP.clear()
