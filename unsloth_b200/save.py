"""LoRA merge for 16-bit export: the consumer of `fast_dequantize` in unsloth/save.py:620-646
(`_merge_lora`), SURVEY.md 8f rank 4.  With libunsloth_b200.so behind fast_dequantize the export
path needs no bitsandbytes.  Not on the training hot path: the rank-r update is one fp32 `addmm_`
(torch / cuBLAS, exactly the reference's call), the NF4 expansion is our kernel.
"""
from __future__ import annotations

import torch

from .kernels.utils import fast_dequantize, get_lora_parameters_bias


@torch.inference_mode
def merge_lora(layer, name: str = ""):
    """Returns (W_merged [out, in] in the weight's 16-bit dtype, bias).  Same arithmetic as the
    reference's `_merge_lora` (save.py:629-645): the update s * B @ A is added to the dequantised
    weight in fp32 and the sum is rounded ONCE to the storage dtype; a non-finite result raises."""
    wrapped = hasattr(layer, "base_layer") or hasattr(getattr(layer, "weight", None), "quant_state")
    if not wrapped:
        return layer.weight, getattr(layer, "bias", None)
    W, quant_state, A, B, s, bias = get_lora_parameters_bias(layer)
    if quant_state is None:
        out_dtype, dense = W.dtype, W
    else:
        out_dtype = quant_state[2] if type(quant_state) is list else quant_state.dtype
        dense = fast_dequantize(W, quant_state)                 # ub200_dequantize_nf4
    merged = dense.float()                                       # [out, in], fp32
    if A is not None:
        merged = torch.addmm(merged, B.float(), A.float(), alpha=s)
        if not bool(torch.isfinite(merged.abs().max())):
            raise ValueError("unsloth_b200: merging the adapter of %r produced non-finite weights" % name)
    return merged.to(out_dtype), bias


def merged_state_dict(model):
    """16-bit state dict of a QLoRA model with every adapter folded into its base weight
    (the tensor set `unsloth_save_model(..., save_method="merged_16bit")` writes, save.py:700-760):
    `...q_proj.weight` instead of `...q_proj.base_layer.weight` + lora_A / lora_B."""
    out = {}
    lora_prefixes = [n for n, m in model.named_modules() if hasattr(m, "base_layer")]
    for n in lora_prefixes:
        W, bias = merge_lora(model.get_submodule(n), n)
        out[n + ".weight"] = W.contiguous()
        if bias is not None:
            out[n + ".bias"] = bias
    for k, v in model.state_dict().items():
        if any(k.startswith(p_ + ".") for p_ in lora_prefixes):
            continue
        out[k] = v
    return out
