"""Minimal PEFT-compatible LoRA wrapper (peft is not installed offline).

Exposes exactly the attribute surface the reference's `get_lora_parameters` scrapes
(unsloth/kernels/utils.py:335-397): `.base_layer.weight` (+ `.quant_state`),
`.lora_A[adapter].weight [r, in]`, `.lora_B[adapter].weight [out, r]`, `.scaling[adapter]`,
`.active_adapters`, `.disable_adapters`, `.merged`.  LoRA A/B are kept in fp32
(models/_utils.py:2482-2496, float32_mixed_precision=True); the kernels cast per call.
Initialisation as PEFT: A ~ kaiming_uniform(a=sqrt(5)), B = 0.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class LoraLinear(nn.Module):
    def __init__(self, base_layer: nn.Module, r: int = 16, lora_alpha: int = 16,
                 adapter_name: str = "default", device=None, init_b_std: float = 0.0):
        super().__init__()
        self.base_layer = base_layer
        in_f = getattr(base_layer, "in_features")
        out_f = getattr(base_layer, "out_features")
        dev = device if device is not None else base_layer.weight.device
        self.in_features, self.out_features = in_f, out_f
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(in_f, r, bias=False, device=dev,
                                                             dtype=torch.float32)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, out_f, bias=False, device=dev,
                                                             dtype=torch.float32)})
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        if init_b_std > 0:
            nn.init.normal_(self.lora_B[adapter_name].weight, std=init_b_std)
        else:
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.scaling = {adapter_name: lora_alpha / r}
        self.r = {adapter_name: r}
        self.lora_dropout = nn.ModuleDict({adapter_name: nn.Identity()})
        self.use_dora = {adapter_name: False}
        self.active_adapters = [adapter_name]
        self.disable_adapters = False
        self.merged = False
        for p in self.base_layer.parameters():
            p.requires_grad_(False)

    @property
    def weight(self):
        return self.base_layer.weight

    def forward(self, x):
        from .kernels.fast_lora import LoRA_W
        from .kernels.utils import get_lora_parameters
        W, Wq, A, B, s = get_lora_parameters(self)
        return LoRA_W.apply(x, W, Wq, A, B, s)
