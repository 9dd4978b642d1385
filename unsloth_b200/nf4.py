"""NF4 double-quantised weight container for the QLoRA path.

bitsandbytes (the reference's 4-bit backend, pyproject.toml:473) is not available offline, so
this module provides the minimal objects the reference's hot path reads
(unsloth/kernels/utils.py:582-598): a packed uint8 weight `[numel/2, 1]` carrying a
`.quant_state` with `absmax` (uint8), `shape`, `dtype`, `blocksize`, `offset` and
`state2.{absmax, code, blocksize}` -- field for field the layout of
`bitsandbytes.functional.QuantState` for `bnb_4bit_quant_type="nf4"`,
`bnb_4bit_use_double_quant=True` (models/llama.py:2620-2626).
"""
from __future__ import annotations

import torch

from . import _lib as L


class QuantState:
    """Field-compatible stand-in for bitsandbytes.functional.QuantState (nested statistics)."""

    def __init__(self, absmax, shape=None, code=None, blocksize=None, quant_type=None, dtype=None,
                 offset=None, state2=None):
        self.absmax = absmax
        self.shape = shape
        self.code = code
        self.dtype = dtype
        self.blocksize = blocksize
        self.quant_type = quant_type
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None

    def to(self, device):
        self.absmax = self.absmax.to(device)
        if self.code is not None:
            self.code = self.code.to(device)
        if self.nested:
            self.offset = self.offset.to(device)
            self.state2.absmax = self.state2.absmax.to(device)
            self.state2.code = self.state2.code.to(device)
        return self


def create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """The 8-bit dynamic code of the second-level quantiser (bitsandbytes' published
    `create_dynamic_map`): 256 sorted values in [-1, 1]."""
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 2 ** total_bits
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


_CODE2 = {}


def _code2(device):
    key = str(device)
    if key not in _CODE2:
        _CODE2[key] = create_dynamic_map().to(device)
    return _CODE2[key]


def quantize_nf4(W: torch.Tensor, blocksize: int = 64, blocksize2: int = 256):
    """NF4 + double quantisation of a CUDA weight.  Returns (packed uint8 [numel/2, 1],
    QuantState).  Stage 1 (per-64 absmax, nearest code, two codes per byte with the first
    element in the HIGH nibble) is ub200_quantize_nf4; stage 2 quantises the mean-centred
    absmax to 8 bits with the dynamic map in blocks of 256 (tiny; torch ops)."""
    L.require_cuda(W)
    assert blocksize == 64 and W.numel() % 64 == 0
    Wc = W.contiguous()
    n = Wc.numel()
    packed = torch.empty((n // 2, 1), dtype=torch.uint8, device=W.device)
    absmax = torch.empty(n // 64, dtype=torch.float32, device=W.device)
    L.call("ub200_quantize_nf4", L.ptr(Wc), L.dt(Wc), L.ptr(packed), L.ptr(absmax), n, blocksize,
           L.stream())
    offset = absmax.mean()
    am = absmax - offset
    nb = am.numel()
    pad = (-nb) % blocksize2
    amp = torch.cat([am, am.new_zeros(pad)]).reshape(-1, blocksize2)
    absmax2 = amp.abs().amax(dim=1)
    sc = (amp / absmax2.clamp_min(1e-30)[:, None]).reshape(-1)
    code2 = _code2(W.device)
    mid = (code2[:-1] + code2[1:]) / 2
    q = torch.searchsorted(mid, sc.contiguous()).to(torch.uint8)[:nb].contiguous()
    state2 = QuantState(absmax=absmax2.contiguous(), code=code2, blocksize=blocksize2,
                        dtype=torch.float32)
    qs = QuantState(absmax=q, shape=torch.Size(W.shape), dtype=W.dtype, blocksize=blocksize,
                    quant_type="nf4", offset=offset, state2=state2)
    return packed, qs


class Params4bit(torch.nn.Parameter):
    """Frozen packed weight carrying `.quant_state` (what `get_lora_parameters` reads through
    `base_layer.weight.quant_state`, unsloth/kernels/utils.py:352-356)."""

    def __new__(cls, data, quant_state=None):
        self = torch.Tensor._make_subclass(cls, data, False)
        self.quant_state = quant_state
        return self


class Linear4bit(torch.nn.Module):
    """Bias-free NF4 linear: the `base_layer` of a QLoRA projection."""

    def __init__(self, in_features, out_features, packed, quant_state):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = Params4bit(packed, quant_state)
        self.bias = None

    @classmethod
    def from_dense(cls, W: torch.Tensor):
        packed, qs = quantize_nf4(W)
        return cls(W.shape[1], W.shape[0], packed, qs)

    def forward(self, x):
        from .kernels.utils import matmul_lora
        return matmul_lora(x, self.weight, self.weight.quant_state, None, None, None)
