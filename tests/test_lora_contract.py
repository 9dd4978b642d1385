"""Host-side PEFT attribute contract of `get_lora_parameters(_bias)` (a10: kernels/utils.py:335-440),
with the cases the reference's tests/test_fast_gemv_dispatch.py:36-63 pins (a 16-bit weight that
carries a `weight_scale` must NOT get a quant state) and the adapter on/off rules of
kernels/utils.py:365-397.  No device work: runs on CPU."""
from types import SimpleNamespace

import torch

from unsloth_b200.kernels.utils import QUANT_STATE, get_lora_parameters, get_lora_parameters_bias
from unsloth_b200.lora import LoraLinear


def _proj(weight, weight_scale=None):
    proj = SimpleNamespace(weight=weight, bias=None, merged=False)
    if weight_scale is not None:
        proj.weight_scale = weight_scale
    return proj


def test_bf16_weight_scale_not_used_as_quant_state():
    proj = _proj(torch.randn(4, 4, dtype=torch.bfloat16), torch.rand(2, 2))
    W, W_quant = get_lora_parameters_bias(proj)[:2]
    assert W_quant is None and W is proj.weight


def test_plain_bf16_has_no_quant_state():
    proj = _proj(torch.randn(4, 4, dtype=torch.bfloat16))
    assert get_lora_parameters_bias(proj)[1] is None
    assert get_lora_parameters(proj)[1:] == (None, None, None, None)       # no adapters on a bare layer
    assert QUANT_STATE(proj.weight) is None


def test_adapter_rules():
    base = torch.nn.Linear(16, 8, bias=True)
    lin = LoraLinear(base, r=4, lora_alpha=8, init_b_std=0.1)
    W, Wq, A, B, s = get_lora_parameters(lin)
    assert W is base.weight and Wq is None
    assert A is lin.lora_A["default"].weight and B is lin.lora_B["default"].weight and s == 2.0
    assert A.shape == (4, 16) and B.shape == (8, 4) and A.dtype == torch.float32
    assert get_lora_parameters_bias(lin)[5] is base.bias
    lin.disable_adapters = True                              # kernels/utils.py:365-371
    assert get_lora_parameters(lin)[2:] == (None, None, None)
    lin.disable_adapters, lin.merged = False, True
    assert get_lora_parameters(lin)[2:] == (None, None, None)
    lin.merged = False
    # a second adapter becomes the active one
    lin.lora_A["other"] = torch.nn.Linear(16, 4, bias=False)
    lin.lora_B["other"] = torch.nn.Linear(4, 8, bias=False)
    lin.scaling["other"] = 0.5
    lin.active_adapters = ["other"]
    _, _, A2, B2, s2 = get_lora_parameters(lin)
    assert A2 is lin.lora_A["other"].weight and B2 is lin.lora_B["other"].weight and s2 == 0.5


def test_quant_state_travels_on_the_weight():
    from unsloth_b200.nf4 import Params4bit, QuantState
    qs = QuantState(torch.zeros(4, dtype=torch.uint8), torch.Size((4, 64)), None, 64, "nf4", torch.bfloat16,
                    torch.tensor(0.0), QuantState(torch.ones(1), code=torch.zeros(256), blocksize=256))
    w = Params4bit(torch.zeros(128, 1, dtype=torch.uint8), qs)
    layer = SimpleNamespace(weight=w, bias=None, merged=False)
    assert get_lora_parameters_bias(layer)[1] is qs and QUANT_STATE(w) is qs
    assert not w.requires_grad
