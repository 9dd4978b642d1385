"""GPU tests of the tcgen05 attention (csrc/attention.cu) against the explicit masked softmax in
fp32 (the definition), and against the library kernels the reference would call (flash-attn 2) on
identical tensors: causal, GQA, sliding window, soft-capping, packed rows, strided projection-buffer
views, D = 64 / 128 / 256, ragged lengths."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16


def ref_attention(Q, K, V, scale, window_left=-1, softcap=0.0, lengths=None):
    """The oracle's fp32 definition (oracle/restate.py::attention), evaluated on the device."""
    from oracle import restate as R
    return R.attention(Q, K, V, scale, window_left, softcap, lengths)


def _mk(B, S, Hq, Hk, D, seed=0, strided=True):
    torch.manual_seed(seed)
    if strided:      # views of [B,S,H*D] projection buffers, as LlamaAttention_fast_forward passes them
        Q = torch.randn(B, S, Hq * D, device=DEV).to(BF).view(B, S, Hq, D)
        K = torch.randn(B, S, Hk * D, device=DEV).to(BF).view(B, S, Hk, D)
        V = torch.randn(B, S, Hk * D, device=DEV).to(BF).view(B, S, Hk, D)
    else:
        Q, K, V = (torch.randn(B, S, h, D, device=DEV).to(BF) for h in (Hq, Hk, Hk))
    return Q, K, V


CASES = [
    # B, S, Hq, Hk, D, window_left, softcap
    (2, 256, 4, 2, 128, -1, 0.0),
    (1, 1000, 8, 2, 128, -1, 0.0),          # ragged length (not a multiple of the tile)
    (2, 512, 4, 4, 64, -1, 0.0),
    (1, 640, 4, 2, 128, 200, 0.0),          # sliding window (Mistral)
    (1, 384, 4, 2, 256, -1, 50.0),          # Gemma-2: D = 256, soft-capping
    (1, 700, 4, 2, 256, 256, 50.0),         # Gemma-2 even layer: window + softcap
    (1, 2048, 8, 2, 128, -1, 0.0),
]


@pytest.mark.parametrize("B,S,Hq,Hk,D,wl,cap", CASES)
def test_attention_forward_vs_definition(B, S, Hq, Hk, D, wl, cap):
    from unsloth_b200.kernels.attention import attention_forward
    Q, K, V = _mk(B, S, Hq, Hk, D, seed=S + D)
    scale = D ** -0.5
    O, lse = attention_forward(Q, K, V, scale, wl, cap)
    Or, lser = ref_attention(Q, K, V, scale, wl, cap)
    err = (O.float() - Or).abs().max().item()
    assert err <= 2e-2 * Or.abs().max().item() + 1e-3, ("O", err)
    assert (lse - lser).abs().max().item() <= 2e-2, ("lse", (lse - lser).abs().max().item())
    # the library kernel on the same tensors (what the reference would run) is not closer to the definition
    from flash_attn import flash_attn_func
    Of = flash_attn_func(Q, K, V, softmax_scale=scale, causal=True, window_size=(wl, wl) if wl >= 0 else (-1, -1),
                         softcap=cap)
    err_f = (Of.float() - Or).abs().max().item()
    assert err <= 2.0 * err_f + 2e-3, (err, err_f)


@pytest.mark.parametrize("lengths,Hq,Hk,D,wl,cap", [([300, 200, 524], 4, 2, 128, -1, 0.0),
                                                    ([130, 894], 4, 2, 128, 100, 0.0),
                                                    ([77, 300, 135], 4, 2, 256, -1, 50.0)])
def test_attention_forward_packed_rows(lengths, Hq, Hk, D, wl, cap):
    from unsloth_b200.kernels.attention import attention_forward
    T_ = sum(lengths)
    Q, K, V = _mk(1, T_, Hq, Hk, D, seed=T_)
    cu = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0)), dtype=torch.int32, device=DEV)
    seq_info = (torch.tensor(lengths, dtype=torch.int32, device=DEV), cu, max(lengths))
    O, lse = attention_forward(Q, K, V, D ** -0.5, wl, cap, seq_info)
    Or, lser = ref_attention(Q, K, V, D ** -0.5, wl, cap, lengths)
    assert (O.float() - Or).abs().max().item() <= 2e-2 * Or.abs().max().item() + 1e-3
    assert (lse - lser).abs().max().item() <= 2e-2


@pytest.mark.parametrize("B,S,Hq,Hk,D,wl,cap", [c for c in CASES if c[1] <= 1000])
def test_attention_backward_vs_definition(B, S, Hq, Hk, D, wl, cap):
    from unsloth_b200.kernels.attention import fast_attention
    Q, K, V = _mk(B, S, Hq, Hk, D, seed=S + D + 1, strided=False)
    dO = (torch.randn(B, S, Hq, D, device=DEV) * 0.1).to(BF)
    q, k, v = (t.clone().requires_grad_() for t in (Q, K, V))
    O = fast_attention(q, k, v, D ** -0.5, (wl, wl), cap)
    O.backward(dO)
    qr, kr, vr = (t.float().clone().requires_grad_() for t in (Q, K, V))
    Or, _ = ref_attention(qr, kr, vr, D ** -0.5, wl, cap)
    Or.backward(dO.float())
    from flash_attn import flash_attn_func
    qf, kf, vf = (t.clone().requires_grad_() for t in (Q, K, V))
    flash_attn_func(qf, kf, vf, softmax_scale=D ** -0.5, causal=True,
                    window_size=(wl, wl) if wl >= 0 else (-1, -1), softcap=cap).backward(dO)
    for n, a, b, f in (("dQ", q.grad, qr.grad, qf.grad), ("dK", k.grad, kr.grad, kf.grad), ("dV", v.grad, vr.grad, vf.grad)):
        err = (a.float() - b).abs().max().item()
        err_f = (f.float() - b).abs().max().item()
        assert err <= 2.0 * err_f + 2e-2 * b.abs().max().item(), (n, err, err_f, b.abs().max().item())


def test_attention_cfg2_properties_and_determinism():
    """cfg2 size (B=4, S=2048, 32/8 heads, D=128): rows of P sum to one (V = ones -> O = ones), run to
    run bitwise equality, and agreement with cuDNN / flash-attn on sampled heads."""
    from unsloth_b200.kernels.attention import attention_forward
    Q, K, V = _mk(4, 2048, 32, 8, 128, seed=5)
    ones = torch.ones_like(V)
    O1, _ = attention_forward(Q, K, ones, 128 ** -0.5)
    assert (O1.float() - 1).abs().max().item() <= 8e-3
    Oa, la = attention_forward(Q, K, V, 128 ** -0.5)
    Ob, lb = attention_forward(Q, K, V, 128 ** -0.5)
    assert torch.equal(Oa, Ob) and torch.equal(la, lb)
    from flash_attn import flash_attn_func
    Of = flash_attn_func(Q, K, V, softmax_scale=128 ** -0.5, causal=True)
    assert (Oa.float() - Of.float()).abs().max().item() <= 2e-2
