"""ncu targets: `python benchmarks/ncu_targets.py <name>` launches a few instances of one kernel
at cfg2 sizes (names: gemm2, skinny, dequant, rms, rope, ce, gemv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200 import _lib as L  # noqa: E402

name = sys.argv[1]
DEV, BF = "cuda", torch.bfloat16
T, H, I = 8192, 4096, 14336
if name == "gemm2":
    A = torch.randn(T, H, device=DEV, dtype=BF); B = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    XA = torch.randn(T, 64, device=DEV, dtype=BF); Bp = torch.randn(I, 64, device=DEV, dtype=BF)
    C = torch.empty(T, I, device=DEV, dtype=BF)
    for _ in range(4):
        K.gemm(T, I, [(A, B, H), (XA, Bp, 64)], C)
elif name == "skinny":
    X = torch.randn(T, H, device=DEV, dtype=BF); A64 = torch.randn(64, H, device=DEV, dtype=BF)
    XA = torch.empty(T, 64, device=DEV, dtype=BF); G = torch.randn(T, 64, device=DEV, dtype=BF)
    o = torch.empty(H, 64, device=DEV, dtype=torch.float32)
    for _ in range(3):
        K.gemm(T, 64, [(X, A64, H)], XA)
        K.gemm(H, 64, [(X, G, T)], o, a_mn=True, b_mn=True, split_k=4)
elif name == "dequant":
    from unsloth_b200.nf4 import quantize_nf4
    W = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    p, q = quantize_nf4(W); out = torch.empty_like(W)
    for _ in range(4):
        K.fast_dequantize(p, q, out=out)
elif name == "gemv":
    from unsloth_b200.nf4 import quantize_nf4
    W = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    p, q = quantize_nf4(W)
    x = torch.randn(1, 1, H, device=DEV, dtype=BF); out = torch.empty(1, 1, I, device=DEV, dtype=BF)
    for _ in range(3):
        K.fast_gemv(x, p, q, out=out)
elif name == "rms":
    X = torch.randn(T, H, device=DEV, dtype=BF); W = torch.ones(H, device=DEV, dtype=BF)
    Y = torch.empty_like(X); r = torch.empty(T, device=DEV); dY = torch.randn(T, H, device=DEV, dtype=BF)
    for _ in range(3):
        L.call("ub200_rms_layernorm_fwd", L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(Y), H, L.ptr(r), T, H, 1e-5, 0, L.BF16, L.stream())
        L.call("ub200_rms_layernorm_bwd", L.ptr(dY), H, L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(r), L.ptr(dY), H, T, H, 0, L.BF16, L.stream())
elif name == "rope":
    from unsloth_b200.kernels.rope_embedding import _launch
    B, S, Hq, Hk, D = 4, 2048, 32, 8, 128
    q = torch.randn(B, S, Hq * D, device=DEV, dtype=BF); k = torch.randn(B, S, Hk * D, device=DEV, dtype=BF)
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D)); fr = torch.outer(torch.arange(S).float(), inv)
    emb = torch.cat([fr, fr], -1); cos, sin = emb.cos().to(DEV, BF), emb.sin().to(DEV, BF)
    for _ in range(4):
        _launch(q.view(B, S, Hq, D).transpose(1, 2), k.view(B, S, Hk, D).transpose(1, 2), cos, sin, None, False, True)
elif name == "glu":
    e = torch.randn(T, I, device=DEV, dtype=BF); g = torch.randn(T, I, device=DEV, dtype=BF); h = torch.empty_like(e)
    for _ in range(3):
        L.call("ub200_glu_fwd", 0, L.ptr(e), L.ptr(g), L.ptr(h), e.numel(), L.BF16, L.stream())
        L.call("ub200_glu_bwd", 0, L.ptr(h), L.ptr(e), L.ptr(g), e.numel(), L.BF16, L.stream())
elif name == "ce":
    from unsloth_b200.kernels.cross_entropy_loss import _ce_backward_, _ce_forward
    V = 128256
    logits = torch.randn(2048, V, device=DEV, dtype=BF); labels = torch.randint(0, V, (2048,), device=DEV)
    dl = torch.full((1,), 1e-3, device=DEV)
    for _ in range(3):
        l, lse = _ce_forward(logits, labels, 0.0, 0.0)
        _ce_backward_(logits, lse, labels, dl, 0, 0.0, 0.0)
elif name in ("attn_fwd", "attn_bwd"):
    from unsloth_b200.kernels.attention import attention_backward, attention_forward
    B, S, Hq, Hk, D = 4, 2048, 32, 8, 128
    q = torch.randn(B, S, Hq * D, device=DEV, dtype=BF).view(B, S, Hq, D)
    k = torch.randn(B, S, Hk * D, device=DEV, dtype=BF).view(B, S, Hk, D)
    v = torch.randn(B, S, Hk * D, device=DEV, dtype=BF).view(B, S, Hk, D)
    for _ in range(2):
        O, lse = attention_forward(q, k, v, D ** -0.5)
    if name == "attn_bwd":
        dO = torch.randn_like(O)
        for _ in range(2):
            attention_backward(dO, q, k, v, O, lse, D ** -0.5)
elif name == "grouped":
    from unsloth_b200.nf4 import quantize_nf4
    os.environ["UB200_GROUPED_BWD"] = "1"
    r = 16
    X = torch.randn(4, 2048, H, device=DEV).to(BF)

    def mk(o, i):
        p_, q_ = quantize_nf4((torch.randn(o, i, device=DEV) * 0.02).to(BF))
        return p_, q_, torch.nn.Parameter((torch.rand(r, i, device=DEV) * 2 - 1) / i ** 0.5), \
            torch.nn.Parameter(torch.randn(o, r, device=DEV) * 0.02)
    qp, kp, vp = mk(H, H), mk(1024, H), mk(1024, H)
    dY = (torch.randn(4, 2048, H, device=DEV) * 0.1).to(BF)
    for _ in range(2):
        x = X.clone().requires_grad_()
        o = K.LoRA_QKV.apply(x, qp[0], qp[1], qp[2], qp[3], 1.0, kp[0], kp[1], kp[2], kp[3], 1.0,
                             vp[0], vp[1], vp[2], vp[3], 1.0, True)
        torch.autograd.backward(list(o), [dY, dY[..., :1024].contiguous(), dY[..., :1024].contiguous()])
torch.cuda.synchronize()
print("ok")
