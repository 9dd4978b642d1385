"""ncu targets: `python benchmarks/ncu_targets.py <name>` launches a few instances of one kernel
at cfg2 sizes (names: gemm2, skinny, dequant, rms, rope, ce)."""
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
elif name == "rms":
    X = torch.randn(T, H, device=DEV, dtype=BF); W = torch.ones(H, device=DEV, dtype=BF)
    Y = torch.empty_like(X); r = torch.empty(T, device=DEV); dY = torch.randn(T, H, device=DEV, dtype=BF)
    for _ in range(3):
        L.call("ub200_rms_layernorm_fwd", L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(Y), H, L.ptr(r), T, H, 1e-5, 0, L.BF16, L.stream())
        L.call("ub200_rms_layernorm_bwd", L.ptr(dY), H, L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(r), L.ptr(dY), H, T, H, 0, L.BF16, L.stream())
torch.cuda.synchronize()
print("ok")
