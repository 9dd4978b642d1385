"""Tiny ncu target: a handful of launches of the dominant kernel at the cfg2 gate/up projection
shape (M=8192 tokens, N=14336, K=4096 + one 64-wide LoRA rank block)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unsloth_b200.kernels as K  # noqa: E402

M, N, Kk = 8192, 14336, 4096
A = torch.randn(M, Kk, device="cuda", dtype=torch.bfloat16)
B = (torch.randn(N, Kk, device="cuda") * 0.02).to(torch.bfloat16)
XA = torch.randn(M, 64, device="cuda", dtype=torch.bfloat16)
Bp = torch.randn(N, 64, device="cuda", dtype=torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    K.gemm(M, N, [(A, B, Kk), (XA, Bp, 64)], C)
torch.cuda.synchronize()
print("ok")
