"""ncu target: the GEMM with the gated-activation epilogue at the cfg2 shape (8192 x 14336 x 4096+64, bf16).
usage: python benchmarks/ncu_glu_target.py bwd|fwd"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsloth_b200 import _lib as L  # noqa: E402
from unsloth_b200.kernels import utils as KU  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
T_, H, I = 8192, 4096, 14336
mode = sys.argv[1] if len(sys.argv) > 1 else "bwd"
torch.manual_seed(0)
e = torch.randn(T_, I, device=DEV).to(BF)
g = torch.randn(T_, I, device=DEV).to(BF)
out = torch.empty(T_, I, device=DEV, dtype=BF)
if mode == "bwd":
    A = (torch.randn(T_, H, device=DEV) * 0.1).to(BF)
    B = (torch.randn(H, I, device=DEV) * 0.02).to(BF)
    A2 = torch.randn(T_, 64, device=DEV).to(BF)
    B2 = (torch.randn(64, I, device=DEV) * 0.01).to(BF)
    for _ in range(3):
        KU.gemm_glu(L.GLU_EPI_BWD, 0, T_, I, [(A, B, H), (A2, B2, 64)], out, e, g, b_mn=True)
else:
    A = torch.randn(T_, H, device=DEV).to(BF)
    B = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    A2 = torch.randn(T_, 64, device=DEV).to(BF)
    B2 = (torch.randn(I, 64, device=DEV) * 0.01).to(BF)
    for _ in range(3):
        KU.gemm_glu(L.GLU_EPI_FWD, 0, T_, I, [(A, B, H), (A2, B2, 64)], out, e, g)
torch.cuda.synchronize()
