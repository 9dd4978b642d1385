"""compute-sanitizer target for the GEMM with the gated-activation epilogue (ub200_gemm_glu): every kernel
instantiation the host can pick (CTA pair / single CTA, tile widths 256 / 128 / 64), both modes, a row tail
(M not a multiple of the tile), two K segments with a K tail.
    compute-sanitizer --tool memcheck python benchmarks/sanitize_glu.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsloth_b200 import _lib as L  # noqa: E402
from unsloth_b200.kernels.utils import gemm_glu  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
torch.manual_seed(0)
for (M, N, K, bn, cg, b_mn, act, dt) in [(300, 512, 136, 256, 2, True, 0, BF), (300, 256, 136, 128, 2, False, 1, torch.float16),
                                         (200, 512, 72, 256, 1, True, 2, BF), (136, 128, 200, 128, 1, False, 0, BF),
                                         (136, 64, 72, 64, 1, False, 0, BF), (600, 768, 264, 0, 0, True, 0, BF)]:
    A = torch.randn(M, K, device=DEV).to(dt)
    B = (torch.randn(N, K, device=DEV) * 0.1).to(dt)
    A2 = torch.randn(M, 64, device=DEV).to(dt)
    B2 = (torch.randn(N, 64, device=DEV) * 0.1).to(dt)
    segs = [(A, B.t().contiguous() if b_mn else B, K), (A2, B2.t().contiguous() if b_mn else B2, 64)]
    for mode in (L.GLU_EPI_FWD, L.GLU_EPI_BWD):
        e = torch.randn(M, N, device=DEV).to(dt)
        g = torch.randn(M, N, device=DEV).to(dt)
        out = torch.empty(M, N, device=DEV, dtype=dt)
        gemm_glu(mode, act, M, N, segs, out, e, g, b_mn=b_mn, block_n=bn, cta_group=cg)
torch.cuda.synchronize()
print("ok")
