"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python benchmarks/sanitize_smoke.py
Shapes are tiny but exercise tails (rows not multiple of the tile, K tails, split-K, both operand
majors, CTA-pair and single-CTA GEMM kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200.nf4 import quantize_nf4  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
torch.manual_seed(0)


class N:
    weight = torch.ones(512, device=DEV, dtype=BF)
    variance_epsilon = 1e-5


X = torch.randn(3, 37, 512, device=DEV, dtype=BF, requires_grad=True)
Y = K.fast_rms_layernorm(N, X); Y.sum().backward() if False else Y.backward(torch.randn_like(Y))
q = torch.randn(2, 33, 8 * 64, device=DEV, dtype=BF); k = torch.randn(2, 33, 2 * 64, device=DEV, dtype=BF)
inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64)); fr = torch.outer(torch.arange(64).float(), inv)
emb = torch.cat([fr, fr], -1); cos, sin = emb.cos().to(DEV, BF), emb.sin().to(DEV, BF)
K.fast_rope_embedding(q.view(2, 33, 8, 64).transpose(1, 2), k.view(2, 33, 2, 64).transpose(1, 2), cos, sin)
K.fast_rope_embedding(q.view(2, 33, 8, 64).transpose(1, 2), k.view(2, 33, 2, 64).transpose(1, 2), cos.float(), sin.float(),
                      torch.randint(0, 64, (66,), device=DEV, dtype=torch.int32))
e = torch.randn(1, 50, 1432, device=DEV, dtype=BF); g = torch.randn_like(e)
h = K.swiglu_fg_kernel(e, g); K.swiglu_DWf_DW_dfg_kernel(h.view(50, -1), e.view(50, -1).clone(), g.view(50, -1).clone())
h = K.geglu_approx_forward_kernel(e, g)
logits = torch.randn(1, 9, 5000, device=DEV, dtype=BF, requires_grad=True)
labels = torch.randint(0, 5000, (1, 9), device=DEV); labels[0, 2] = -100
K.fast_cross_entropy_loss(logits * 1.0, labels, 30.0, 0).backward()
W = (torch.randn(320, 512, device=DEV) * 0.02).to(BF)
p, qs = quantize_nf4(W); D = K.fast_dequantize(p, qs)
for (a_mn, b_mn, M, Nn, Kk, bn, cg, sk) in [(0, 0, 300, 200, 136, 0, 1, 1), (0, 1, 304, 264, 200, 128, 2, 1),
                                            (1, 1, 520, 64, 1000, 64, 1, 4), (0, 0, 1024, 512, 256, 256, 2, 1),
                                            (1, 0, 264, 136, 72, 0, 0, 2)]:
    A = torch.randn(M, Kk, device=DEV).to(BF); B = torch.randn(Nn, Kk, device=DEV).to(BF)
    Aop = A.t().contiguous() if a_mn else A
    Bop = B.t().contiguous() if b_mn else B
    if a_mn and (M % 8):
        continue
    out = torch.empty(M, Nn, device=DEV, dtype=torch.float32)
    K.gemm(M, Nn, [(Aop, Bop, Kk)], out, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, cta_group=cg, split_k=sk)
    ref = A.float() @ B.float().t()
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-3
A_ = torch.rand(16, 512, device=DEV) - 0.5; B_ = torch.randn(320, 16, device=DEV) * 0.05
Xs = torch.randn(1, 70, 512, device=DEV, dtype=BF, requires_grad=True)
A_p, B_p = torch.nn.Parameter(A_), torch.nn.Parameter(B_)
out = K.LoRA_W.apply(Xs * 1.0, p, qs, A_p, B_p, 2.0); out.backward(torch.randn_like(out))
hid = torch.randn(1, 40, 512, device=DEV, dtype=BF, requires_grad=True)
Wlm = (torch.randn(3000, 512, device=DEV) * 0.05).to(BF)
K.unsloth_fused_ce_loss(None, hid * 1.0, Wlm, None, torch.randint(0, 3000, (1, 40), device=DEV), None, None, None,
                        chunk_rows=128).backward()
torch.cuda.synchronize()
print("sanitize smoke ok")
