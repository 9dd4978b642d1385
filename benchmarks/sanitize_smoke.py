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
# ---- round 2: grouped launches (in-kernel dependencies, split-K last-arriver reduction), LoRA_QKV / LoRA_MLP through
# every schedule, fused-dequant GEMM, attention forward / backward (window, softcap, packed rows, D = 64 / 128 / 256)
from unsloth_b200.kernels.utils import gemm_nf4  # noqa: E402
for mode in (("1", "1"), ("1", "2"), ("0", "0")):
    os.environ["UB200_GROUPED"] = "1" if mode[0] == "1" else "0"
    os.environ["UB200_GROUPED_FWD"], os.environ["UB200_GROUPED_BWD"] = mode
    def mk(o, i):
        p_, q_ = quantize_nf4((torch.randn(o, i, device=DEV) * 0.02).to(BF))
        return p_, q_, torch.nn.Parameter((torch.rand(16, i, device=DEV) - 0.5) / 8), torch.nn.Parameter(torch.randn(o, 16, device=DEV) * 0.05)
    qp, kp, vp = mk(512, 512), mk(128, 512), mk(128, 512)
    x = torch.randn(2, 150, 512, device=DEV, dtype=BF, requires_grad=True)
    Qo, Ko, Vo = K.LoRA_QKV.apply(x * 1, qp[0], qp[1], qp[2], qp[3], 1.0, kp[0], kp[1], kp[2], kp[3], 1.0,
                                  vp[0], vp[1], vp[2], vp[3], 1.0, True)
    torch.autograd.backward([Qo, Ko, Vo], [torch.randn_like(Qo), torch.randn_like(Ko), torch.randn_like(Vo)])
    gt, up, dn = mk(1408, 512), mk(1408, 512), mk(512, 1408)
    o = K.LoRA_MLP.apply(x * 1, gt[0], gt[1], gt[2], gt[3], 1.0, up[0], up[1], up[2], up[3], 1.0, dn[0], dn[1], dn[2], dn[3], 1.0,
                         K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel, True)
    o.backward(torch.randn_like(o))
for k_ in ("UB200_GROUPED", "UB200_GROUPED_FWD", "UB200_GROUPED_BWD"):
    os.environ.pop(k_, None)
Xn = torch.randn(300, 512, device=DEV, dtype=BF)
gemm_nf4(Xn, p, qs, torch.empty(300, 320, device=DEV, dtype=BF))
for (Bq, Sq, Hq, Hk, Dh, wl, cap, lengths) in [(2, 200, 4, 2, 128, -1, 0.0, None), (1, 333, 4, 2, 64, 100, 0.0, None),
                                               (1, 260, 2, 1, 256, 64, 30.0, None), (1, 300, 4, 2, 128, -1, 0.0, [130, 170])]:
    qa, ka, va = (torch.randn(Bq, Sq, h_, Dh, device=DEV, dtype=BF, requires_grad=True) for h_ in (Hq, Hk, Hk))
    si = None
    if lengths:
        cu = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0)), dtype=torch.int32, device=DEV)
        si = (torch.tensor(lengths, dtype=torch.int32, device=DEV), cu, max(lengths))
    Oa = K.fast_attention(qa, ka, va, Dh ** -0.5, (wl, wl), cap, si)
    Oa.backward(torch.randn_like(Oa))
torch.cuda.synchronize()
print("sanitize smoke ok")
