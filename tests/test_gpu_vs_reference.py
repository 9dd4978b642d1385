"""bf16 parity against the REFERENCE ITSELF on the B200:
the reference's own Triton kernels and LoRA autograd functions, compiled natively by Triton for
the device (oracle/ref_shim.load_reference_kernels_native -- the unmodified reference install under
baseline/_ref), run on IDENTICAL tensors at BASELINE.json cfg2 sizes (Llama-3-8B: T = 4 x 2048,
H = 4096, I = 14336, V = 128256; plus Gemma-2 / Mistral variants), next to the CUDA path called
through the C ABI.

Every comparison records (and prints) the achieved numbers -- fraction of elements that differ at
all, fraction outside rtol 1e-3 (+ atol 1e-3 max|ref|), max distance in bf16 ulps, max abs error --
into gpurun_out/parity_vs_reference.jsonl so the gates below can be tightened from evidence
(summaries are committed under profiles/).

Gates.  Bandwidth kernels (RMSNorm, RoPE, GLU, CE): north_star's rtol 1e-3; since one bf16 ulp is
3.9e-3 relative, this is "bit-identical except for a bounded fraction of 1-ulp roundings"
(SURVEY section 9): <= 0.2 % of the elements may sit outside and none further than 2 ulps / 1.6e-2 max.
GEMM-based functions (LoRA_MLP / LoRA_QKV / LoRA_W, fused CE): the reference rounds to bf16 after
every term while the fused kernels round once, so both are measured against an fp32 truth and
ours must not be further from it than 1.25 x the reference's own error (+ 2e-3 of the scale).
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF = torch.bfloat16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_vs_reference.jsonl")


@pytest.fixture(scope="module")
def refk():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference not installed (baseline/_ref absent): see DESIGN.md section 5")
    torch.backends.cuda.matmul.allow_tf32 = False
    return ref_shim.load_reference_kernels_native()


def _ordered(t):
    """bf16/fp16 bit patterns mapped to integers that are monotone in the value."""
    i = t.contiguous().view(torch.int16).to(torch.int32)
    return torch.where(i < 0, -(i & 0x7FFF), i)


def record(name, ours, ref, rtol=1e-3):
    o, r = ours.detach(), ref.detach()
    assert o.shape == r.shape, (name, o.shape, r.shape)
    of, rf = o.float(), r.float()
    scale = rf.abs().max().item()
    diff = (of - rf).abs()
    outside = (diff > (1e-3 * scale + rtol * rf.abs())).float().mean().item()
    rec = {"name": name, "numel": o.numel(), "differ_frac": (of != rf).float().mean().item(),
           "outside_rtol1e-3_frac": outside, "max_abs": diff.max().item(), "scale": scale}
    if o.dtype in (torch.bfloat16, torch.float16) and r.dtype == o.dtype:
        rec["max_ulp"] = int((_ordered(o) - _ordered(r)).abs().max().item())
    print("PARITY", json.dumps(rec))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    return rec


def gate_elementwise(name, ours, ref, max_outside=0.002, max_ulp=2):
    rec = record(name, ours, ref)
    assert rec["outside_rtol1e-3_frac"] <= max_outside, rec
    assert rec["max_abs"] <= 1.6e-2 * rec["scale"] + 1e-6, rec
    if "max_ulp" in rec and max_ulp is not None:
        # ulp distance is only meaningful away from zero: re-measure on |ref| >= 1e-2 scale
        m = ref.detach().float().abs() >= 1e-2 * rec["scale"]
        if m.any():
            u = (_ordered(ours.detach())[m] - _ordered(ref.detach())[m]).abs().max().item()
            assert u <= max_ulp, (rec, u)
    return rec


def gate_gemm(name, ours, ref, truth):
    rec = record(name, ours, ref)
    t = truth.detach().double()
    e_ours = (ours.detach().double() - t).abs().max().item()
    e_ref = (ref.detach().double() - t).abs().max().item()
    scale = t.abs().max().item()
    rec2 = {"name": name + ":vs_fp32_truth", "err_ours": e_ours, "err_reference": e_ref, "scale": scale,
            "rel_ours": e_ours / scale, "rel_reference": e_ref / scale}
    print("PARITY", json.dumps(rec2))
    try:
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec2) + "\n")
    except OSError:
        pass
    assert e_ours <= 1.25 * e_ref + 2e-3 * scale, rec2


class Norm:
    def __init__(self, w, eps):
        self.weight, self.variance_epsilon = w, eps


# -------------------------------------------------------------------------------------------------
# a1 RMSNorm
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gemma,H,eps", [(False, 4096, 1e-5), (True, 3584, 1e-6)])
def test_rmsnorm_vs_reference(refk, gemma, H, eps):
    from unsloth_b200.kernels import fast_rms_layernorm
    torch.manual_seed(3407)
    X = torch.randn(4, 2048, H, device=DEV).to(BF)
    W = (torch.randn(H, device=DEV) * 0.3 + (0 if gemma else 1)).to(BF)
    dY = torch.randn(4, 2048, H, device=DEV).to(BF)
    Xr = X.clone().requires_grad_()
    Yr = refk.rms_layernorm.fast_rms_layernorm(Norm(W, eps), Xr, gemma=gemma)
    Yr_c = Yr.detach().clone()
    Yr.backward(dY.clone())
    Xo = X.clone().requires_grad_()
    Yo = fast_rms_layernorm(Norm(W, eps), Xo, gemma=gemma)
    Yo_c = Yo.detach().clone()
    Yo.backward(dY.clone())
    tag = "rms_%s_H%d" % ("gemma" if gemma else "llama", H)
    gate_elementwise(tag + ":Y", Yo_c, Yr_c)
    gate_elementwise(tag + ":dX", Xo.grad, Xr.grad)


# -------------------------------------------------------------------------------------------------
# a2 RoPE (both call forms, in place on transposed views of the projection buffers)
# -------------------------------------------------------------------------------------------------
def _tables(S, D, base, dtype):
    inv = 1.0 / (base ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
    emb = torch.cat((fr, fr), -1)
    return emb.cos().to(DEV, dtype), emb.sin().to(DEV, dtype)


@pytest.mark.parametrize("indexed", [False, True])
@pytest.mark.parametrize("B,S,Hq,Hk,D,tdt,adt", [(4, 2048, 32, 8, 128, BF, BF), (1, 4096, 16, 8, 256, torch.float32, BF),
                                                  (2, 1024, 32, 8, 128, torch.float16, torch.float16)])
def test_rope_vs_reference(refk, indexed, B, S, Hq, Hk, D, tdt, adt):
    from unsloth_b200.kernels import fast_rope_embedding
    torch.manual_seed(7)
    cos, sin = _tables(S, D, 500000.0, tdt)
    q0 = torch.randn(B, S, Hq * D, device=DEV).to(adt)
    k0 = torch.randn(B, S, Hk * D, device=DEV).to(adt)
    dq = torch.randn(B, Hq, S, D, device=DEV).to(adt)
    dk = torch.randn(B, Hk, S, D, device=DEV).to(adt)
    idx = None
    if indexed:        # packed-style position ids: two documents per row
        pos = torch.cat([torch.arange(S // 3), torch.arange(S - S // 3)]).repeat(B)
        idx = pos.to(DEV, torch.int32)

    def run(fn):
        qb, kb = q0.clone().requires_grad_(), k0.clone().requires_grad_()
        Q = (qb * 1).view(B, S, Hq, D).transpose(1, 2)
        K = (kb * 1).view(B, S, Hk, D).transpose(1, 2)
        Qo, Ko = fn(Q, K, cos, sin, idx)
        Qc, Kc = Qo.detach().clone(), Ko.detach().clone()
        torch.autograd.backward([Qo, Ko], [dq.clone(), dk.clone()])
        return Qc, Kc, qb.grad, kb.grad
    r = run(refk.rope_embedding.fast_rope_embedding)
    o = run(fast_rope_embedding)
    tag = "rope_%s_D%d_%s" % ("indexed" if indexed else "noindex", D, str(adt).split(".")[-1])
    for n, a, b in zip(("Q", "K", "dQ", "dK"), o, r):
        gate_elementwise(tag + ":" + n, a.contiguous(), b.contiguous())


# -------------------------------------------------------------------------------------------------
# a11 SwiGLU / GEGLU
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["swiglu", "geglu_approx", "geglu_exact"])
def test_glu_vs_reference(refk, act):
    import unsloth_b200.kernels as K
    torch.manual_seed(11)
    T_, I = 8192, 14336
    e = torch.randn(1, T_, I, device=DEV).to(BF)
    g = torch.randn(1, T_, I, device=DEV).to(BF)
    DW = torch.randn(T_, I, device=DEV).to(BF)
    rf, rb = {"swiglu": (refk.swiglu.swiglu_fg_kernel, refk.swiglu.swiglu_DWf_DW_dfg_kernel),
              "geglu_approx": (refk.geglu.geglu_approx_forward_kernel, refk.geglu.geglu_approx_backward_kernel),
              "geglu_exact": (refk.geglu.geglu_exact_forward_kernel, refk.geglu.geglu_exact_backward_kernel)}[act]
    of, ob = {"swiglu": (K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel),
              "geglu_approx": (K.geglu_approx_forward_kernel, K.geglu_approx_backward_kernel),
              "geglu_exact": (K.geglu_exact_forward_kernel, K.geglu_exact_backward_kernel)}[act]
    hr = rf(e, g)
    ho = of(e, g)
    # the tanh / erf tails are where libdevice and our MUFU paths may differ by an ulp
    gate_elementwise(act + ":h", ho, hr, max_ulp=3)
    r3 = rb(DW.clone(), e.view(T_, I).clone(), g.view(T_, I).clone())
    o3 = ob(DW.clone(), e.view(T_, I).clone(), g.view(T_, I).clone())
    for n, a, b in zip(("h", "df", "de"), o3, r3):
        gate_elementwise(act + ":bwd_" + n, a, b, max_ulp=4)


# -------------------------------------------------------------------------------------------------
# a3 cross entropy on materialised logits
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,rows,softcap,scale", [(128256, 2048, 0.0, 0.0), (32768, 4096, 0.0, 0.0),
                                                   (256000, 1024, 30.0, 0.0), (128256, 512, 0.0, 0.5)])
def test_cross_entropy_vs_reference(refk, V, rows, softcap, scale):
    from unsloth_b200.kernels.cross_entropy_loss import Fast_CrossEntropyLoss
    torch.manual_seed(13)
    logits = (torch.randn(rows, V, device=DEV) * 2).to(BF)
    labels = torch.randint(0, V, (rows,), device=DEV)
    labels[::17] = -100
    dl = torch.rand(rows, device=DEV) / rows

    def run(fn):
        lg = logits.clone().requires_grad_()
        loss = fn(lg * 1, labels, softcap, scale)
        lc = loss.detach().clone()
        (lg_grad,) = torch.autograd.grad(loss, lg, dl.clone())
        return lc, lg_grad
    lr, gr = run(refk.cross_entropy_loss.Fast_CrossEntropyLoss.apply)
    lo, go = run(Fast_CrossEntropyLoss.apply)
    tag = "ce_V%d_cap%g_scale%g" % (V, softcap, scale)
    rec = record(tag + ":loss", lo, lr)
    assert rec["max_abs"] <= 1e-3 * rec["scale"] + 1e-4, rec        # fp32 losses
    gate_elementwise(tag + ":dlogits", go, gr, max_ulp=None)


# -------------------------------------------------------------------------------------------------
# a5-a8 LoRA_MLP / LoRA_QKV / LoRA_W (16-bit and NF4 base weights) at cfg2 sizes
# -------------------------------------------------------------------------------------------------
def _mk(o, i, r, quant):
    from unsloth_b200.nf4 import quantize_nf4
    W = (torch.randn(o, i, device=DEV) * 0.02).to(BF)
    A = ((torch.rand(r, i, device=DEV) * 2 - 1) / i ** 0.5)
    B = torch.randn(o, r, device=DEV) * 0.02
    if quant:
        packed, qs = quantize_nf4(W)
        from unsloth_b200.kernels import fast_dequantize
        Wd = fast_dequantize(packed, qs).clone()
        return packed, qs, A, B, Wd
    return W, None, A, B, W


def _lora_truth(X32, Wd, A, B, s):
    return X32 @ Wd.float().t() + s * (X32 @ A.t()) @ B.t()


@pytest.mark.parametrize("quant", [False, True])
def test_lora_mlp_vs_reference(refk, quant):
    import unsloth_b200.kernels as K
    torch.manual_seed(21)
    T_, H, I, r, s = 8192, 4096, 14336, 16, 1.0
    X = torch.randn(4, T_ // 4, H, device=DEV).to(BF)
    dY = (torch.randn(4, T_ // 4, H, device=DEV) * 0.1).to(BF)
    gate, up, down = _mk(I, H, r, quant), _mk(I, H, r, quant), _mk(H, I, r, quant)

    def run(mod, fwd, bwd):
        P = [t.clone().requires_grad_() for p in (gate, up, down) for t in (p[2], p[3])]
        Xg = X.clone().requires_grad_()
        out = mod.apply(Xg * 1, gate[0], gate[1], P[0], P[1], s, up[0], up[1], P[2], P[3], s,
                        down[0], down[1], P[4], P[5], s, fwd, bwd, True)
        oc = out.detach().clone()
        out.backward(dY.clone())
        return [oc, Xg.grad] + [p.grad for p in P]
    rr = run(refk.fast_lora.LoRA_MLP, refk.swiglu.swiglu_fg_kernel, refk.swiglu.swiglu_DWf_DW_dfg_kernel)
    oo = run(K.LoRA_MLP, K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel)
    # fp32 truth by autograd on the (dequantised) weights
    P32 = [t.clone().requires_grad_() for p in (gate, up, down) for t in (p[2], p[3])]
    X32 = X.float().reshape(T_, H).requires_grad_()
    e = _lora_truth(X32, gate[4], P32[0], P32[1], s)
    g = _lora_truth(X32, up[4], P32[2], P32[3], s)
    h = torch.nn.functional.silu(e) * g
    o = _lora_truth(h, down[4], P32[4], P32[5], s)
    o.backward(dY.float().reshape(T_, H))
    truth = [o.detach().view(4, T_ // 4, H), X32.grad.view(4, T_ // 4, H)] + [p.grad for p in P32]
    names = ["out", "dX", "d_gateA", "d_gateB", "d_upA", "d_upB", "d_downA", "d_downB"]
    for n, a, b, t in zip(names, oo, rr, truth):
        gate_gemm("lora_mlp_%s:%s" % ("nf4" if quant else "bf16", n), a.float() if a.dtype != b.dtype else a,
                  b.float() if a.dtype != b.dtype else b, t)


@pytest.mark.parametrize("quant", [False, True])
def test_lora_qkv_o_vs_reference(refk, quant):
    import unsloth_b200.kernels as K
    torch.manual_seed(23)
    T_, H, KV, r, s = 8192, 4096, 1024, 16, 1.0
    X = torch.randn(4, T_ // 4, H, device=DEV).to(BF)
    dQ = (torch.randn(4, T_ // 4, H, device=DEV) * 0.1).to(BF)
    dK = (torch.randn(4, T_ // 4, KV, device=DEV) * 0.1).to(BF)
    dV = (torch.randn(4, T_ // 4, KV, device=DEV) * 0.1).to(BF)
    q, k, v = _mk(H, H, r, quant), _mk(KV, H, r, quant), _mk(KV, H, r, quant)

    def run(mod):
        P = [t.clone().requires_grad_() for p in (q, k, v) for t in (p[2], p[3])]
        Xg = X.clone().requires_grad_()
        Q, Kk, V = mod.apply(Xg * 1, q[0], q[1], P[0], P[1], s, k[0], k[1], P[2], P[3], s,
                             v[0], v[1], P[4], P[5], s, True)
        outs = [t.detach().clone() for t in (Q, Kk, V)]
        torch.autograd.backward([Q, Kk, V], [dQ.clone(), dK.clone(), dV.clone()])
        return outs + [Xg.grad] + [p.grad for p in P]
    rr, oo = run(refk.fast_lora.LoRA_QKV), run(K.LoRA_QKV)
    P32 = [t.clone().requires_grad_() for p in (q, k, v) for t in (p[2], p[3])]
    X32 = X.float().reshape(T_, H).requires_grad_()
    Q32 = _lora_truth(X32, q[4], P32[0], P32[1], s)
    K32 = _lora_truth(X32, k[4], P32[2], P32[3], s)
    V32 = _lora_truth(X32, v[4], P32[4], P32[5], s)
    torch.autograd.backward([Q32, K32, V32], [dQ.float().reshape(T_, H), dK.float().reshape(T_, KV),
                                              dV.float().reshape(T_, KV)])
    truth = [Q32.detach().view(4, -1, H), K32.detach().view(4, -1, KV), V32.detach().view(4, -1, KV),
             X32.grad.view(4, -1, H)] + [p.grad for p in P32]
    names = ["Q", "K", "V", "dX", "d_qA", "d_qB", "d_kA", "d_kB", "d_vA", "d_vB"]
    for n, a, b, t in zip(names, oo, rr, truth):
        gate_gemm("lora_qkv_%s:%s" % ("nf4" if quant else "bf16", n), a.float() if a.dtype != b.dtype else a,
                  b.float() if a.dtype != b.dtype else b, t)
    # o_proj: LoRA_W
    o = _mk(H, H, r, quant)
    dO = (torch.randn(4, T_ // 4, H, device=DEV) * 0.1).to(BF)

    def run_w(mod):
        A, B = o[2].clone().requires_grad_(), o[3].clone().requires_grad_()
        Xg = X.clone().requires_grad_()
        out = mod.apply(Xg * 1, o[0], o[1], A, B, s)
        oc = out.detach().clone()
        out.backward(dO.clone())
        return [oc, Xg.grad, A.grad, B.grad]
    rr, oo = run_w(refk.fast_lora.LoRA_W), run_w(K.LoRA_W)
    A32, B32 = o[2].clone().requires_grad_(), o[3].clone().requires_grad_()
    X32 = X.float().reshape(T_, H).requires_grad_()
    O32 = _lora_truth(X32, o[4], A32, B32, s)
    O32.backward(dO.float().reshape(T_, H))
    truth = [O32.detach().view(4, -1, H), X32.grad.view(4, -1, H), A32.grad, B32.grad]
    for n, a, b, t in zip(["O", "dX", "d_oA", "d_oB"], oo, rr, truth):
        gate_gemm("lora_w_%s:%s" % ("nf4" if quant else "bf16", n), a.float() if a.dtype != b.dtype else a,
                  b.float() if a.dtype != b.dtype else b, t)


def test_fast_dequantize_reference_host_path(refk):
    """The reference's own `fast_dequantize` (kernels/utils.py:567-679: two-stage absmax
    reconstruction, `+= offset`, then the NF4 expansion) driving the bitsandbytes-signature symbols
    exported by libunsloth_b200.so, against the one-launch `ub200_dequantize_nf4`: bit-exact."""
    from unsloth_b200.kernels import fast_dequantize
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(3)
    for shape in ((14336, 4096), (1024, 4096), (4096, 14336)):
        W = (torch.randn(*shape, device=DEV) * 0.02).to(BF)
        packed, qs = quantize_nf4(W)
        ours = fast_dequantize(packed, qs)
        ref = refk.utils.fast_dequantize(packed, qs)
        rec = record("fast_dequantize_%dx%d" % shape, ours, ref)
        assert rec["differ_frac"] == 0.0, rec
        ref_t = refk.utils.fast_dequantize(packed.t(), qs)         # transposed-weight contract (:678-679)
        assert ref_t.shape == (shape[1], shape[0])
        assert torch.equal(fast_dequantize(packed.t(), qs), ref_t)


# -------------------------------------------------------------------------------------------------
# a4 logits-free fused CE at the bench's own shape vs the reference's logits path
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,V,softcap", [(4096, 128256, 0.0), (3584, 256000, 30.0)])
def test_fused_ce_vs_reference_logits_path(refk, H, V, softcap):
    """unsloth_fused_ce_loss (chunk 2048, as the bench runs it) against the reference's
    UNSLOTH_RETURN_LOGITS route (models/llama.py:1525-1562): lm_head matmul (cuBLAS, bf16) ->
    shift -> reference `fast_cross_entropy_loss` -> autograd to the hidden states."""
    from unsloth_b200.kernels import unsloth_fused_ce_loss
    torch.manual_seed(29)
    B, S = 2, 2048
    hidden = torch.randn(B, S, H, device=DEV).to(BF)
    Wlm = (torch.randn(V, H, device=DEV) * 0.02).to(BF)
    labels = torch.randint(0, V, (B, S), device=DEV)
    labels[0, 5:40] = -100
    hr = hidden.clone().requires_grad_()
    logits = torch.nn.functional.linear(hr * 1, Wlm)
    shift = torch.full_like(labels, -100)
    shift[..., :-1] = labels[..., 1:]
    loss_r = refk.cross_entropy_loss.fast_cross_entropy_loss(logits, shift, logit_softcapping=softcap)
    loss_r.backward()
    del logits
    ho = hidden.clone().requires_grad_()
    loss_o = unsloth_fused_ce_loss(None, ho * 1, Wlm, None, labels, None, None, None,
                                   logit_softcapping=softcap, chunk_rows=2048)
    loss_o.backward()
    # fp32 truth on sampled rows (full fp32 logits would be 4 GB+)
    rows = torch.arange(0, B * S, 37, device=DEV)
    h32 = hidden.float().reshape(-1, H)[rows].requires_grad_()
    lg = h32 @ Wlm.float().t()
    if softcap:
        lg = softcap * torch.tanh(lg / softcap)
    n_items = (shift != -100).sum()
    l32 = torch.nn.functional.cross_entropy(lg, shift.reshape(-1)[rows], ignore_index=-100, reduction="sum") / n_items
    l32.backward()
    rec = record("fused_ce_H%d_V%d:loss" % (H, V), loss_o.reshape(1), loss_r.reshape(1))
    assert abs(loss_o.item() - loss_r.item()) <= 1e-3 * abs(loss_r.item()), rec
    gate_gemm("fused_ce_H%d_V%d:dH" % (H, V), ho.grad.reshape(-1, H)[rows], hr.grad.reshape(-1, H)[rows], h32.grad)
