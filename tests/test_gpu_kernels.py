"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C ABI
(unsloth_b200._lib -> libunsloth_b200.so), against
  (a) the golden vectors produced by the REFERENCE's Triton kernels (tests/golden/*.npz, fp32,
      tolerance 1e-5 -- the north_star fp32 gate), and
  (b) the CPU oracle (oracle/restate.py) on seeded bf16 inputs, including the reference's
      rounding points; gate rtol 1e-3 with atol = 1e-3 * max|ref|, plus the
      fraction of elements that differ at all (expected ~0 for the elementwise kernels), and
  (c) size-independent properties at BASELINE.json cfg2 sizes (Llama-3-8B, T = 4 x 2048).
"""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"
F32 = dict(rtol=1e-5, atol=1e-5)


def T(x, dtype=None):
    t = torch.from_numpy(np.asarray(x)).to(DEV)
    return t.to(dtype) if dtype is not None else t


def close(a, b, **kw):
    kw = {**F32, **kw}
    b = b if torch.is_tensor(b) else T(b)
    torch.testing.assert_close(a.detach().float().cpu(), b.detach().float().cpu(), **kw)


def bf16_gate(ours, ref, max_mismatch=0.002, rtol=1e-3):
    """rtol 1e-3 with atol 1e-3*max|ref|; elements outside must be < max_mismatch of all and
    within one bf16 ulp-ish (4e-3 relative of max) -- see SURVEY.md section 9."""
    o, r = ours.detach().float().cpu(), ref.detach().float().cpu()
    atol = 1e-3 * r.abs().max().item() + 1e-12
    bad = (o - r).abs() > (atol + rtol * r.abs())
    frac = bad.float().mean().item()
    assert frac <= max_mismatch, "mismatch fraction %.5f" % frac
    assert (o - r).abs().max().item() <= 1.6e-2 * r.abs().max().item() + 1e-6


class Norm:
    def __init__(self, w, eps):
        self.weight, self.variance_epsilon = w, eps


# -------------------------------------------------------------------------------------------
# RMSNorm
# -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rms_llama_512", "rms_llama_odd", "rms_gemma_256",
                                  "rms_selftest_512", "rms_selftest_1024"])
def test_rmsnorm_golden(golden, name):
    from unsloth_b200.kernels import fast_rms_layernorm
    g = golden(name)
    X = T(g["X"]).requires_grad_()
    Y = fast_rms_layernorm(Norm(T(g["W"]), float(g["eps"])), X, gemma=bool(g["gemma"]))
    close(Y, g["Y"])
    Y.backward(T(g["dY"]).clone())
    close(X.grad, g["dX"], atol=2e-5)
    if "dX_hf" in g:  # the reference's own self-test bar (rms_layernorm.py:326)
        assert (X.grad.cpu() - torch.from_numpy(g["dX_hf"])).abs().max() <= 0.05


@pytest.mark.parametrize("gemma,H", [(False, 4096), (True, 3584), (False, 2048), (False, 8192)])
def test_rmsnorm_bf16_vs_oracle(gemma, H):
    from unsloth_b200.kernels import fast_rms_layernorm
    torch.manual_seed(3407)
    X = torch.randn(3, 67, H).to(torch.bfloat16)
    W = (torch.randn(H) * 0.3 + (0 if gemma else 1)).to(torch.bfloat16)
    dY = torch.randn(3, 67, H).to(torch.bfloat16)
    Yr, r = R.rms_layernorm_fwd(X, W, 1e-5, gemma)
    dXr = R.rms_layernorm_bwd(dY, X, W, r, gemma)
    Xg = X.to(DEV).requires_grad_()
    Y = fast_rms_layernorm(Norm(W.to(DEV), 1e-5), Xg, gemma=gemma)
    dYg = dY.to(DEV).clone()
    Y.backward(dYg)
    bf16_gate(Y, Yr)
    bf16_gate(Xg.grad, dXr)
    if not gemma:  # in-place contract: dX is written over dY
        assert Xg.grad.data_ptr() == dYg.data_ptr()


@pytest.mark.parametrize("H,dt", [(4096, torch.bfloat16), (2048, torch.bfloat16), (256, torch.float16),
                                  (8192, torch.bfloat16)])
def test_add_rmsnorm_fused_vs_oracle_and_separate_ops(H, dt):
    """Residual add fused with the next RMSNorm: S and Y are BIT-EXACT against the two separate
    ops (torch's 16-bit add, then fast_rms_layernorm); the backward (norm gradient accumulated in
    place into the residual gradient) against the oracle's rms_layernorm_bwd + dS in fp32."""
    from unsloth_b200.kernels import fast_add_rms_layernorm, fast_rms_layernorm
    torch.manual_seed(11)
    A = torch.randn(2, 45, H).to(dt)
    B = (torch.randn(2, 45, H) * 0.5).to(dt)
    W = (torch.randn(H) * 0.3 + 1).to(dt)
    dS = torch.randn(2, 45, H).to(dt)
    dY = torch.randn(2, 45, H).to(dt)
    Ag, Bg = A.to(DEV).requires_grad_(), B.to(DEV).requires_grad_()
    norm = Norm(W.to(DEV), 1e-5)
    S, Y = fast_add_rms_layernorm(norm, Ag, Bg)
    S_sep = A.to(DEV) + B.to(DEV)
    Y_sep = fast_rms_layernorm(norm, S_sep)
    assert torch.equal(S, S_sep) and torch.equal(Y, Y_sep)
    Sr = A + B                                             # oracle: torch CPU add, one rounding
    assert torch.equal(S.cpu(), Sr)
    Yr, r = R.rms_layernorm_fwd(Sr, W, 1e-5, False)
    bf16_gate(Y, Yr)
    dSg = dS.to(DEV).clone()
    torch.autograd.backward([S, Y], [dSg, dY.to(DEV).clone()])
    ref = dS.float() + R.rms_layernorm_bwd(dY.float(), Sr, W, r, False).float()
    bf16_gate(Ag.grad, ref)
    assert torch.equal(Ag.grad, Bg.grad)
    assert torch.equal(dSg, Ag.grad)                       # accumulated in place into dS
    # S unused downstream (last norm of the stack): plain norm backward
    A2, B2 = A.to(DEV).requires_grad_(), B.to(DEV).requires_grad_()
    _, Y2 = fast_add_rms_layernorm(norm, A2, B2)
    Y2.backward(dY.to(DEV).clone())
    bf16_gate(A2.grad, R.rms_layernorm_bwd(dY, Sr, W, r, False))
    # fp32 activations fall back to the two separate ops (same results as fast_rms_layernorm)
    S3, Y3 = fast_add_rms_layernorm(Norm(W.float().to(DEV), 1e-5), A.float().to(DEV), B.float().to(DEV))
    Y3r, _ = R.rms_layernorm_fwd(A.float() + B.float(), W.float(), 1e-5, False)
    close(Y3, Y3r)


def test_rmsnorm_cfg2_properties():
    """Full cfg2 size (T=8192, H=4096): unit RMS of the output for W=1 and scale invariance."""
    from unsloth_b200.kernels import fast_rms_layernorm
    torch.manual_seed(0)
    X = torch.randn(4, 2048, 4096, device=DEV, dtype=torch.bfloat16)
    W = torch.ones(4096, device=DEV, dtype=torch.bfloat16)
    Y = fast_rms_layernorm(Norm(W, 1e-5), X)
    rms = Y.float().pow(2).mean(-1).sqrt()
    assert (rms - 1).abs().max() < 2e-2
    Y2 = fast_rms_layernorm(Norm(W, 0.0), X * 4)
    assert (Y2.float() - Y.float()).abs().max() < 4e-2


# -------------------------------------------------------------------------------------------
# RoPE
# -------------------------------------------------------------------------------------------
def test_rope_golden(golden):
    from unsloth_b200.kernels import fast_rope_embedding
    for name in ("rope_noindex", "rope_index"):
        g = golden(name)
        Q, K = T(g["Q"]).requires_grad_(), T(g["K"]).requires_grad_()
        idx = T(g["idx"]) if "idx" in g else None
        Qo, Ko = fast_rope_embedding(Q * 1.0, K * 1.0, T(g["cos"]), T(g["sin"]), idx)
        close(Qo, g["Qo"]); close(Ko, g["Ko"])
        torch.autograd.backward([Qo, Ko], [T(g["dQ"]).clone(), T(g["dK"]).clone()])
        close(Q.grad, g["gQ"]); close(K.grad, g["gK"])


def _tables(S, D, base=500000.0, dtype=torch.bfloat16):
    inv = 1.0 / (base ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(S).float(), inv)
    emb = torch.cat([fr, fr], -1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


@pytest.mark.parametrize("D,Hq,Hk,tdt", [(128, 32, 8, torch.bfloat16), (64, 32, 8, torch.bfloat16),
                                          (256, 16, 8, torch.float32)])
def test_rope_bf16_vs_oracle_inplace_strided(D, Hq, Hk, tdt):
    """The projection-buffer call form of models/llama.py:703-730: Q is a transposed VIEW of the
    contiguous [B,S,H*D] buffer and is rotated in place."""
    from unsloth_b200.kernels import fast_rope_embedding
    torch.manual_seed(1)
    B, S = 2, 77
    cos, sin = _tables(128, D, dtype=tdt)
    qbuf = torch.randn(B, S, Hq * D).to(torch.bfloat16)
    kbuf = torch.randn(B, S, Hk * D).to(torch.bfloat16)
    # native-Triton rounding (one product rounded, the other fused): oracle/restate.py::_rot
    Qr = R.rope_noindex(qbuf.view(B, S, Hq, D), cos, sin, contract=True).transpose(1, 2)
    Kr = R.rope_noindex(kbuf.view(B, S, Hk, D), cos, sin, contract=True).transpose(1, 2)
    qg, kg = qbuf.to(DEV), kbuf.to(DEV)
    Q = qg.view(B, S, Hq, D).transpose(1, 2)
    K = kg.view(B, S, Hk, D).transpose(1, 2)
    Qo, Ko = fast_rope_embedding(Q, K, cos.to(DEV), sin.to(DEV))
    assert Qo.data_ptr() == qg.data_ptr() and Ko.data_ptr() == kg.data_ptr()  # in place
    bf16_gate(Qo, Qr, max_mismatch=1e-4); bf16_gate(Ko, Kr, max_mismatch=1e-4)
    # indices path == explicit positions; round trip fwd -> bwd restores the input (rotation)
    idx = torch.randint(0, 128, (B * S,), dtype=torch.int32)
    Q2r, K2r = R.rope_qk(qbuf.view(B, S, Hq, D).transpose(1, 2), kbuf.view(B, S, Hk, D).transpose(1, 2),
                         cos, sin, idx, contract=True)
    qg2, kg2 = qbuf.to(DEV), kbuf.to(DEV)
    Q2, K2 = fast_rope_embedding(qg2.view(B, S, Hq, D).transpose(1, 2),
                                 kg2.view(B, S, Hk, D).transpose(1, 2), cos.to(DEV), sin.to(DEV),
                                 idx.to(DEV))
    bf16_gate(Q2, Q2r, max_mismatch=1e-4); bf16_gate(K2, K2r, max_mismatch=1e-4)


def test_rope_cfg2_roundtrip():
    """cfg2 size: forward then backward (rotation by -theta) restores Q and K."""
    from unsloth_b200 import _lib as L
    from unsloth_b200.kernels.rope_embedding import _launch
    torch.manual_seed(2)
    B, S, Hq, Hk, D = 4, 2048, 32, 8, 128
    cos, sin = _tables(S, D, dtype=torch.float32)
    cos, sin = cos.to(DEV), sin.to(DEV)
    q = torch.randn(B, S, Hq * D, device=DEV, dtype=torch.bfloat16)
    k = torch.randn(B, S, Hk * D, device=DEV, dtype=torch.bfloat16)
    q0, k0 = q.clone(), k.clone()
    Q, K = q.view(B, S, Hq, D).transpose(1, 2), k.view(B, S, Hk, D).transpose(1, 2)
    _launch(Q, K, cos, sin, None, False, True)
    assert (q.float() - q0.float()).abs().max() > 0.1
    n0 = q0.float().view(B, S, Hq, D).norm(dim=-1)
    assert ((q.float().view(B, S, Hq, D).norm(dim=-1) - n0).abs() / n0).max() < 2e-2  # isometry
    _launch(Q, K, cos, sin, None, True, True)
    assert (q.float() - q0.float()).abs().max() < 6e-2
    assert (k.float() - k0.float()).abs().max() < 6e-2


# -------------------------------------------------------------------------------------------
# SwiGLU / GEGLU
# -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["swiglu", "geglu_approx", "geglu_exact"])
def test_glu_golden_and_bf16(golden, name):
    import unsloth_b200.kernels as K
    fwd = {"swiglu": K.swiglu_fg_kernel, "geglu_approx": K.geglu_approx_forward_kernel,
           "geglu_exact": K.geglu_exact_forward_kernel}[name]
    bwd = {"swiglu": K.swiglu_DWf_DW_dfg_kernel, "geglu_approx": K.geglu_approx_backward_kernel,
           "geglu_exact": K.geglu_exact_backward_kernel}[name]
    ofwd = {"swiglu": R.swiglu_fwd, "geglu_approx": R.geglu_approx_fwd, "geglu_exact": R.geglu_exact_fwd}[name]
    obwd = {"swiglu": R.swiglu_bwd, "geglu_approx": R.geglu_approx_bwd, "geglu_exact": R.geglu_exact_bwd}[name]
    g = golden(name)
    e, up, DW = T(g["e"]), T(g["g"]), T(g["DW"])
    close(fwd(e, up), g["h"])
    h, df, de = bwd(DW.clone(), e.reshape(DW.shape).clone(), up.reshape(DW.shape).clone())
    close(h, g["bh"]); close(df, g["bdf"]); close(de, g["bde"], atol=2e-5)
    # bf16 vs oracle (rounding points), odd-ish size
    torch.manual_seed(5)
    eb = (torch.randn(3, 50, 1432) * 2).to(torch.bfloat16)
    gb = torch.randn(3, 50, 1432).to(torch.bfloat16)
    DWb = torch.randn(150, 1432).to(torch.bfloat16)
    bf16_gate(fwd(eb.to(DEV), gb.to(DEV)), ofwd(eb, gb), max_mismatch=2e-3)
    hr, dfr, der = obwd(DWb, eb.reshape(150, -1), gb.reshape(150, -1))
    DWg, eg, gg = DWb.to(DEV), eb.reshape(150, -1).to(DEV).contiguous(), gb.reshape(150, -1).to(DEV).contiguous()
    h2, df2, de2 = bwd(DWg, eg, gg)
    assert h2.data_ptr() == DWg.data_ptr() and df2.data_ptr() == eg.data_ptr() and de2.data_ptr() == gg.data_ptr()
    bf16_gate(h2, hr, max_mismatch=2e-3); bf16_gate(df2, dfr, max_mismatch=2e-3)
    bf16_gate(de2, der, max_mismatch=2e-3)


# -------------------------------------------------------------------------------------------
# Cross entropy
# -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ce_v1000", "ce_v70000_chunked", "ce_softcap30", "ce_scale"])
def test_cross_entropy_golden(golden, name):
    from unsloth_b200.kernels import fast_cross_entropy_loss
    g = golden(name)
    logits = T(g["logits"]).requires_grad_()
    lg = logits * 1.0
    lg.retain_grad()
    loss = fast_cross_entropy_loss(lg, T(g["labels"]), float(g["softcap"]), float(g["scale"]))
    close(loss, g["loss"])
    loss.backward()
    close(logits.grad, g["dlogits"], atol=1e-6)


@pytest.mark.parametrize("V,softcap", [(128256, 0.0), (32768, 0.0), (256000, 30.0)])
def test_cross_entropy_bf16_big_vocab(V, softcap):
    from unsloth_b200.kernels.cross_entropy_loss import _ce_forward, _ce_backward_
    torch.manual_seed(7)
    Tn = 6
    logits = (torch.randn(Tn, V) * 3).to(torch.bfloat16)
    labels = torch.randint(0, V, (Tn,)); labels[2] = -100
    lr, lser = R.cross_entropy_fwd(logits, labels, softcap, 0.0)
    dl = torch.full((Tn,), 0.25)
    dr = R.cross_entropy_bwd(logits, lser, labels, dl, softcap, 0.0)
    lg = logits.to(DEV)
    l, lse = _ce_forward(lg, labels.to(DEV), softcap, 0.0)
    close(l, lr, rtol=1e-5, atol=1e-4); close(lse, lser, rtol=1e-5, atol=1e-4)
    _ce_backward_(lg, lse, labels.to(DEV), dl.to(DEV), 1, softcap, 0.0)
    bf16_gate(lg, dr, max_mismatch=2e-3)
    assert lg[2].abs().max() == 0  # ignored row has zero gradient


# -------------------------------------------------------------------------------------------
# NF4
# -------------------------------------------------------------------------------------------
def test_nf4_dequant_bit_exact_and_bnb_symbols():
    from unsloth_b200 import _lib as L
    from unsloth_b200.kernels import fast_dequantize
    from unsloth_b200.nf4 import quantize_nf4, QuantState
    torch.manual_seed(3407)
    W = (torch.randn(1024, 4096) * 0.02).to(torch.bfloat16)
    packed_r, qs_r = R.quantize_nf4(W)
    packed, qs = quantize_nf4(W.to(DEV))
    assert torch.equal(packed.cpu(), packed_r)                       # integer/byte work: bit exact
    assert torch.equal(qs.absmax.cpu(), qs_r.absmax)
    close(qs.state2.absmax, qs_r.state2.absmax, rtol=0, atol=0)
    # dequantise the ORACLE's state on the GPU: byte-for-byte the oracle's answer.  (The GPU
    # quantiser's own `offset` = absmax.mean() differs from the CPU mean in the last bit, so
    # the two quantisers' states are compared field by field above and below instead.)
    assert abs(qs.offset.item() - qs_r.offset.item()) <= 1e-7 * abs(qs_r.offset.item()) + 1e-9
    Dr = R.dequantize_nf4(packed_r, qs_r)
    s2r = QuantState(qs_r.state2.absmax.to(DEV), code=qs_r.state2.code.to(DEV), blocksize=256)
    qs_o = QuantState(qs_r.absmax.to(DEV), qs_r.shape, None, 64, "nf4", torch.bfloat16,
                      qs_r.offset.to(DEV), s2r)
    Do = fast_dequantize(packed_r.to(DEV), qs_o)
    assert torch.equal(Do.cpu().view(torch.int16), Dr.view(torch.int16))   # bit exact
    D = fast_dequantize(packed, qs)
    assert D.dtype == torch.bfloat16 and D.shape == W.shape
    assert (D.cpu().view(torch.int16) != Dr.view(torch.int16)).float().mean() < 0.02
    assert (D.float().cpu() - Dr.float()).abs().max() <= 2.0 ** -8 * Dr.float().abs().max()
    # transposed-call contract and passthrough (kernels/utils.py:578-579, 678-679)
    assert fast_dequantize(packed.t(), qs).shape == (4096, 1024)
    assert fast_dequantize(W, None) is W
    out = torch.empty(1024, 4096, dtype=torch.bfloat16, device=DEV)
    assert fast_dequantize(packed, qs, out=out).data_ptr() == out.data_ptr()
    # old list-form quant_state (kernels/utils.py:594-598)
    s2 = qs.state2
    lst = [qs.absmax, qs.shape, qs.dtype, qs.blocksize, [qs.offset, [s2.absmax, s2.code, s2.blocksize, None, None, None, None]], None, None]
    assert torch.equal(fast_dequantize(packed, lst), D)
    # the bitsandbytes symbols, called the way the reference does (kernels/utils.py:650-675)
    n_abs = qs.absmax.numel()
    out_abs = torch.empty(n_abs, dtype=torch.float32, device=DEV)
    st = L.stream()
    L.lib.cdequantize_blockwise_fp32(L.ptr(s2.code), L.ptr(qs.absmax), L.ptr(s2.absmax), L.ptr(out_abs),
                                     s2.blocksize, n_abs, st)
    out_abs += qs.offset
    close(out_abs, R.dequantize_absmax(qs_r), rtol=1e-6, atol=0)
    out2 = torch.empty(1024, 4096, dtype=torch.bfloat16, device=DEV)
    L.lib.cdequantize_blockwise_bf16_nf4(None, L.ptr(packed), L.ptr(out_abs), L.ptr(out2), 64, out2.numel(), st)
    assert torch.equal(out2, D)
    out3 = torch.empty(1024, 4096, dtype=torch.float16, device=DEV)
    L.lib.cdequantize_blockwise_fp16_nf4(None, L.ptr(packed), L.ptr(out_abs), L.ptr(out3), 64, out3.numel(), st)
    qs16 = QuantState(qs.absmax, qs.shape, None, 64, "nf4", torch.float16, qs.offset, qs.state2)
    assert torch.equal(out3, fast_dequantize(packed, qs16))


@pytest.mark.parametrize("m,k,dt", [(4096, 4096, torch.bfloat16), (14336, 4096, torch.bfloat16),
                                    (1000, 1024, torch.float16), (4096, 14336, torch.bfloat16),
                                    (264, 1088, torch.bfloat16), (40, 96, torch.float16),
                                    (1000, 11008, torch.bfloat16), (50, 1152, torch.float16)])
def test_fast_gemv_nf4_vs_oracle_and_bnb_symbol(m, k, dt):
    """Decode-time GEMV on the packed weight: against the oracle's exact-product
    restatement, against dequantise-then-matmul on the GPU, and through the bitsandbytes symbol
    with the argument list of kernels/utils.py:955-973."""
    from unsloth_b200 import _lib as L
    from unsloth_b200.kernels import fast_dequantize, fast_gemv
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(m + k)
    W = (torch.randn(m, k) * 0.02).to(dt).to(DEV)
    packed, qs = quantize_nf4(W)
    x = torch.randn(1, 1, k).to(dt).to(DEV)
    out = fast_gemv(x, packed, qs)
    assert out.shape == (1, 1, m) and out.dtype == dt
    from types import SimpleNamespace as NS
    qs_cpu = NS(absmax=qs.absmax.cpu(), shape=qs.shape, dtype=dt, blocksize=64, offset=qs.offset.cpu(),
                state2=NS(absmax=qs.state2.absmax.cpu(), code=qs.state2.code.cpu(), blocksize=256))
    # the kernel's arithmetic: code table in the 16-bit dtype (as bitsandbytes' GEMV holds it), exact products,
    # fp32 absmax and accumulation, one rounding: at most one ulp of the output dtype from the restatement
    # (shapes outside the pair-table kernel's domain -- k % 128, k < 1024 -- run the fp32-table kernel)
    pair = k % 128 == 0 and k >= 1024
    ref = R.gemv_nf4(x.cpu(), packed.cpu(), qs_cpu, code_dtype=dt if pair else None)
    o, rf = out.view(-1).float().cpu(), ref.float()
    assert (o - rf).abs().max() <= 8e-3 * rf.abs().max()
    assert ((o - rf).abs() > 1e-3 * rf.abs().max()).float().mean() < 0.02
    # and against the fp32-table form (no rounding of the codes): the table rounding is worth < 1 output ulp
    rf32 = R.gemv_nf4(x.cpu(), packed.cpu(), qs_cpu).float()
    assert (o - rf32).abs().max() <= 1.2e-2 * rf32.abs().max()
    assert (o - rf32).abs().mean() <= 1.5e-3 * rf32.abs().max()
    Wd = fast_dequantize(packed, qs)
    ref2 = (Wd.float() @ x.view(-1).float())
    assert (out.view(-1).float() - ref2).abs().max() <= 1.2e-2 * ref2.abs().max()
    # the bitsandbytes route: fp32 absmax first (cdequantize_blockwise_fp32 + offset), then the GEMV
    absmax = torch.empty(qs.absmax.numel(), dtype=torch.float32, device=DEV)
    L.lib.cdequantize_blockwise_fp32(L.ptr(qs.state2.code), L.ptr(qs.absmax), L.ptr(qs.state2.absmax),
                                       L.ptr(absmax), 256, absmax.numel(), L.stream())
    absmax += qs.offset
    code = torch.tensor(R.NF4_CODE.tolist(), dtype=torch.float32, device=DEV)
    out2 = torch.empty((1, 1, m), dtype=dt, device=DEV)
    fx = L.lib.cgemm_4bit_inference_naive_bf16 if dt == torch.bfloat16 else L.lib.cgemm_4bit_inference_naive_fp16
    fx(m, 1, k, L.ptr(x), L.ptr(packed), L.ptr(absmax), L.ptr(code), L.ptr(out2), m, (k + 1) // 2, m, 64,
       L.stream())
    assert (out2.float() - out.float()).abs().max() <= 8e-3 * out.float().abs().max()


def test_fast_linear_forward_decode_and_merge_lora():
    """fast_linear_forward (kernels/utils.py:1082-1125) at bsz == q_len == 1 with LoRA in the GEMV
    epilogue, at bsz > 1 through the GEMM, on a dense weight, and `merge_lora` (save.py:620-646):
    x @ merged.T must equal the unmerged projection."""
    from unsloth_b200.kernels import fast_linear_forward, get_lora_parameters, fast_dequantize
    from unsloth_b200.lora import LoraLinear
    from unsloth_b200.nf4 import Linear4bit
    from unsloth_b200.save import merge_lora
    torch.manual_seed(21)
    k, m = 2048, 3072
    W = (torch.randn(m, k) * 0.02).to(torch.bfloat16).to(DEV)
    proj = LoraLinear(Linear4bit.from_dense(W), r=16, lora_alpha=32, init_b_std=0.05)
    Wq, qs, A, B, s = get_lora_parameters(proj)
    Wd = fast_dequantize(Wq, qs).float()
    full = Wd + s * (B.float() @ A.float())
    x1 = torch.randn(1, 1, k, device=DEV).to(torch.bfloat16)
    y1 = fast_linear_forward(proj, x1)
    assert y1.shape == (1, 1, m)
    ref1 = x1.view(-1).float() @ full.t()
    assert (y1.view(-1).float() - ref1).abs().max() <= 1.2e-2 * ref1.abs().max()
    from types import SimpleNamespace as NS
    qs_cpu = NS(absmax=qs.absmax.cpu(), shape=qs.shape, dtype=torch.bfloat16, blocksize=64,
                offset=qs.offset.cpu(),
                state2=NS(absmax=qs.state2.absmax.cpu(), code=qs.state2.code.cpu(), blocksize=256))
    ref_o = R.fast_linear_forward(x1.cpu(), Wq.cpu(), qs_cpu, A.detach().cpu(), B.detach().cpu(), s).float()
    assert (y1.view(-1).float().cpu() - ref_o).abs().max() <= 8e-3 * ref_o.abs().max()
    base_only = x1.view(-1).float() @ Wd.t()
    assert (ref1 - base_only).abs().max() > 0.05 * ref1.abs().max()        # the LoRA term matters
    x4 = torch.randn(4, 1, k, device=DEV).to(torch.bfloat16)
    y4 = fast_linear_forward(proj, x4)
    ref4 = x4.view(4, k).float() @ full.t()
    assert y4.shape == (4, 1, m)
    assert (y4.view(4, m).float() - ref4).abs().max() <= 1.5e-2 * ref4.abs().max()
    # dense 16-bit base (LoRA without quantisation) and the lm_head-style GEMV
    dense = LoraLinear(torch.nn.Linear(k, m, bias=False, device=DEV, dtype=torch.bfloat16), r=16,
                       lora_alpha=32, init_b_std=0.05)
    Wn, _, A2, B2, s2 = get_lora_parameters(dense)
    yd = fast_linear_forward(dense, x1)
    refd = x1.view(-1).float() @ (Wn.float() + s2 * (B2.float() @ A2.float())).t()
    assert (yd.view(-1).float() - refd).abs().max() <= 1.2e-2 * refd.abs().max()
    # merge: one rounding of W + sBA to bf16
    Wm, bias = merge_lora(proj, "proj")
    assert bias is None and Wm.dtype == torch.bfloat16 and Wm.shape == (m, k)
    # one rounding of an fp32 sum; cuBLAS may order the fp32 operations differently from `full`
    ref_m = full.to(torch.bfloat16)
    assert (Wm != ref_m).float().mean() < 1e-3
    assert (Wm.float() - full).abs().max() <= 2.0 ** -8 * full.abs().max()


def test_nf4_cfg2_size_properties():
    """A full Llama-3-8B gate_proj (14336 x 4096): quantise -> dequantise -> requantise is
    idempotent and the error is bounded by the NF4 grid."""
    from unsloth_b200.kernels import fast_dequantize
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(0)
    W = (torch.randn(14336, 4096, device=DEV) * 0.02).to(torch.bfloat16)
    packed, qs = quantize_nf4(W)
    D = fast_dequantize(packed, qs)
    blk = W.float().view(-1, 64)
    err = (D.float().view(-1, 64) - blk).abs().amax(1) / blk.abs().amax(1)
    assert err.max() < 0.2
    packed2, _ = quantize_nf4(D)
    assert (packed2 == packed).float().mean() > 0.995


# -------------------------------------------------------------------------------------------
# edge cases: empty inputs, ragged rows, argument errors surface as exceptions (never silently)
# -------------------------------------------------------------------------------------------
def test_edge_cases_empty_ragged_and_errors():
    import unsloth_b200.kernels as K
    H = 256
    norm = Norm(torch.ones(H, device=DEV, dtype=torch.bfloat16), 1e-5)
    # empty batch: nothing launched, shapes preserved
    Y = K.fast_rms_layernorm(norm, torch.empty(0, 5, H, device=DEV, dtype=torch.bfloat16))
    assert Y.shape == (0, 5, H)
    assert K.swiglu_fg_kernel(torch.empty(1, 0, 64, device=DEV, dtype=torch.bfloat16),
                              torch.empty(1, 0, 64, device=DEV, dtype=torch.bfloat16)).numel() == 0
    # a single row / ragged row counts (not multiples of any tile)
    for rows in (1, 3, 129, 1000):
        X = torch.randn(rows, H, device=DEV, dtype=torch.bfloat16)
        Yr, _ = R.rms_layernorm_fwd(X.cpu(), norm.weight.cpu(), 1e-5)
        bf16_gate(K.fast_rms_layernorm(norm, X), Yr)
    # hidden size that is not a multiple of the 16-byte vector -> explicit error, not garbage
    with pytest.raises(RuntimeError, match="bad argument"):
        K.fast_rms_layernorm(Norm(torch.ones(100, device=DEV, dtype=torch.bfloat16), 1e-5),
                             torch.randn(4, 100, device=DEV, dtype=torch.bfloat16))
    # GEMM operand with a leading dimension that TMA cannot address
    A = torch.randn(64, 100, device=DEV, dtype=torch.bfloat16)      # ld = 100 (not a multiple of 8)
    B = torch.randn(64, 100, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="bad argument"):
        K.gemm(64, 64, [(A, B, 100)], torch.empty(64, 64, device=DEV, dtype=torch.bfloat16))
    # mixed operand dtypes are rejected on the host side
    with pytest.raises(RuntimeError, match="mixed operand dtypes"):
        K.gemm(64, 64, [(A[:, :96].contiguous(), B[:, :96].contiguous().half(), 96)],
               torch.empty(64, 64, device=DEV, dtype=torch.bfloat16))
    # all labels ignored: loss 0 / n_items guards are the caller's (reference semantics), grads are zero
    logits = torch.randn(1, 4, 300, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    lab = torch.full((1, 4), -100, device=DEV)
    loss = K.fast_cross_entropy_loss(logits * 1.0, lab, n_items=1)
    loss.backward()
    assert loss.item() == 0 and logits.grad.abs().max().item() == 0


def test_fp16_path_matches_oracle():
    """fp16 activations (the reference's T4 CI dtype): RMSNorm, RoPE, SwiGLU and the GEMM."""
    import unsloth_b200.kernels as K
    torch.manual_seed(4)
    X = torch.randn(5, 40, 1024).half(); W = (torch.randn(1024) * 0.2 + 1).half()
    Yr, _ = R.rms_layernorm_fwd(X, W, 1e-5)
    Y = K.fast_rms_layernorm(Norm(W.to(DEV), 1e-5), X.to(DEV))
    assert (Y.float().cpu() - Yr.float()).abs().max() <= 2e-3 * Yr.float().abs().max()
    e, g = torch.randn(1, 30, 512).half(), torch.randn(1, 30, 512).half()
    assert (K.swiglu_fg_kernel(e.to(DEV), g.to(DEV)).float().cpu() - R.swiglu_fwd(e, g).float()).abs().max() < 4e-3
    cos, sin = _tables(64, 64, dtype=torch.float16)
    q, k = torch.randn(1, 20, 4 * 64).half(), torch.randn(1, 20, 2 * 64).half()
    Qr = R.rope_noindex(q.view(1, 20, 4, 64), cos, sin).transpose(1, 2)
    Qo, _ = K.fast_rope_embedding(q.to(DEV).view(1, 20, 4, 64).transpose(1, 2), k.to(DEV).view(1, 20, 2, 64).transpose(1, 2),
                                  cos.to(DEV), sin.to(DEV))
    assert (Qo.float().cpu() - Qr.float()).abs().max() < 4e-3
    A = torch.randn(200, 256, device=DEV).half(); B = torch.randn(136, 256, device=DEV).half()
    out = torch.empty(200, 136, device=DEV, dtype=torch.float16)
    K.gemm(200, 136, [(A, B, 256)], out)
    ref = A.float() @ B.float().t()
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 2e-3


def test_cast_pad_multi_and_accumulate_multi_match_the_single_tensor_forms():
    """ub200_cast_pad_multi / ub200_accumulate_multi (descriptors by value, chunks of 40 per launch): every
    descriptor = ub200_cast_pad_2d / `dst += view`, bit for bit, over more than two chunks of mixed shapes,
    offsets, scales, transposes and (for the accumulation) transposed / column-sliced source views."""
    from unsloth_b200 import _lib as L
    from unsloth_b200.kernels.utils import cast_pad
    torch.manual_seed(5)
    n = 97
    srcs, dsts, refs, arr = [], [], [], (L.CastDesc * n)()
    for i in range(n):
        r, c = [(16, 4096), (8, 1024), (4096, 16), (33, 70), (1, 5)][i % 5]
        tr = i % 3 == 1
        dt = [torch.bfloat16, torch.float16, torch.float32][i % 3]
        src = torch.randn(r, c, device=DEV)
        pr, pc = (c, r) if tr else (r, c)
        ro, co = (i % 2) * 3, (i % 4) * 5
        dst = torch.full((pr + ro + 2, pc + co + 7), 7.0, dtype=dt, device=DEV)
        ref = torch.full_like(dst, 7.0)
        cast_pad(src, ref, ro, co, 0.5 + i, tr)
        d = arr[i]
        d.src, d.dst, d.src_ld, d.dst_ld = src.data_ptr(), dst.data_ptr(), src.stride(0), dst.stride(0)
        d.src_dtype, d.dst_dtype, d.rows, d.cols = L.dt(src), L.dt(dst), r, c
        d.dst_rows, d.dst_cols, d.row_off, d.col_off, d.scale, d.transpose = dst.shape[0], dst.shape[1], ro, co, 0.5 + i, int(tr)
        srcs.append(src); dsts.append(dst); refs.append(ref)
    L.call("ub200_cast_pad_multi", arr, n, L.stream())
    for a, b in zip(dsts, refs):
        assert torch.equal(a, b)
    # the 16-byte-store path of the batched kernel: the shapes a LoRA step really casts (rank blocks padded to
    # 64, 16-bit destinations with 8-column granularity), against the single-tensor form
    cases = [((16, 4096), (64, 4096), 16, 0, False), ((16, 14336), (64, 14336), 0, 0, False),
             ((14336, 16), (14336, 64), 0, 32, False), ((1024, 16), (64, 1024), 16, 0, True),
             ((4096, 16), (4096, 64), 0, 0, False), ((24, 520), (40, 528), 8, 8, False)]
    m = len(cases) * 2
    arr2, keep = (L.CastDesc * m)(), []
    for i in range(m):
        (r, c), dshape, ro, co, tr = cases[i % len(cases)]
        dt = [torch.bfloat16, torch.float16][i // len(cases)]
        src = torch.randn(r, c, device=DEV)
        dst = torch.full(dshape, 7.0, dtype=dt, device=DEV)
        ref = torch.full_like(dst, 7.0)
        cast_pad(src, ref, ro, co, 1.5 + i, tr)
        d = arr2[i]
        d.src, d.dst, d.src_ld, d.dst_ld = src.data_ptr(), dst.data_ptr(), src.stride(0), dst.stride(0)
        d.src_dtype, d.dst_dtype, d.rows, d.cols = L.dt(src), L.dt(dst), r, c
        d.dst_rows, d.dst_cols, d.row_off, d.col_off, d.scale, d.transpose = dshape[0], dshape[1], ro, co, 1.5 + i, int(tr)
        keep.append((src, dst, ref))
    L.call("ub200_cast_pad_multi", arr2, m, L.stream())
    for _, a, b in keep:
        assert torch.equal(a, b)
    acc = (L.AccDesc * n)()
    gs, views, exp = [], [], []
    for i in range(n):
        r, c = [(16, 4096), (4096, 16), (8, 1024), (5, 3)][i % 4]
        full = torch.randn(c, 64, device=DEV) if i % 2 == 0 else torch.randn(r, 64, device=DEV)
        view = full[:, 3:3 + r].t() if i % 2 == 0 else full[:, 7:7 + c]
        if view.shape != (r, c):
            full = torch.randn(r, max(c, 64) + 9, device=DEV); view = full[:, 7:7 + c]
        g = torch.randn(r, c, device=DEV)
        exp.append(g + view)
        a = acc[i]
        a.src, a.dst, a.src_rs, a.src_cs, a.rows, a.cols = view.data_ptr(), g.data_ptr(), view.stride(0), view.stride(1), r, c
        gs.append(g); views.append((full, view))
    L.call("ub200_accumulate_multi", acc, n, L.stream())
    for g, e in zip(gs, exp):
        assert torch.equal(g, e)
