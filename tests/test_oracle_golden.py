"""Pins oracle/restate.py (the CPU oracle) against the committed golden vectors that
oracle/make_golden.py produced by executing the REFERENCE's own Triton kernels
(TRITON_INTERPRET=1, fp32).  Tolerance: 1e-5 (north_star fp32 gate)."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

TOL = dict(rtol=1e-5, atol=1e-5)


def T(x):
    return torch.from_numpy(np.asarray(x))


def close(a, b, **kw):
    kw = {**TOL, **kw}
    torch.testing.assert_close(a, T(b) if not torch.is_tensor(b) else b, **kw)


@pytest.mark.parametrize("name", ["rms_llama_512", "rms_llama_odd", "rms_gemma_256",
                                  "rms_selftest_512", "rms_selftest_1024"])
def test_rmsnorm(golden, name):
    g = golden(name)
    X, W, dY = T(g["X"]), T(g["W"]), T(g["dY"])
    gemma = bool(g["gemma"])
    Y, r = R.rms_layernorm_fwd(X, W, float(g["eps"]), gemma)
    close(Y, g["Y"])
    dX = R.rms_layernorm_bwd(dY, X, W, r, gemma)
    close(dX, g["dX"], atol=2e-5)
    if "Y_hf" in g:  # reference self-test bar (kernels/rms_layernorm.py:326): amax<=0.05
        assert (dX - T(g["dX_hf"])).abs().max() <= 0.05
        close(Y, g["Y_hf"])


def test_rope_noindex(golden):
    g = golden("rope_noindex")
    Q, K, cos, sin = (T(g[k]) for k in ("Q", "K", "cos", "sin"))
    # the reference transposes to [B,S,H,D] before its kernel (rope_embedding.py:276)
    Qo = R.rope_noindex(Q.transpose(1, 2), cos, sin).transpose(1, 2)
    Ko = R.rope_noindex(K.transpose(1, 2), cos, sin).transpose(1, 2)
    close(Qo, g["Qo"]); close(Ko, g["Ko"])
    gQ = R.rope_noindex(T(g["dQ"]).transpose(1, 2), cos, sin, backward=True).transpose(1, 2)
    gK = R.rope_noindex(T(g["dK"]).transpose(1, 2), cos, sin, backward=True).transpose(1, 2)
    close(gQ, g["gQ"]); close(gK, g["gK"])


def test_rope_index(golden):
    g = golden("rope_index")
    Q, K, cos, sin, idx = (T(g[k]) for k in ("Q", "K", "cos", "sin", "idx"))
    Qo, Ko = R.rope_qk(Q, K, cos, sin, idx)
    close(Qo, g["Qo"]); close(Ko, g["Ko"])
    gQ, gK = R.rope_qk(T(g["dQ"]), T(g["dK"]), cos, sin, idx, backward=True)
    close(gQ, g["gQ"]); close(gK, g["gK"])
    # no indices == arange positions
    Q2, K2 = R.rope_qk(Q, K, cos, sin, None)
    Q3, K3 = R.rope_qk(Q, K, cos, sin, torch.arange(Q.shape[2]).repeat(Q.shape[0]))
    close(Q2, Q3); close(K2, K3)


@pytest.mark.parametrize("name", ["ce_v1000", "ce_v70000_chunked", "ce_softcap30", "ce_scale"])
def test_cross_entropy(golden, name):
    g = golden(name)
    logits, labels = T(g["logits"]), T(g["labels"])
    loss, dlogits = R.fast_cross_entropy_loss(logits, labels, float(g["softcap"]), float(g["scale"]))
    close(loss, g["loss"])
    close(dlogits, g["dlogits"], atol=1e-6)


@pytest.mark.parametrize("name,fwd,bwd", [
    ("swiglu", R.swiglu_fwd, R.swiglu_bwd),
    ("geglu_approx", R.geglu_approx_fwd, R.geglu_approx_bwd),
    ("geglu_exact", R.geglu_exact_fwd, R.geglu_exact_bwd)])
def test_glu(golden, name, fwd, bwd):
    g = golden(name)
    e, up, DW = T(g["e"]), T(g["g"]), T(g["DW"])
    close(fwd(e, up), g["h"])
    h, df, de = bwd(DW, e.reshape(DW.shape), up.reshape(DW.shape))
    close(h, g["bh"]); close(df, g["bdf"]); close(de, g["bde"], atol=2e-5)


@pytest.mark.parametrize("act", ["swiglu", "geglu_approx"])
def test_lora_mlp(golden, act):
    g = golden("lora_mlp_" + act)
    t = {k: T(v) for k, v in g.items()}
    s = float(g["s"])
    gate = (t["gW"], None, t["gA"], t["gB"], s)
    up = (t["uW"], None, t["uA"], t["uB"], s)
    down = (t["dW"], None, t["dA"], t["dB"], s)
    out, e, gg = R.lora_mlp_fwd(t["X"], gate, up, down, act)
    close(out, g["out"])
    dX, (dgA, dgB), (duA, duB), (ddA, ddB) = R.lora_mlp_bwd(t["dY"], t["X"], e, gg, gate, up, down, act)
    close(dX, g["dX"])
    for mine, key in ((dgA, "d_gA"), (dgB, "d_gB"), (duA, "d_uA"), (duB, "d_uB"),
                      (ddA, "d_dA"), (ddB, "d_dB")):
        close(mine, g[key], atol=2e-5)


def test_lora_qkv(golden):
    g = golden("lora_qkv")
    t = {k: T(v) for k, v in g.items()}
    s = float(g["s"])
    q = (t["qW"], None, t["qA"], t["qB"], s)
    k = (t["kW"], None, t["kA"], t["kB"], s)
    v = (t["vW"], None, t["vA"], t["vB"], s)
    Q, K, V = R.lora_qkv_fwd(t["X"], q, k, v)
    close(Q, g["Q"]); close(K, g["K"]); close(V, g["V"])
    dX, gq, gk, gv = R.lora_qkv_bwd(t["dQ"], t["dK"], t["dV"], t["X"], q, k, v)
    close(dX, g["dX"])
    for (a, b), n in ((gq, "q"), (gk, "k"), (gv, "v")):
        close(a, g["d_%sA" % n], atol=2e-5); close(b, g["d_%sB" % n], atol=2e-5)


def test_lora_w(golden):
    g = golden("lora_w")
    t = {k: T(v) for k, v in g.items()}
    o = (t["oW"], None, t["oA"], t["oB"], float(g["s"]))
    close(R.lora_w_fwd(t["X"], o), g["O"])
    dX, (dA, dB) = R.lora_w_bwd(t["dY"], t["X"], o)
    close(dX, g["dX"]); close(dA, g["d_oA"], atol=2e-5); close(dB, g["d_oB"], atol=2e-5)


def test_nf4_roundtrip():
    """NF4 double-quant restatement (parity unpinned): self-consistency properties."""
    torch.manual_seed(3407)
    W = torch.randn(128, 256) * 0.02
    packed, qs = R.quantize_nf4(W.to(torch.bfloat16))
    assert packed.shape == (128 * 256 // 2, 1) and packed.dtype == torch.uint8
    assert qs.absmax.dtype == torch.uint8 and qs.absmax.numel() == 128 * 256 // 64
    assert qs.state2.code.numel() == 256 and qs.state2.absmax.numel() == 2
    D = R.dequantize_nf4(packed, qs)
    assert D.shape == W.shape and D.dtype == torch.bfloat16
    # quantisation error bounded by half the largest NF4 gap times the block absmax
    blk = W.reshape(-1, 64)
    err = (D.float().reshape(-1, 64) - blk).abs().amax(1) / blk.abs().amax(1)
    assert err.max() < 0.2
    # idempotence: re-quantising the dequantised weight reproduces the same codes
    packed2, _ = R.quantize_nf4(D)
    assert (packed2 == packed).float().mean() > 0.99
    # transposed-call contract (kernels/utils.py:678-679)
    assert R.fast_dequantize(packed.t(), qs).shape == (256, 128)
    assert R.fast_dequantize(W, None) is W


def test_fused_linear_ce_matches_logits_path():
    torch.manual_seed(0)
    B, S, H, V = 2, 7, 32, 101
    hidden = torch.randn(B, S, H)
    Wlm = torch.randn(V, H) * 0.1
    labels = torch.randint(0, V, (B, S)); labels[0, 3] = -100
    loss, dH = R.fused_linear_cross_entropy(hidden, Wlm, labels)
    h = hidden.clone().requires_grad_()
    logits = h @ Wlm.t()
    ref = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1),
                                            ignore_index=-100)
    ref.backward()
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dH, h.grad, rtol=1e-4, atol=1e-6)


def test_packing_label_vector(golden):
    """tests/utils/test_packing.py:1489-1525 of the reference: boundary-masked labels."""
    g = golden("packing_labels")
    from unsloth_b200.packing import mask_packed_boundary_labels
    out = mask_packed_boundary_labels(T(g["labels"])[None].clone(), T(g["packed_seq_lengths"]))
    assert out[0].tolist() == g["expected"].tolist()


def test_oracle_gemv_restatement_matches_its_dequant():
    """`gemv_nf4` / `fast_linear_forward` (decode path, kernels/utils.py:874-973, :1082-1125) are the
    same NF4 expansion as `dequantize_nf4`, contracted with x: the two restatements must agree to
    the rounding of the 16-bit dequantised weights."""
    import torch
    from oracle import restate as R
    torch.manual_seed(0)
    W = (torch.randn(96, 256) * 0.02).to(torch.bfloat16)
    p, qs = R.quantize_nf4(W)
    x = torch.randn(256).to(torch.bfloat16)
    y = R.gemv_nf4(x, p, qs).float()
    yd = R.dequantize_nf4(p, qs).float() @ x.float()
    assert (y - yd).abs().max() <= 8e-3 * yd.abs().max()
    A = torch.randn(8, 256) * 0.05
    B = torch.randn(96, 8) * 0.05
    full = R.fast_linear_forward(x, p, qs, A, B, 2.0).float()
    ref = yd + 2.0 * (B.to(torch.bfloat16).float() @ (A.to(torch.bfloat16).float() @ x.float()))
    assert (full - ref).abs().max() <= 8e-3 * ref.abs().max()
    assert (full - y).abs().max() > 1e-2 * ref.abs().max()          # the LoRA term is present


# ----------------------------------------------------------------------------------------------
# 16-bit rounding points: the reference's Triton kernels run in fp16 under
# the interpreter; the oracle's emulation of every rounding point must reproduce them to the last
# bit up to FMA-contraction / libm noise: at most one fp16 ulp on at most 0.5 % of the elements.
# ----------------------------------------------------------------------------------------------
def ulp_close(ours, ref, max_frac=5e-3, max_ulps=1):
    ours, ref = ours.detach(), T(ref) if not torch.is_tensor(ref) else ref
    assert ours.dtype == torch.float16 and ref.dtype == torch.float16
    a, b = ours.float().reshape(-1), ref.float().reshape(-1)
    ulp = torch.maximum(b.abs(), torch.tensor(6.1e-5)) * 2.0 ** -10      # fp16: 10 stored mantissa bits
    bad = (a - b).abs() > 0
    assert bad.float().mean().item() <= max_frac, bad.float().mean().item()
    # 1e-6 absolute floor: the subnormal tail, where (1 + erf) / (1 + tanh) cancel catastrophically
    ok = ((a - b).abs() <= (max_ulps + 0.001) * ulp) | ((a - b).abs() <= 1e-6)
    assert ok.all(), ((a - b).abs() / ulp).max().item()


def test_fp16_rounding_points_rmsnorm(golden):
    g = golden("fp16_rms_llama")
    X, W, dY = T(g["X"]), T(g["W"]), T(g["dY"])
    Y, r = R.rms_layernorm_fwd(X, W, float(g["eps"]), False)
    ulp_close(Y, g["Y"], max_frac=0.0)                 # `normed -> W.dtype` then `* W`: bit exact
    ulp_close(R.rms_layernorm_bwd(dY, X, W, r, False), g["dX"])


@pytest.mark.parametrize("name,fwd,bwd", [
    ("fp16_swiglu", R.swiglu_fwd, R.swiglu_bwd),
    ("fp16_geglu_approx", R.geglu_approx_fwd, R.geglu_approx_bwd),
    ("fp16_geglu_exact", R.geglu_exact_fwd, R.geglu_exact_bwd)])
def test_fp16_rounding_points_glu(golden, name, fwd, bwd):
    g = golden(name)
    e, up, DW = T(g["e"]), T(g["g"]), T(g["DW"])
    # tanh / erf come from different libms (Triton interpreter vs torch): allow 2 / 4 ulps there
    mu = {"fp16_swiglu": 1, "fp16_geglu_approx": 2, "fp16_geglu_exact": 4}[name]
    ulp_close(fwd(e, up), g["h"], max_ulps=mu)
    h, df, de = bwd(DW, e.reshape(-1, e.shape[-1]), up.reshape(-1, up.shape[-1]))
    ulp_close(h, g["bh"], max_ulps=mu); ulp_close(df, g["bdf"], max_ulps=mu); ulp_close(de, g["bde"], max_ulps=mu)


def test_fp16_rounding_points_rope(golden):
    g = golden("fp16_rope")
    Q, K, cos, sin, idx = (T(g[k]) for k in ("Q", "K", "cos", "sin", "idx"))
    Qo = R.rope_noindex(Q.transpose(1, 2), cos, sin).transpose(1, 2)
    Ko = R.rope_noindex(K.transpose(1, 2), cos, sin).transpose(1, 2)
    ulp_close(Qo, g["Qo"]); ulp_close(Ko, g["Ko"])
    Qi, Ki = R.rope_qk(Q, K, cos, sin, idx)
    ulp_close(Qi, g["Qi"]); ulp_close(Ki, g["Ki"])


def test_fp16_rounding_points_cross_entropy(golden):
    g = golden("fp16_ce")
    logits, labels = T(g["logits"]), T(g["labels"])
    V = logits.shape[-1]
    loss, lse = R.cross_entropy_fwd(logits.view(-1, V), labels.view(-1), 0.0, 0.0)
    n = (labels != -100).sum()
    close(loss.sum() / n, g["loss"], rtol=1e-5, atol=1e-5)
    dl = torch.full((labels.numel(),), 1.0 / n.item())
    dlog = R.cross_entropy_bwd(logits.view(-1, V), lse, labels.view(-1), dl, 0.0, 0.0)
    ulp_close(dlog.view_as(logits), g["dlogits"], max_frac=2e-2)


def _eq16(mine, ref, what):
    """LoRA rounding points in 16 bit: the reference's functions and the oracle do the same fp32-
    accumulated products and round at the same places, so results agree to the bit (a 1-ulp
    difference on <= 0.5 % of the elements is tolerated for BLAS summation-order noise)."""
    ref = T(ref)
    if ref.dtype == torch.float16:
        ulp_close(mine, ref)
    else:       # adapter gradients: 16-bit values stored in fp32
        torch.testing.assert_close(mine.float(), ref.float(), rtol=1e-3, atol=1e-6, msg=what)
        assert (mine.float() != ref.float()).float().mean() <= 5e-3, what


def test_fp16_rounding_points_lora(golden):
    g = golden("fp16_lora_w")
    t = {k: T(v) for k, v in g.items()}
    o = (t["oW"], None, t["oA"], t["oB"], float(g["s"]))
    _eq16(R.lora_w_fwd(t["X"], o), g["O"], "O")
    dX, (dA, dB) = R.lora_w_bwd(t["dY"], t["X"], o)
    _eq16(dX, g["dX"], "dX"); _eq16(dA, g["d_oA"], "dA"); _eq16(dB, g["d_oB"], "dB")

    g = golden("fp16_lora_qkv")
    t = {k: T(v) for k, v in g.items()}
    s = float(g["s"])
    q, k, v = ((t[n + "W"], None, t[n + "A"], t[n + "B"], s) for n in "qkv")
    Q, K, V = R.lora_qkv_fwd(t["X"], q, k, v)
    _eq16(Q, g["Q"], "Q"); _eq16(K, g["K"], "K"); _eq16(V, g["V"], "V")
    dX, gq, gk, gv = R.lora_qkv_bwd(t["dQ"], t["dK"], t["dV"], t["X"], q, k, v)
    _eq16(dX, g["dX"], "dX")
    for (a, b), n in ((gq, "q"), (gk, "k"), (gv, "v")):
        _eq16(a, g["d_%sA" % n], n + "A"); _eq16(b, g["d_%sB" % n], n + "B")

    g = golden("fp16_lora_mlp_swiglu")
    t = {k: T(v) for k, v in g.items()}
    s = float(g["s"])
    gate, up, down = ((t[n + "W"], None, t[n + "A"], t[n + "B"], s) for n in "gud")
    out, e, gg = R.lora_mlp_fwd(t["X"], gate, up, down, "swiglu")
    _eq16(out, g["out"], "out")
    dX, (dgA, dgB), (duA, duB), (ddA, ddB) = R.lora_mlp_bwd(t["dY"], t["X"], e, gg, gate, up, down, "swiglu")
    _eq16(dX, g["dX"], "dX")
    for mine, key in ((dgA, "d_gA"), (dgB, "d_gB"), (duA, "d_uA"), (duB, "d_uB"), (ddA, "d_dA"), (ddB, "d_dB")):
        _eq16(mine, g[key], key)


def test_attention_oracle_vs_torch_sdpa_and_hf_masks():
    """oracle/restate.py::attention (the definition the GPU attention tests use) against torch's own
    scaled_dot_product_attention (plain causal, GQA) and against the explicit sliding-window /
    soft-capping formula of transformers' eager Gemma-2 attention."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    B, S, Hq, Hk, D = 2, 37, 4, 2, 16
    Q, K, V = torch.randn(B, S, Hq, D), torch.randn(B, S, Hk, D), torch.randn(B, S, Hk, D)
    O, lse = R.attention(Q, K, V, D ** -0.5)
    ref = F.scaled_dot_product_attention(Q.transpose(1, 2), K.transpose(1, 2).repeat_interleave(2, 1),
                                         V.transpose(1, 2).repeat_interleave(2, 1), is_causal=True).transpose(1, 2)
    torch.testing.assert_close(O, ref, rtol=1e-5, atol=1e-5)
    # window w: HF masks i - j >= sliding_window, the reference's flash-attn window (w, w) keeps w + 1 keys
    w, cap = 5, 20.0
    O2, _ = R.attention(Q, K, V, D ** -0.5, window_left=w, softcap=cap)
    q, k, v = Q.transpose(1, 2), K.transpose(1, 2).repeat_interleave(2, 1), V.transpose(1, 2).repeat_interleave(2, 1)
    s = (q @ k.transpose(-1, -2)) * D ** -0.5
    s = torch.tanh(s / cap) * cap                                    # modeling_gemma2.py eager_attention_forward
    i, j = torch.arange(S)[:, None], torch.arange(S)[None, :]
    s = s.masked_fill(~((j <= i) & (i - j < w + 1)), float("-inf"))
    torch.testing.assert_close(O2, (torch.softmax(s, -1) @ v).transpose(1, 2), rtol=1e-5, atol=1e-5)
    # packed rows == separate documents
    lengths = [10, 27]
    Op, _ = R.attention(Q[:1], K[:1], V[:1], D ** -0.5, lengths=lengths)
    Oa, _ = R.attention(Q[:1, :10], K[:1, :10], V[:1, :10], D ** -0.5)
    Ob, _ = R.attention(Q[:1, 10:], K[:1, 10:], V[:1, 10:], D ** -0.5)
    torch.testing.assert_close(Op, torch.cat([Oa, Ob], 1), rtol=1e-5, atol=1e-5)


def test_nf4_code_book_is_the_published_normal_float_construction():
    """bitsandbytes is absent, so the 16 NF4 levels hard-coded in oracle/restate.py, unsloth_b200/nf4.py and
    csrc/nf4.cu cannot be read off its kernels.  They CAN be re-derived: NF4 is defined (QLoRA, Dettmers et al.
    2023, appendix E; bitsandbytes.functional.create_normal_map(offset=0.9677083, use_extra_value=True)) as the
    normalised quantiles of N(0, 1) -- 8 levels on the positive side, 7 on the negative side, plus an exact
    zero.  Evaluating that construction with scipy reproduces the table BIT FOR BIT in fp32, which pins the
    table independently of our own transcription; the same check for the 256-entry dynamic map
    (create_dynamic_map, the second-level code of double quantisation) pins its end points and symmetry."""
    from scipy.stats import norm
    offset = 0.9677083
    pos = norm.ppf(torch.linspace(offset, 0.5, 9)[:-1]).tolist()
    neg = (-norm.ppf(torch.linspace(offset, 0.5, 8)[:-1])).tolist()
    v = torch.tensor(pos + [0.0] + neg, dtype=torch.float32).sort().values
    v = v / v.max()
    assert torch.equal(v, R.NF4_CODE)
    from unsloth_b200 import nf4 as product_nf4
    table = getattr(product_nf4, "NF4_CODE", None)
    if table is not None:
        assert torch.equal(torch.as_tensor(table, dtype=torch.float32).cpu(), R.NF4_CODE)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unsloth_b200", "csrc",
                            "nf4.cu")).read()
    for x in R.NF4_CODE.tolist():
        if x not in (0.0, 1.0, -1.0):
            assert repr(abs(x))[:12] in src, x        # the kernel's constant table holds the same literals
    code = R.create_dynamic_map()
    assert code.numel() == 256 and torch.equal(code, code.sort().values)
    assert code[-1] == 1.0 and (code == 0).sum() == 1 and abs(code[0].item() + 0.9929687380790710) < 1e-7
    # symmetric but for the extra +1.0: every negative level has its positive mirror image
    assert torch.allclose(-code[code < 0].flip(0), code[code > 0][:-1], rtol=0, atol=0)
