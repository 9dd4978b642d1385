"""HOST LOGIC of the shipped Python layer (kernels/*.py, patch.py) on a CPU-only box: every C-ABI
call is interpreted by tests/abi_emulator.py (oracle arithmetic on the raw pointers / strides the
shims pass), so what is exercised here is everything ABOVE the ABI -- views and strides handed to
the kernels, in-place contracts, label shifting and item counts, chunking of the fused CE, the
rank-block orchestration, the fused add+norm chain, packed batches, UNSLOTH_RETURN_LOGITS.
References: the golden vectors produced by the reference's own kernels, and the stock HuggingFace
model (the configs[0] reference path).  The GPU suite checks the same calls on the real library."""
import copy

import numpy as np
import pytest
import torch

from oracle import restate as R
import abi_emulator as EMU


def T(x):
    return torch.from_numpy(np.asarray(x))


def close(a, b, rtol=1e-5, atol=1e-5):
    torch.testing.assert_close(a.detach().float(), (b if torch.is_tensor(b) else T(b)).float(), rtol=rtol, atol=atol)


class Norm:
    def __init__(self, w, eps):
        self.weight, self.variance_epsilon = w, eps


@pytest.fixture
def emu(monkeypatch):
    return EMU.install(monkeypatch)


# ------------------------------------------------------------------------------------------------
# kernel shims
# ------------------------------------------------------------------------------------------------
def test_rmsnorm_shim(golden, emu):
    from unsloth_b200.kernels import fast_rms_layernorm
    for name in ("rms_llama_512", "rms_llama_odd", "rms_gemma_256"):
        g = golden(name)
        X = T(g["X"]).clone().requires_grad_()
        Y = fast_rms_layernorm(Norm(T(g["W"]), float(g["eps"])), X, gemma=bool(g["gemma"]))
        close(Y, g["Y"])
        dY = T(g["dY"]).clone()
        Y.backward(dY)
        close(X.grad, g["dX"], atol=2e-5)
        if not bool(g["gemma"]):
            assert X.grad.data_ptr() == dY.data_ptr()          # in place over dY (rms_layernorm.py:218)
    assert set(emu) == {"ub200_rms_layernorm_fwd", "ub200_rms_layernorm_bwd"}


def test_rope_shim_in_place_on_strided_views(golden, emu):
    from unsloth_b200.kernels import fast_rope_embedding
    g = golden("rope_noindex")
    Q, K = T(g["Q"]), T(g["K"])
    B, Hq, S, D = Q.shape
    # the layout of the model: Q is a transposed VIEW of the [B, S, H*D] projection buffer
    qbuf = Q.transpose(1, 2).reshape(B, S, Hq * D).clone()
    kbuf = K.transpose(1, 2).reshape(B, S, K.shape[1] * D).clone()
    Qv = qbuf.view(B, S, Hq, D).transpose(1, 2).requires_grad_()
    Kv = kbuf.view(B, S, K.shape[1], D).transpose(1, 2).requires_grad_()
    Qo, Ko = fast_rope_embedding(Qv, Kv, T(g["cos"]), T(g["sin"]))
    close(Qo, g["Qo"]); close(Ko, g["Ko"])
    assert Qo.data_ptr() == qbuf.data_ptr() and Ko.data_ptr() == kbuf.data_ptr()      # rotated in place, no copy
    torch.autograd.backward([Qo, Ko], [T(g["dQ"]).clone(), T(g["dK"]).clone()])
    close(Qv.grad, g["gQ"]); close(Kv.grad, g["gK"])
    g = golden("rope_index")
    Qv, Kv = T(g["Q"]).clone().requires_grad_(), T(g["K"]).clone().requires_grad_()
    Qo, Ko = fast_rope_embedding(Qv, Kv, T(g["cos"]), T(g["sin"]), T(g["idx"]))
    close(Qo, g["Qo"]); close(Ko, g["Ko"])
    torch.autograd.backward([Qo, Ko], [T(g["dQ"]).clone(), T(g["dK"]).clone()])
    close(Qv.grad, g["gQ"]); close(Kv.grad, g["gK"])
    assert emu.count("ub200_rope_qk") == 4                      # Q and K in ONE launch per direction


@pytest.mark.parametrize("name", ["ce_v1000", "ce_v70000_chunked", "ce_softcap30", "ce_scale"])
def test_cross_entropy_shim(golden, emu, name):
    from unsloth_b200.kernels import fast_cross_entropy_loss
    g = golden(name)
    logits = T(g["logits"]).clone().requires_grad_()
    lg = logits * 1.0
    loss = fast_cross_entropy_loss(lg, T(g["labels"]), float(g["softcap"]), float(g["scale"]))
    close(loss, g["loss"])
    loss.backward()
    close(logits.grad, g["dlogits"], atol=1e-6)


def test_glu_shims(golden, emu):
    import unsloth_b200.kernels as K
    for name, fwd, bwd in (("swiglu", K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel),
                           ("geglu_approx", K.geglu_approx_forward_kernel, K.geglu_approx_backward_kernel),
                           ("geglu_exact", K.geglu_exact_forward_kernel, K.geglu_exact_backward_kernel)):
        g = golden(name)
        e, up, DW = T(g["e"]), T(g["g"]), T(g["DW"]).clone()
        close(fwd(e, up), g["h"])
        e2, g2 = e.reshape(-1, e.shape[-1]).clone(), up.reshape(-1, up.shape[-1]).clone()
        h, df, de = bwd(DW, e2, g2)
        close(h, g["bh"]); close(df, g["bdf"]); close(de, g["bde"])
        assert h.data_ptr() == DW.data_ptr() and df.data_ptr() == e2.data_ptr() and de.data_ptr() == g2.data_ptr()


def test_nf4_quantise_dequantise_and_matmul_lora(emu):
    from unsloth_b200.kernels import fast_dequantize, matmul_lora
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(0)
    W = torch.randn(96, 128) * 0.02
    packed, qs = quantize_nf4(W)
    packed_r, qs_r = R.quantize_nf4(W)
    assert torch.equal(packed, packed_r) and torch.equal(qs.absmax, qs_r.absmax)
    D = fast_dequantize(packed, qs)
    assert torch.equal(D, R.dequantize_nf4(packed_r, qs_r))
    assert fast_dequantize(packed.t(), qs).shape == (128, 96)            # transposed contract (utils.py:678)
    assert fast_dequantize(W, None) is W
    X = torch.randn(2, 5, 128)
    A, B = torch.randn(8, 128) * 0.1, torch.randn(96, 8) * 0.1
    out = matmul_lora(X, packed, qs, A, B, 0.5)
    close(out, X @ D.t() + 0.5 * (X @ A.t()) @ B.t(), atol=1e-5)


@pytest.mark.parametrize("softcap,n_items", [(0.0, None), (30.0, 13), (0.0, 7)])
def test_fused_ce_host_logic(emu, softcap, n_items):
    """Chunking, internal label shift, n_items and the gradient wrt the hidden states of the
    logits-free loss against the oracle's `hidden @ W.T -> shift -> CE` (llama.py:1525-1562)."""
    from unsloth_b200.kernels import unsloth_fused_ce_loss
    torch.manual_seed(1)
    Bz, S, H, V = 2, 100, 32, 211            # T = 200 rows -> two 128-row chunks
    hidden = torch.randn(Bz, S, H, requires_grad=True)
    Wt = torch.randn(V, H) * 0.2
    labels = torch.randint(0, V, (Bz, S)); labels[0, 3] = -100; labels[1, 0] = -100
    ref = R.fused_linear_cross_entropy(hidden.detach(), Wt, labels, n_items=n_items, softcap=softcap)
    loss = unsloth_fused_ce_loss(trainer=None, hidden_states=hidden, lm_head_weight=Wt, lm_head_bias=None,
                                 labels=labels, mask=None, n_items=n_items, scaling=None, target_gb=None,
                                 torch_compile=False, logit_softcapping=softcap, chunk_rows=128)
    close(loss, ref[0], atol=2e-5)
    loss.backward()
    close(hidden.grad, ref[1], atol=2e-5)
    h2 = hidden.detach().clone().requires_grad_()
    logits = h2 @ Wt.t()
    if softcap:
        logits = softcap * torch.tanh(logits / softcap)
    shift = torch.full_like(labels, -100); shift[:, :-1] = labels[:, 1:]
    den = n_items if n_items is not None else (shift != -100).sum()
    (torch.nn.functional.cross_entropy(logits.view(-1, V), shift.view(-1), ignore_index=-100, reduction="sum") / den).backward()
    close(hidden.grad, h2.grad, atol=2e-5)


# ------------------------------------------------------------------------------------------------
# whole model: patch.install + fast forwards, against stock HuggingFace (the configs[0] reference)
# ------------------------------------------------------------------------------------------------
TINY = dict(hidden_size=64, intermediate_size=128, num_attention_heads=4, num_key_value_heads=2,
            head_dim=16, vocab_size=256)


def attention_double(Q, K_, V, scale, window, softcap, seq_info=None):
    """The external attention call ([B,S,H,D] in, [B,S,Hq,D] out) in plain torch: causal, optional
    sliding window / soft cap, block-diagonal per document when `seq_info` is given."""
    B, S, Hq, D = Q.shape
    rep = Hq // K_.shape[2]
    q, k, v = (t.float() for t in (Q, K_, V))
    if seq_info is not None:
        q, k, v = (t.reshape(1, B * S, t.shape[2], D) for t in (q, k, v))
    n = q.shape[1]
    q, k, v = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3).repeat_interleave(rep, 1), v.permute(0, 2, 1, 3).repeat_interleave(rep, 1)
    s = (q @ k.transpose(-1, -2)) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    i, j = torch.arange(n)[:, None], torch.arange(n)[None, :]
    mask = j <= i
    if window != (-1, -1):
        mask = mask & (j >= i - window[0])
    if seq_info is not None:
        doc = torch.repeat_interleave(torch.arange(seq_info[0].numel()), seq_info[0].long())
        mask = mask & (doc[:, None] == doc[None, :])
    o = torch.softmax(s.masked_fill(~mask, float("-inf")), -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B, S, Hq, D).to(Q.dtype)


@pytest.fixture
def cpu_model(emu, monkeypatch):
    import unsloth_b200.patch as P
    from unsloth_b200.kernels import utils as KU
    monkeypatch.setattr(P, "_attention", attention_double)
    KU.set_keep_dequant(False)
    yield P
    KU.set_keep_dequant(None)
    P.FUSE_ADD_NORM = True


def _build(P, name, dtype=torch.float32, layers=2, **extra):
    return P.build_qlora_model(name, r=4, lora_alpha=8, device="cpu", dtype=dtype, num_hidden_layers=layers,
                               init_b_std=0.05, **dict(TINY, **extra))


def _grads(P, model):
    return torch.cat([p.grad.float().flatten() for p in P.lora_parameters(model)])


def _zero(P, model):
    for p in P.lora_parameters(model):
        p.grad = None


@pytest.mark.parametrize("name,extra,seq", [("llama-3-8b", {}, 24), ("llama-3.2-1b", {}, 20),
                                            ("mistral-7b-v0.3", {"sliding_window": 8}, 24),
                                            ("gemma-2-9b", {"query_pre_attn_scalar": 16, "sliding_window": 8}, 20)])
def test_patched_model_matches_stock_hf_on_cpu(cpu_model, name, extra, seq):
    """fp32 end to end: NF4 weights dequantised into a stock HF model with plain LoRA (the reference
    CPU path) must give the same loss and the same LoRA gradients as the patched model."""
    from test_gpu_model import _reference_from
    P = cpu_model
    model = _build(P, name, **extra)
    ref_extra = dict(extra)
    if "sliding_window" in ref_extra:
        # the reference hands flash-attn window_size = (sw, sw) (mistral.py:112-128, gemma2.py:139-150):
        # key j is visible iff j >= i - sw, i.e. sw + 1 keys; HF's own mask keeps i - j < sliding_window
        ref_extra["sliding_window"] += 1
    cfg = P.hf_config(name, 2, **dict(TINY, **ref_extra))
    ref = _reference_from(model, cfg)
    torch.manual_seed(1)
    ids = torch.randint(0, TINY["vocab_size"], (2, seq))
    labels = ids.clone(); labels[0, :3] = -100
    out = model(input_ids=ids, labels=labels)
    out.loss.backward()
    ref_out = ref(input_ids=ids, labels=labels)
    ref_out.loss.backward()
    assert abs(out.loss.item() - ref_out.loss.item()) <= 2e-4 * abs(ref_out.loss.item())
    from unsloth_b200.kernels import get_lora_parameters
    for lo, lr in zip(model.model.layers, ref.model.layers):
        for po, pr in ((lo.self_attn, lr.self_attn), (lo.mlp, lr.mlp)):
            for pn in P.TARGET_MODULES:
                if not hasattr(po, pn):
                    continue
                _, _, A, B, _ = get_lora_parameters(getattr(po, pn))
                close(A.grad, getattr(pr, pn).A.grad, rtol=2e-3, atol=2e-6)
                close(B.grad, getattr(pr, pn).B.grad, rtol=2e-3, atol=2e-6)


def test_fused_add_norm_chain_on_cpu(cpu_model):
    """bf16: the (residual, normed) chain of fast_add_rms_layernorm against the layer-by-layer form:
    identical forward, gradients equal up to the one rounding the fusion removes."""
    P = cpu_model
    model = _build(P, "llama-3-8b", dtype=torch.bfloat16, layers=3)
    torch.manual_seed(2)
    ids = torch.randint(0, TINY["vocab_size"], (2, 16))
    res = {}
    for fuse in (True, False):
        P.FUSE_ADD_NORM = fuse
        _zero(P, model)
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        res[fuse] = (loss.item(), _grads(P, model))
    assert res[True][0] == res[False][0]
    a, b = res[True][1], res[False][1]
    assert torch.dot(a, b) / (a.norm() * b.norm()) > 0.999


def test_packed_row_on_cpu(cpu_model):
    P = cpu_model
    model = _build(P, "mistral-7b-v0.3", sliding_window=6)
    torch.manual_seed(3)
    L1, L2 = 9, 14
    d1 = torch.randint(0, TINY["vocab_size"], (1, L1)); d2 = torch.randint(0, TINY["vocab_size"], (1, L2))
    row = torch.cat([d1, d2], 1)
    lens = torch.tensor([L1, L2], dtype=torch.int32)
    with torch.no_grad():
        hp = model(input_ids=row, packed_seq_lengths=lens).hidden_states
        h1, h2 = model(input_ids=d1).hidden_states, model(input_ids=d2).hidden_states
        leaked = model(input_ids=row).hidden_states
    close(hp[:, :L1], h1, rtol=1e-4, atol=1e-5); close(hp[:, L1:], h2, rtol=1e-4, atol=1e-5)
    assert (leaked[:, L1:] - h2).abs().max() > 1e-2 * h2.abs().max()
    n = L1 + L2 - 2
    _zero(P, model)
    lp = model(input_ids=row, labels=row, packed_seq_lengths=lens, num_items_in_batch=n).loss
    lp.backward()
    gp = _grads(P, model)
    _zero(P, model)
    ls = 0.0
    for d in (d1, d2):
        l = model(input_ids=d, labels=d, num_items_in_batch=n).loss
        l.backward()
        ls += l.item()
    assert abs(lp.item() - ls) <= 1e-4 * abs(ls)
    close(gp, _grads(P, model), rtol=2e-3, atol=1e-6)
    from unsloth_b200.packing import num_items_in_batch
    assert num_items_in_batch(row, lens) == n


def test_return_logits_path_on_cpu(cpu_model, monkeypatch):
    """UNSLOTH_RETURN_LOGITS=1 (llama.py:1525-1562) == the logits-free default, incl. packed rows."""
    P = cpu_model
    for name, extra in (("llama-3-8b", {}), ("gemma-2-9b", {"query_pre_attn_scalar": 16})):
        model = _build(P, name, **extra)
        torch.manual_seed(4)
        ids = torch.randint(0, TINY["vocab_size"], (1, 18))
        labels = ids.clone(); labels[0, :4] = -100
        lens = torch.tensor([7, 11], dtype=torch.int32)
        res = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("UNSLOTH_RETURN_LOGITS", flag)
            _zero(P, model)
            out = model(input_ids=ids, labels=labels, packed_seq_lengths=lens)
            out.loss.backward()
            res[flag] = (out.loss.item(), _grads(P, model), out.logits)
        assert res["0"][2] is None and res["1"][2].shape == (1, 18, TINY["vocab_size"])
        assert abs(res["0"][0] - res["1"][0]) <= 1e-5 * abs(res["0"][0])
        close(res["1"][1], res["0"][1], rtol=1e-3, atol=1e-6)


def test_keep_dequant_on_cpu(cpu_model):
    from unsloth_b200.kernels import utils as KU
    P = cpu_model
    model = _build(P, "llama-3-8b")
    ids = torch.randint(0, TINY["vocab_size"], (2, 12), generator=torch.Generator().manual_seed(5))
    res = {}
    import unsloth_b200._lib as L
    for keep in (True, False):
        KU.set_keep_dequant(keep)
        KU.bump_param_epoch()
        _zero(P, model)
        n0 = L.launch_count
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        res[keep] = (loss.item(), _grads(P, model), L.launch_count - n0)
    assert res[True][0] == res[False][0] and torch.equal(res[True][1], res[False][1])
    assert res[False][2] - res[True][2] == 2 * 7 - 3        # q/k/v of layer 0 form no dX


def test_decode_path_host_logic(emu):
    """fast_gemv / fast_linear_forward (kernels/utils.py:874-973, :1082-1125) and merge_lora
    (save.py:620-646): branch selection, LoRA temp, epilogue arguments, bias, `out=` reuse."""
    from unsloth_b200.kernels import fast_gemv, fast_linear_forward, fast_dequantize, get_lora_parameters
    from unsloth_b200.lora import LoraLinear
    from unsloth_b200.nf4 import Linear4bit
    from unsloth_b200.save import merge_lora
    torch.manual_seed(8)
    k, m = 128, 96
    W = (torch.randn(m, k) * 0.05).to(torch.bfloat16)
    proj = LoraLinear(Linear4bit.from_dense(W), r=8, lora_alpha=16, init_b_std=0.1)
    Wq, qs, A, B, s = get_lora_parameters(proj)
    Wd = fast_dequantize(Wq, qs).float()
    full = Wd + s * (B.float() @ A.float())
    x = torch.randn(1, 1, k).to(torch.bfloat16)
    emu.clear()
    y = fast_gemv(x, Wq, qs)
    assert emu == ["ub200_gemv_nf4"] and y.shape == (1, 1, m)                       # ONE launch
    close(y.view(-1), Wd @ x.view(-1).float(), rtol=2e-2, atol=2e-2)
    out = torch.empty(1, 1, m, dtype=torch.bfloat16)
    assert fast_gemv(x, Wq, qs, out=out).data_ptr() == out.data_ptr()
    emu.clear()
    y1 = fast_linear_forward(proj, x)
    assert emu == ["ub200_gemv_dense", "ub200_gemv_nf4"]                            # A x, then GEMV + epilogue
    close(y1.view(-1), full @ x.view(-1).float(), rtol=2e-2, atol=3e-2)
    ref = R.fast_linear_forward(x, Wq, qs, A.detach(), B.detach(), s)
    close(y1.view(-1), ref, rtol=8e-3, atol=8e-3)
    # q_len != 1 falls through to the training primitive; bsz > 1 uses the GEMM on [bsz, in] rows
    x3 = torch.randn(1, 3, k).to(torch.bfloat16)
    close(fast_linear_forward(proj, x3).view(3, m), x3.view(3, k).float() @ full.t(), rtol=2e-2, atol=3e-2)
    x4 = torch.randn(4, 1, k).to(torch.bfloat16)
    y4 = fast_linear_forward(proj, x4)
    assert y4.shape == (4, 1, m)
    close(y4.view(4, m), x4.view(4, k).float() @ full.t(), rtol=2e-2, atol=3e-2)
    # dense 16-bit base with a bias
    base = torch.nn.Linear(k, m, bias=True, dtype=torch.bfloat16)
    dense = LoraLinear(base, r=8, lora_alpha=16, init_b_std=0.1)
    Wn, _, A2, B2, s2 = get_lora_parameters(dense)
    yd = fast_linear_forward(dense, x)
    close(yd.view(-1), (Wn.float() + s2 * (B2.float() @ A2.float())) @ x.view(-1).float() + base.bias.float(),
          rtol=2e-2, atol=3e-2)
    # merge: W + s B A formed in fp32, rounded once
    Wm, bias = merge_lora(proj, "proj")
    assert bias is None and torch.equal(Wm, full.to(torch.bfloat16))
    proj.disable_adapters = True
    assert torch.equal(merge_lora(proj)[0], Wd.to(torch.bfloat16))


@pytest.mark.parametrize("name,dtype,fuse,packed", [
    ("llama-3-8b", torch.bfloat16, True, False), ("llama-3-8b", torch.float32, False, False),
    ("gemma-2-9b", torch.float32, True, False), ("mistral-7b-v0.3", torch.bfloat16, True, True)])
def test_gradient_checkpointing_recompute_is_exact(cpu_model, name, dtype, fuse, packed):
    """Per-layer recompute in the backward (install(..., gradient_checkpointing=True)) must give the
    same loss and bit-identical LoRA gradients: the in-place kernels (RoPE on the projection
    buffers, dX into the saved X, dY -> dX of RMSNorm, the residual-gradient accumulation) have to
    survive the second forward.  The emulator reproduces the in-place memory semantics."""
    P = cpu_model
    extra = {"query_pre_attn_scalar": 16} if name.startswith("gemma") else {}
    model = _build(P, name, dtype=dtype, layers=3, **extra)
    P.FUSE_ADD_NORM = fuse
    torch.manual_seed(6)
    ids = torch.randint(0, TINY["vocab_size"], (1 if packed else 2, 20))
    kw = dict(packed_seq_lengths=torch.tensor([8, 12], dtype=torch.int32)) if packed else {}
    res = {}
    for ck in (False, True):
        model.model._ub_gradient_checkpointing = ck
        _zero(P, model)
        loss = model(input_ids=ids, labels=ids, **kw).loss
        loss.backward()
        res[ck] = (loss.item(), _grads(P, model))
    assert res[True][0] == res[False][0]
    assert torch.equal(res[True][1], res[False][1])
    assert res[True][1].abs().max() > 0


def test_class_level_routes_serve_a_stock_hf_model(emu):
    """SURVEY 8b, class-level alternative: `patch_rms_layernorm` (rms_layernorm.py:261-274) and
    `patch_loss_functions` (cross_entropy_loss.py:459-473) let the kernels serve a STOCK HF Llama
    (no fast forwards): same loss and gradients as the unpatched model."""
    import unsloth_b200.kernels as K
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_attention_heads=4, num_key_value_heads=2,
                      num_hidden_layers=2, vocab_size=256, max_position_embeddings=64)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    ids = torch.randint(0, 256, (2, 14))
    labels = ids.clone(); labels[0, :4] = -100
    base = LlamaForCausalLM(cfg).float()
    sd = copy.deepcopy(base.state_dict())
    out0 = base(input_ids=ids, labels=labels)
    out0.loss.backward()
    g0 = base.model.layers[0].mlp.down_proj.weight.grad.clone()
    K.patch_rms_layernorm(); K.patch_loss_functions()
    try:
        patched = LlamaForCausalLM(cfg).float()              # built AFTER the class swap
        patched.load_state_dict(sd)
        emu.clear()
        out1 = patched(input_ids=ids, labels=labels)
        out1.loss.backward()
        assert "ub200_rms_layernorm_fwd" in emu and "ub200_cross_entropy_fwd" in emu
        assert "ub200_rms_layernorm_bwd" in emu and "ub200_cross_entropy_bwd" in emu
    finally:
        K.unpatch_rms_layernorm(); K.unpatch_loss_functions()
    close(out1.loss, out0.loss, rtol=1e-5, atol=1e-6)
    close(patched.model.layers[0].mlp.down_proj.weight.grad, g0, rtol=1e-4, atol=1e-7)
    import transformers.loss.loss_utils as lu
    assert lu.LOSS_MAPPING["ForCausalLM"].__name__ == "ForCausalLMLoss"             # restored


def test_rope_promotion_and_single_tensor_forms(emu):
    """Gemma-style fp32 tables with 16-bit Q/K: the no-index form computes in the TABLE dtype
    (rope_embedding.py:154), so fp32 tables give fp32 products and one final rounding;
    `Fast_RoPE_Embedding` (single [B,S,H,D] tensor, :169-261) shares the kernel, in place."""
    from unsloth_b200.kernels import fast_rope_embedding, Fast_RoPE_Embedding
    torch.manual_seed(0)
    B, H, S, D = 2, 3, 7, 16
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(16).float(), inv)
    cos, sin = torch.cat([fr, fr], -1).cos(), torch.cat([fr, fr], -1).sin()
    Q = torch.randn(B, H, S, D).to(torch.bfloat16); K = torch.randn(B, 1, S, D).to(torch.bfloat16)
    Qo, Ko = fast_rope_embedding(Q.clone(), K.clone(), cos, sin)
    ref = R.rope_noindex(Q.transpose(1, 2), cos, sin).transpose(1, 2)
    assert Qo.dtype == torch.bfloat16 and torch.equal(Qo, ref.to(torch.bfloat16))
    # position ids == arange must reproduce the no-index result when everything is fp32
    Qf, Kf = Q.float(), K.float()
    a, _ = fast_rope_embedding(Qf.clone(), Kf.clone(), cos, sin)
    b, _ = fast_rope_embedding(Qf.clone(), Kf.clone(), cos, sin, torch.arange(S).repeat(B))
    close(a, b)
    # single-tensor API on [B, S, H, D]
    X0 = torch.randn(B, S, H, D)
    X = X0.clone().requires_grad_()
    Y = Fast_RoPE_Embedding.apply(X, cos, sin)                  # rotates X's storage in place
    close(Y, R.rope_noindex(X0, cos, sin))
    assert Y.data_ptr() == X.data_ptr()
    Y.backward(torch.ones_like(Y))
    close(X.grad, R.rope_noindex(torch.ones(B, S, H, D), cos, sin, backward=True))


def test_cross_entropy_argument_forms(emu):
    """`n_items` as int / tensor / None, `logit_scaling`, all-ignored rows (loss 0, zero gradient)."""
    from unsloth_b200.kernels import fast_cross_entropy_loss
    torch.manual_seed(3)
    logits = torch.randn(2, 5, 50)
    labels = torch.randint(0, 50, (2, 5)); labels[0] = -100
    for n_items in (None, 5, torch.tensor(5)):
        lg = logits.clone().requires_grad_()
        loss = fast_cross_entropy_loss(lg * 1.0, labels, 0, 0.5, n_items=n_items)
        ref = torch.nn.functional.cross_entropy((0.5 * logits).view(-1, 50), labels.view(-1), ignore_index=-100,
                                                reduction="sum") / 5
        close(loss, ref)
        loss.backward()
        assert lg.grad[0].abs().max() == 0
    lg = logits.clone().requires_grad_()
    all_ignored = torch.full((2, 5), -100)
    loss = fast_cross_entropy_loss(lg * 1.0, all_ignored, n_items=1)
    assert loss.item() == 0
    loss.backward()
    assert lg.grad.abs().max() == 0


def test_fp16_model_runs_through_the_16bit_paths(cpu_model):
    """fp16 end to end (the reference's T4 CI dtype): fused add+norm, packed RoPE tables in fp16,
    loss finite and close to the fp32 model built from the same seed."""
    P = cpu_model
    m16 = _build(P, "llama-3-8b", dtype=torch.float16)
    m32 = _build(P, "llama-3-8b", dtype=torch.float32)
    ids = torch.randint(0, TINY["vocab_size"], (2, 16), generator=torch.Generator().manual_seed(1))
    l16 = m16(input_ids=ids, labels=ids).loss
    l32 = m32(input_ids=ids, labels=ids).loss
    l16.backward()
    assert torch.isfinite(l16) and abs(l16.item() - l32.item()) < 2e-2 * abs(l32.item())
    assert all(torch.isfinite(p.grad).all() for p in P.lora_parameters(m16))


# ------------------------------------------------------------------------------------------------
# round-2 advisor findings
# ------------------------------------------------------------------------------------------------
def test_lora_casts_never_go_stale_after_raw_parameter_updates(cpu_model):
    """Optimisers that write parameters behind autograd's back (`p.data.add_()`, raw pointers,
    8-bit / fused optimisers) bump neither `_version` nor PARAM_EPOCH: every forward rebuilds its
    16-bit adapter casts, so the next step must see the new adapters."""
    P = cpu_model
    model = _build(P, "llama-3-8b")
    torch.manual_seed(2)
    ids = torch.randint(0, TINY["vocab_size"], (1, 12))
    l0 = model(input_ids=ids, labels=ids).loss.item()
    l0b = model(input_ids=ids, labels=ids).loss.item()
    assert l0 == l0b
    with torch.no_grad():
        for p in P.lora_parameters(model):
            p.data.add_(0.05 * torch.randn_like(p))          # no _version bump on `p`
    l1 = model(input_ids=ids, labels=ids).loss.item()
    assert abs(l1 - l0) > 1e-4, (l0, l1)
    # the decode-time cast cache is refreshed by ANY torch.optim step (global post-step hook)
    from unsloth_b200.kernels import utils as KU
    e0 = KU.PARAM_EPOCH
    opt = torch.optim.SGD(P.lora_parameters(model), lr=0.1)
    model(input_ids=ids, labels=ids).loss.backward()
    opt.step()
    assert KU.PARAM_EPOCH > e0


def test_fused_ce_trainable_lm_head_bias_and_loss_scaler(emu):
    """unsloth_fused_ce_loss with a trainable lm_head (+ bias), a vocabulary that is not a multiple
    of 8, and the fp16 GradScaler the reference call site passes as `scaling`
    (models/llama.py:1505): same loss and gradients as autograd on the logits path; the scale is
    folded into the in-pass gradient and divided out again."""
    from unsloth_b200.kernels import unsloth_fused_ce_loss
    torch.manual_seed(4)
    Bz, S, H, V = 2, 70, 32, 203
    hidden = torch.randn(Bz, S, H, requires_grad=True)
    Wt = (torch.randn(V, H) * 0.2).requires_grad_()
    bias = (torch.randn(V) * 0.1).requires_grad_()
    labels = torch.randint(0, V, (Bz, S)); labels[0, 3] = -100

    class Scaler:                                    # GradScaler surface used by the loss
        def get_scale(self): return 1024.0
        def is_enabled(self): return True
    loss = unsloth_fused_ce_loss(trainer=None, hidden_states=hidden, lm_head_weight=Wt, lm_head_bias=bias,
                                 labels=labels, mask=None, n_items=None, scaling=Scaler(), target_gb=None,
                                 torch_compile=False, logit_softcapping=0, chunk_rows=128)
    (loss * 3.0).backward()                          # an upstream factor must flow through unchanged
    h2, W2, b2 = (t.detach().clone().requires_grad_() for t in (hidden, Wt, bias))
    logits = h2 @ W2.t() + b2
    shift = torch.full_like(labels, -100); shift[:, :-1] = labels[:, 1:]
    ref = torch.nn.functional.cross_entropy(logits.view(-1, V), shift.view(-1), ignore_index=-100)
    (ref * 3.0).backward()
    close(loss, ref, atol=2e-5)
    close(hidden.grad, h2.grad, atol=2e-5)
    close(Wt.grad, W2.grad, atol=2e-5)
    close(bias.grad, b2.grad, atol=2e-5)
    # frozen lm_head (the QLoRA default): no weight-gradient work, no gradient
    Wf = Wt.detach().clone()
    h3 = hidden.detach().clone().requires_grad_()
    unsloth_fused_ce_loss(hidden_states=h3, lm_head_weight=Wf, labels=labels, chunk_rows=128).backward()
    h4 = hidden.detach().clone().requires_grad_()
    torch.nn.functional.cross_entropy((h4 @ Wf.t()).view(-1, V), shift.view(-1), ignore_index=-100).backward()
    close(h3.grad, h4.grad, atol=2e-5)
    # no_grad: loss only
    with torch.no_grad():
        l = unsloth_fused_ce_loss(hidden_states=hidden, lm_head_weight=Wf, labels=labels, chunk_rows=128)
    assert not l.requires_grad


def test_rope_scaling_types():
    import unsloth_b200.patch as P
    base = P.RotaryCache(64, 10000.0, "cpu", torch.float32).inv_freq()
    lin = P.RotaryCache(64, 10000.0, "cpu", torch.float32, dict(rope_type="linear", factor=4.0)).inv_freq()
    assert torch.allclose(lin, base / 4.0)
    with pytest.raises(NotImplementedError):
        P.RotaryCache(64, 10000.0, "cpu", torch.float32, dict(rope_type="yarn", factor=4.0)).inv_freq()


def test_no_grad_forward_does_not_keep_dequantised_weights(cpu_model):
    from unsloth_b200.kernels import utils as KU
    P = cpu_model
    model = _build(P, "llama-3-8b")
    KU.set_keep_dequant(True)
    try:
        ids = torch.randint(0, TINY["vocab_size"], (1, 8))
        seen = []
        orig = KU.fast_dequantize

        def spy(W, qs=None, out=None, use_global_buffer=False, _slot=0):
            seen.append(use_global_buffer)
            return orig(W, qs, out, use_global_buffer, _slot)
        KU.fast_dequantize = spy
        try:
            with torch.no_grad():
                model(input_ids=ids, labels=ids)
            assert seen and all(seen), "no_grad forward must reuse the global buffers"
            seen.clear()
            model(input_ids=ids, labels=ids).loss.backward()
            assert not any(seen), "training forward keeps private expansions for the backward"
        finally:
            KU.fast_dequantize = orig
    finally:
        KU.set_keep_dequant(False)


@pytest.mark.parametrize("name", ["llama-3-8b", "gemma-2-9b"])
def test_offloaded_checkpoint_and_tiled_mlp_are_exact(cpu_model, name):
    """`gradient_checkpointing="unsloth"` (layer inputs parked on the host, layer recomputed in the
    backward) and `tiled_mlp=n` (MLP over token shards, recomputed shard-wise) change memory, not
    arithmetic: same loss, same LoRA gradients as the plain step (fp32 through the ABI emulator)."""
    P = cpu_model
    extra = {"query_pre_attn_scalar": 16} if name == "gemma-2-9b" else {}
    torch.manual_seed(3)
    ids = torch.randint(0, TINY["vocab_size"], (2, 600))          # >= 512 tokens: offloading engages
    labels = ids.clone(); labels[0, :7] = -100
    base = _build(P, name, **extra)
    out = base(input_ids=ids, labels=labels)
    out.loss.backward()
    g0 = _grads(P, base)
    for kw in (dict(gradient_checkpointing="unsloth"), dict(tiled_mlp=3), dict(gradient_checkpointing=True, tiled_mlp=2)):
        m = P.build_qlora_model(name, r=4, lora_alpha=8, device="cpu", dtype=torch.float32, num_hidden_layers=2,
                                init_b_std=0.05, **dict(TINY, **extra), **kw)
        o = m(input_ids=ids, labels=labels)
        o.loss.backward()
        assert abs(o.loss.item() - out.loss.item()) <= 1e-6 * abs(out.loss.item()), kw
        torch.testing.assert_close(_grads(P, m), g0, rtol=1e-5, atol=1e-7)
    from unsloth_b200.kernels import utils as KU
    KU.KEEP_DEQUANT_BLOCKED = False


@pytest.mark.parametrize("name,extra", [("llama-3-8b", {}), ("mistral-7b-v0.3", {"sliding_window": 6}),
                                        ("gemma-2-9b", {"query_pre_attn_scalar": 16, "sliding_window": 6})])
def test_kv_cache_decode_loop_matches_full_forward(cpu_model, name, extra):
    """generate(): prefill + single-token KV-cache steps (llama.py:352-602) reproduce what the full
    forward over the growing sequence predicts (greedy tokens identical, fp32 through the emulator)."""
    from unsloth_b200.generate import generate
    P = cpu_model
    model = _build(P, name, **extra)
    torch.manual_seed(5)
    ids = torch.randint(0, TINY["vocab_size"], (1, 9))
    out = generate(model, ids, max_new_tokens=6)
    assert out.shape == (1, 15) and torch.equal(out[:, :9], ids)
    cur = ids
    with torch.no_grad():
        for _ in range(6):
            hidden = P.Model_fast_forward(model.model, cur)
            logits = torch.nn.functional.linear(hidden[:, -1:], model.lm_head.weight).float()
            if model._ub_final_softcap:
                logits = model._ub_final_softcap * torch.tanh(logits / model._ub_final_softcap)
            cur = torch.cat([cur, logits[:, -1].argmax(-1, keepdim=True)], 1)
    assert torch.equal(out, cur)


@pytest.mark.parametrize("name,extra", [("llama-3-8b", {}), ("gemma-2-9b", {"query_pre_attn_scalar": 16, "sliding_window": 8})])
def test_step_plan_batches_casts_and_gradient_accumulation(cpu_model, emu, name, extra):
    """FlatLoRABucket.begin_step() / end_backward() (kernels/utils.py::StepPlan): from the second step on the
    per-adapter cast launches collapse into ub200_cast_pad_multi and the d_A / d_B of every projection reach the
    flat gradient bucket through ub200_accumulate_multi instead of autograd's AccumulateGrad -- with the SAME
    losses and gradients as the plain autograd route, also after a raw in-place parameter update between steps."""
    from unsloth_b200.ddp import FlatLoRABucket
    from unsloth_b200.kernels import utils as KU
    P = cpu_model
    torch.manual_seed(11)
    model = _build(P, name, **extra)
    torch.manual_seed(11)
    ref = _build(P, name, **extra)
    ids = torch.randint(0, TINY["vocab_size"], (2, 16))
    bucket = FlatLoRABucket(P.lora_parameters(model))
    n_params = len(bucket.params)
    for step in range(3):
        # reference route: autograd accumulates each gradient on its own
        _zero(P, ref)
        lr_ = ref(input_ids=ids, labels=ids).loss
        lr_.backward()
        g_ref = _grads(P, ref)
        del emu[:]
        bucket.begin_step()
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        bucket.end_backward()
        assert KU.ACTIVE_PLAN is None and not bucket.plan.in_step
        assert loss.item() == lr_.item()
        g = torch.cat([p.grad.flatten() for p in P.lora_parameters(model)])
        assert torch.equal(g, g_ref), (step, (g - g_ref).abs().max())
        assert emu.count("ub200_accumulate_multi") == 1 and len(bucket.plan.grad_pairs) == 0
        if step == 0:
            assert "ub200_cast_pad_2d" in emu and "ub200_cast_pad_multi" not in emu       # the recording step
            n_casts = sum(len(e[4]) for e in bucket.plan.entries.values())
            assert n_casts >= n_params                              # every adapter tensor is cast at least once
        else:
            assert "ub200_cast_pad_2d" not in emu and emu.count("ub200_cast_pad_multi") == 1
            assert sum(len(e[4]) for e in bucket.plan.entries.values()) == n_casts
        # a raw in-place update on BOTH copies (no _version bump, no optimiser hook): the next step must see it
        with torch.no_grad():
            torch.manual_seed(100 + step)
            for p, q in zip(P.lora_parameters(model), P.lora_parameters(ref)):
                d = 0.05 * torch.randn_like(p)
                p.data.add_(d); q.data.add_(d)
    # outside a step nothing is taken from the plan: a plain forward rebuilds its casts
    del emu[:]
    model(input_ids=ids, labels=ids)
    assert "ub200_cast_pad_2d" in emu and "ub200_cast_pad_multi" not in emu


@pytest.mark.parametrize("name,dtype", [("llama-3-8b", torch.bfloat16), ("gemma-2-9b", torch.float16)])
def test_glu_epilogue_schedule_matches_two_launch_schedule(cpu_model, emu, monkeypatch, name, dtype):
    """Host logic of the fused gated activation (ub200_gemm_glu, default for 16-bit tensors): the up
    projection's launch leaves g and h, the DW launch overwrites e / g with df / de in place.  Same loss and
    the same LoRA gradients as the two-launch schedule (UB200_FUSED_GLU=0), and the fused entry point is
    what actually ran."""
    P = cpu_model
    extra = {"query_pre_attn_scalar": 16, "sliding_window": 8} if name.startswith("gemma") else {}
    ids = torch.randint(0, TINY["vocab_size"], (2, 12), generator=torch.Generator().manual_seed(3))
    res = {}
    for flag in ("1", "0", "fwd", "bwd"):
        monkeypatch.setenv("UB200_FUSED_GLU", flag)
        torch.manual_seed(0)
        m = _build(P, name, dtype=dtype, **extra)
        del emu[:]
        loss = m(input_ids=ids, labels=ids).loss
        loss.backward()
        n_fused = sum(1 for c in emu if c == "ub200_gemm_glu")
        n_plain = sum(1 for c in emu if c in ("ub200_glu_fwd", "ub200_glu_bwd"))
        res[flag] = (loss.detach().float(), _grads(P, m), n_fused, n_plain)
    layers = 2
    assert res["1"][2:] == (2 * layers, 0) and res["0"][2:] == (0, 2 * layers)
    assert res["fwd"][2:] == (layers, layers) and res["bwd"][2:] == (layers, layers)
    for flag in ("0", "fwd", "bwd"):
        assert torch.equal(res["1"][0], res[flag][0])
        torch.testing.assert_close(res["1"][1], res[flag][1], rtol=0, atol=0)


def test_glu_epilogue_falls_back_when_the_kernel_cannot_take_the_shape(cpu_model, emu, monkeypatch):
    """ub200_gemm_glu needs 16-bit tensors and whole tiles (N % 64 == 0 for N <= 64, % 128 up to 128, % 256
    above).  An intermediate size of 96, and an fp32 model, must take the two-launch route on their own -- same
    numbers as with the epilogues switched off, no fused call issued."""
    P = cpu_model
    ids = torch.randint(0, TINY["vocab_size"], (1, 10), generator=torch.Generator().manual_seed(4))
    for dtype, extra in ((torch.bfloat16, {"intermediate_size": 96}), (torch.float32, {})):
        res = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("UB200_FUSED_GLU", flag)
            torch.manual_seed(0)
            m = _build(P, "llama-3-8b", dtype=dtype, **extra)
            del emu[:]
            loss = m(input_ids=ids, labels=ids).loss
            loss.backward()
            assert "ub200_gemm_glu" not in emu
            assert emu.count("ub200_glu_fwd") == 2 and emu.count("ub200_glu_bwd") == 2
            res[flag] = (loss.detach().float(), _grads(P, m))
        assert torch.equal(res["1"][0], res["0"][0]) and torch.equal(res["1"][1], res["0"][1])
