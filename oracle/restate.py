"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the QLoRA hot path.

A plain-PyTorch (device='cpu') restatement of every function on the hot path of
unslothai/unsloth, written from the reference's algorithm and
mirroring its rounding points.  Each function cites the reference
file:line it follows (paths relative to /root/reference/unsloth/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module -- and only as the checker.  The product path (unsloth_b200/*)
never imports it and fails loudly when the CUDA library is missing.

Parity pinning status
  * RMSNorm / RoPE / SwiGLU / GEGLU / CE / LoRA_MLP / LoRA_QKV / LoRA_W: pinned against
    the reference's own Triton kernels executed under TRITON_INTERPRET=1 (fp32) through
    oracle/ref_shim.py; the outputs are committed as tests/golden/*.npz by
    oracle/make_golden.py and re-checked in tests/test_oracle_golden.py.  The 16-bit ROUNDING
    POINTS are pinned too: fp16 runs of the same reference kernels and
    LoRA functions (tests/golden/fp16_*.npz) are reproduced to the ulp (LoRA: bit for bit).
  * NF4 double-quant dequantisation (bitsandbytes >=0.45.5, not vendored, not installed):
    restated from its published algorithm and the reference call site
    kernels/utils.py:582-598, 650-675; likewise the 4-bit GEMV of `fast_gemv` (:874-973).
    **parity unpinned**.
  * Logits-free fused linear cross-entropy (unsloth_zoo >= 2026.8.13, not vendored, not
    installed): semantics restated from the logits path it replaces
    (models/llama.py:1525-1562 + kernels/cross_entropy_loss.py:421-449).
    **parity unpinned**.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

# --------------------------------------------------------------------------------------
# RMSNorm  (kernels/rms_layernorm.py)
# --------------------------------------------------------------------------------------


def rms_layernorm_fwd(X, W, eps, gemma=False):
    """kernels/rms_layernorm.py:21-59 (Llama) and :123-159 (Gemma).

    Returns (Y in X.dtype, r fp32[T]).  Llama form rounds `normed` to W.dtype before the
    multiply by W (:57); the Gemma form is all-fp32 with (W + 1).
    """
    shape = X.shape
    X2 = X.reshape(-1, shape[-1])
    Xf = X2.float()
    var = (Xf * Xf).sum(-1) / X2.shape[-1]
    r = torch.rsqrt(var + torch.tensor(eps, dtype=torch.float32))
    normed = Xf * r[:, None]
    if gemma:
        Y = (normed * (W.float() + 1.0)).to(X.dtype)
    else:
        Y = (normed.to(W.dtype) * W).to(X.dtype)
    return Y.reshape(shape), r


def rms_layernorm_bwd(dY, X, W, r, gemma=False):
    """kernels/rms_layernorm.py:62-120.  dX only (norm weights are frozen)."""
    shape = dY.shape
    n = shape[-1]
    dYf = dY.reshape(-1, n).float()
    Xf = X.reshape(-1, n).float()
    Wf = W.float()
    normed = Xf * r[:, None].float()
    dY_W = dYf * (Wf + 1.0) if gemma else dYf * Wf
    rowsum = (dY_W * normed).sum(-1, keepdim=True)
    dX = r[:, None] / n * (n * dY_W - normed * rowsum)
    return dX.to(dY.dtype).reshape(shape)


# --------------------------------------------------------------------------------------
# RoPE  (kernels/rope_embedding.py)
# --------------------------------------------------------------------------------------


# How the reference's `q0*cos - q1*sin` / `q1*cos + q0*sin` round (rope_embedding.py:82-83, 162-163):
#   False  every product and the sum round to the operand dtype -- what TRITON_INTERPRET=1 (numpy)
#          does; pins the fp32 / fp16 golden vectors generated in the build container;
#   True   what Triton 3.6 emits NATIVELY on sm_100 (PTX read off the GPU box, profiles/
#          r2_triton_rope_ptx.txt): `fma(q0, cos, -rn(q1*sin))` and `fma(q0, sin, rn(q1*cos))`,
#          i.e. one product rounded and the other fused.  The GPU parity tests use this form.
ROPE_NATIVE_CONTRACTION = False


def _rot(q1, q2, cos, sin, contract=None):
    contract = ROPE_NATIVE_CONTRACTION if contract is None else contract
    dt = torch.promote_types(torch.promote_types(q1.dtype, cos.dtype), q2.dtype)
    if not contract or dt == torch.float32:
        # products and sums each round to the operand dtype, exactly as separate tensor ops would
        return q1 * cos - q2 * sin, q2 * cos + q1 * sin
    # 16-bit operands: a product of two 8/11-bit significands is exact in fp32
    f = torch.float32
    m_s = (q2 * sin).to(f)                      # rn16(q1*sin)
    m_c = (q2 * cos).to(f)                      # rn16(q1*cos)
    o1 = (q1.to(f) * cos.to(f) - m_s).to(dt)
    o2 = (q1.to(f) * sin.to(f) + m_c).to(dt)
    return o1, o2


def rope_noindex(Q, cos, sin, backward=False, contract=None):
    """kernels/rope_embedding.py:104-166 (+ wrapper :169-261).

    Q: [B, S, n_heads, D] (contiguous).  cos/sin: [>=S, D]; only the first D/2 columns
    are read (:118-127).  Math is done in the *table* dtype (:154-155), result stored in
    Q.dtype.  Returns a new tensor (the reference works in place).
    """
    B, S, H, D = Q.shape
    half = D // 2
    cos1 = cos[:S, :half][None, :, None, :]
    sin1 = sin[:S, :half][None, :, None, :]
    if backward:
        sin1 = -sin1
    q1 = Q[..., :half].to(cos.dtype)
    q2 = Q[..., half:].to(cos.dtype)
    o1, o2 = _rot(q1, q2, cos1, sin1, contract)
    return torch.cat([o1, o2], dim=-1).to(Q.dtype)


def rope_qk(Q, K, cos, sin, indices=None, backward=False, contract=None):
    """kernels/rope_embedding.py:23-98 (+ wrapper :283-399).

    Q: [B, Hq, S, D], K: [B, Hk, S, D].  Row of cos/sin is `indices[b*S+s]` when given,
    else s (:51-58).  Operands keep their own dtype (type promotion with the table).
    """
    B, Hq, S, D = Q.shape
    half = D // 2
    if indices is not None:
        pos = indices.reshape(-1).long()
    else:
        pos = torch.arange(S).repeat(B)
    cos1 = cos[pos, :half].reshape(B, 1, S, half)
    sin1 = sin[pos, :half].reshape(B, 1, S, half)
    if backward:
        sin1 = -sin1

    def one(T):
        t1, t2 = T[..., :half], T[..., half:]
        o1, o2 = _rot(t1, t2, cos1, sin1, contract)
        return torch.cat([o1, o2], dim=-1).to(T.dtype)

    return one(Q), one(K)


# --------------------------------------------------------------------------------------
# SwiGLU / GEGLU  (kernels/swiglu.py, kernels/geglu.py)
# --------------------------------------------------------------------------------------


def swiglu_fwd(e, g):
    """kernels/swiglu.py:27-47: h = (e*sigmoid(e)).to(g.dtype) * g."""
    ef = e.float()
    f = (ef * torch.sigmoid(ef)).to(g.dtype)
    return f * g


def swiglu_bwd(DW, e, g):
    """kernels/swiglu.py:67-109.  Returns (h, df, de) -- the reference writes them into
    the buffers of (DW, e, g)."""
    ef = e.float()
    se = torch.sigmoid(ef)
    f = (se * ef).to(DW.dtype)
    h = f * g
    df = DW * f
    dg = DW * g
    de = (dg.float() * se * (1.0 + ef * (1.0 - se))).to(DW.dtype)
    return h, df, de


_S = 0.7978845608028654  # sqrt(2/pi), kernels/geglu.py:155


def geglu_approx_fwd(e, g):
    """kernels/geglu.py:142-167."""
    ef = e.float()
    f = 0.5 * ef * (torch.tanh(_S * ef * (1.0 + 0.044715 * ef * ef)) + 1.0)
    return f.to(g.dtype) * g


def geglu_approx_bwd(DW, e, g):
    """kernels/geglu.py:188-244."""
    ef = e.float()
    a = _S * ef
    b = a * 0.044715 * ef * ef
    T = 1.0 + torch.tanh(a + b)
    T2 = 0.5 * T
    Q2 = -T2 * (T - 2.0) * (a + 3.0 * b)
    df_de = T2 + Q2
    f = (T2 * ef).to(DW.dtype)
    h = f * g
    df = DW * f
    dg = DW * g
    de = (dg.float() * df_de).to(DW.dtype)
    return h, df, de


def geglu_exact_fwd(e, g):
    """kernels/geglu.py:31-53."""
    ef = e.float()
    f = 0.5 * ef * (torch.erf(ef * (1.0 / math.sqrt(2.0))) + 1.0)
    return f.to(g.dtype) * g


def geglu_exact_bwd(DW, e, g):
    """kernels/geglu.py:74-123."""
    ef = e.float()
    fp = 0.5 * (torch.erf(ef * (1.0 / math.sqrt(2.0))) + 1.0)
    f = (fp * ef).to(DW.dtype)
    h = f * g
    df = DW * f
    dg = DW * g
    df_de = fp + 0.3989422804014327 * ef * torch.exp(-0.5 * ef * ef)
    de = (dg.float() * df_de).to(DW.dtype)
    return h, df, de


# --------------------------------------------------------------------------------------
# Cross entropy on materialised logits  (kernels/cross_entropy_loss.py)
# --------------------------------------------------------------------------------------


def _ce_transform(x, softcap, scale):
    if scale != 0:
        x = scale * x
    if softcap != 0:
        x = softcap * torch.tanh(x / softcap)
    return x


def cross_entropy_fwd(logits, labels, softcap=0.0, scale=0.0):
    """kernels/cross_entropy_loss.py:35-111 / :114-199 + host combine :368-370.

    logits [T, V] any float dtype, labels int64 [T] (already shifted).  Returns
    (loss fp32 [T], logsumexp fp32 [T]); loss is 0 where label == -100.
    """
    x = _ce_transform(logits.float(), softcap, scale)
    lse = torch.logsumexp(x, dim=-1)
    lab = labels.clone()
    valid = lab != -100
    lab[~valid] = 0
    xl = x.gather(1, lab[:, None])[:, 0]
    loss = torch.where(valid, lse - xl, torch.zeros_like(lse))
    return loss, lse


def cross_entropy_bwd(logits, lse, labels, dloss, softcap=0.0, scale=0.0):
    """kernels/cross_entropy_loss.py:202-285: d logits (stored in logits.dtype)."""
    x = logits.float()
    if scale != 0:
        x = x * scale
    partial = x
    if softcap != 0:
        partial = torch.tanh(x / softcap)
        x = softcap * partial
    y = torch.exp(x - lse[:, None])
    valid = labels != -100
    lab = labels.clone()
    lab[~valid] = 0
    onehot = torch.zeros_like(y)
    onehot.scatter_(1, lab[:, None], 1.0)
    y = y - onehot
    if scale != 0:
        y = y * scale
    if softcap != 0:
        y = y * (1.0 - partial * partial)
    dl = torch.where(valid, dloss.float(), torch.zeros_like(dloss, dtype=torch.float32))
    return (dl[:, None] * y).to(logits.dtype)


def fast_cross_entropy_loss(logits, labels, softcap=0.0, scale=0.0, n_items=None):
    """kernels/cross_entropy_loss.py:421-449: sum(loss)/n_items, plus d logits for
    d(loss)=1.  logits [B,S,V], labels [B,S]."""
    B, S, V = logits.shape
    loss, lse = cross_entropy_fwd(logits.reshape(-1, V), labels.reshape(-1), softcap, scale)
    if n_items is None:
        n_items = torch.count_nonzero(labels != -100)
    total = loss.sum() / n_items
    dloss = torch.full_like(loss, 1.0) / n_items
    dlogits = cross_entropy_bwd(logits.reshape(-1, V), lse, labels.reshape(-1), dloss,
                                softcap, scale)
    return total, dlogits.reshape(B, S, V)


def fused_linear_cross_entropy(hidden, lm_head_weight, labels, n_items=None,
                               softcap=0.0, scale=0.0, lm_head_bias=None):
    """Restatement of unsloth_zoo.loss_utils.unsloth_fused_ce_loss (EXTERNAL, absent;
    parity unpinned) from its call site models/llama.py:1497-1509 and the logits path it
    replaces (:1525-1562): logits = lm_head(hidden) in model dtype, labels shifted by one
    *inside* (llama.py:1479-1482), mean over n_items.

    hidden [B,S,H], labels [B,S] unshifted.  Returns (loss, dHidden [B,S,H] for dloss=1).
    """
    B, S, H = hidden.shape
    logits = (hidden.reshape(-1, H).float() @ lm_head_weight.float().t())
    if lm_head_bias is not None:
        logits = logits + lm_head_bias.float()
    logits = logits.to(hidden.dtype).reshape(B, S, -1)
    shift = torch.full_like(labels, -100)
    shift[..., :-1] = labels[..., 1:]
    loss, dlogits = fast_cross_entropy_loss(logits, shift, softcap, scale, n_items)
    dH = (dlogits.reshape(B * S, -1).float() @ lm_head_weight.float()).to(hidden.dtype)
    return loss, dH.reshape(B, S, H)


# --------------------------------------------------------------------------------------
# NF4 double quantisation (bitsandbytes algorithm; parity unpinned)
# --------------------------------------------------------------------------------------

# 16-entry NormalFloat4 code book hard-coded in bitsandbytes' kernels (the reference
# passes code=NULL, kernels/utils.py:668).  External knowledge -- see SURVEY 8c.
NF4_CODE = torch.tensor([
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453,
    -0.28444138169288635, -0.18477343022823334, -0.09105003625154495, 0.0,
    0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
    0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=torch.float32)


def create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """bitsandbytes.functional.create_dynamic_map (published algorithm): the 256-entry
    8-bit dynamic code used for the second-level (absmax) quantisation."""
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 2 ** total_bits
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


def quantize_nf4(W, blocksize=64, blocksize2=256):
    """bitsandbytes quantize_4bit(quant_type='nf4', compress_statistics=True) restated:
    per-64 absmax, nearest NF4 code, two codes per byte (first element in the HIGH
    nibble); absmax is then mean-centred and quantised to 8 bit with the dynamic map in
    blocks of 256.  Returns a quant_state-like namespace with the field names the
    reference reads (kernels/utils.py:582-598).
    """
    shape = tuple(W.shape)
    flat = W.reshape(-1).float()
    n = flat.numel()
    assert n % blocksize == 0
    blocks = flat.reshape(-1, blocksize)
    absmax = blocks.abs().amax(dim=1)
    scaled = blocks / absmax.clamp_min(1e-30)[:, None]
    idx = (scaled[..., None] - NF4_CODE).abs().argmin(dim=-1).to(torch.uint8).reshape(-1)
    packed = ((idx[0::2] << 4) | idx[1::2]).to(torch.uint8).reshape(-1, 1)
    # second level
    offset = absmax.mean()
    am = absmax - offset
    code2 = create_dynamic_map()
    nb = am.numel()
    pad = (-nb) % blocksize2
    amp = torch.cat([am, am.new_zeros(pad)]).reshape(-1, blocksize2)
    absmax2 = amp.abs().amax(dim=1)
    sc = amp / absmax2.clamp_min(1e-30)[:, None]
    # nearest code (searchsorted on midpoints keeps memory small)
    mid = (code2[:-1] + code2[1:]) / 2
    q = torch.searchsorted(mid, sc.reshape(-1).contiguous()).to(torch.uint8)[:nb]
    state2 = SimpleNamespace(absmax=absmax2.contiguous(), code=code2, blocksize=blocksize2)
    qs = SimpleNamespace(absmax=q.contiguous(), shape=torch.Size(shape), dtype=W.dtype,
                         blocksize=blocksize, offset=offset.clone(), state2=state2,
                         quant_type="nf4")
    return packed, qs


def dequantize_absmax(qs):
    """Stage 1, kernels/utils.py:650-661: cdequantize_blockwise_fp32 then `+= offset`."""
    a8 = qs.absmax.long()
    blk = torch.arange(a8.numel()) // qs.state2.blocksize
    return qs.state2.code[a8] * qs.state2.absmax[blk] + qs.offset


def dequantize_nf4(packed, qs):
    """Stage 2, kernels/utils.py:663-679: out[2j] = NF4[W[j]>>4]*absmax[2j//bs],
    out[2j+1] = NF4[W[j]&15]*..., cast to quant_state.dtype, shape quant_state.shape."""
    absmax = dequantize_absmax(qs)
    b = packed.reshape(-1).long()
    idx = torch.stack([b >> 4, b & 0xF], dim=1).reshape(-1)
    vals = NF4_CODE[idx] * absmax[torch.arange(idx.numel()) // qs.blocksize]
    return vals.to(qs.dtype).reshape(qs.shape)


def fast_dequantize(W, quant_state=None):
    """kernels/utils.py:567-679 semantics: quant_state None -> W unchanged; returns
    out.t() iff W.shape[0] == 1 (the caller passed the packed weight transposed)."""
    if quant_state is None:
        return W
    out = dequantize_nf4(W, quant_state)
    return out.t() if W.shape[0] == 1 else out


def gemv_nf4(x, packed, qs, code_dtype=None):
    """`fast_gemv`, kernels/utils.py:874-973 (bsz == q_len == 1): fp32 absmax (:938-948), then
    bitsandbytes' 4-bit GEMV out[m] = sum_k NF4[nibble(m,k)] * absmax[(m*K+k)//bs] * x[k], rounded
    once to quant_state.dtype.  bitsandbytes is absent (parity unpinned): this is the exact-product
    fp32 form; the original casts the 16-entry code table (and absmax, and every product) to the
    16-bit dtype before accumulating in fp32.  `code_dtype` = that dtype restates the first of those
    roundings only -- the arithmetic of csrc/gemv.cu's pair-table kernel."""
    absmax = dequantize_absmax(qs)
    b = packed.reshape(-1).long()
    idx = torch.stack([b >> 4, b & 0xF], dim=1).reshape(-1)
    code = NF4_CODE if code_dtype is None else NF4_CODE.to(code_dtype).float()
    Wf = (code[idx] * absmax[torch.arange(idx.numel()) // qs.blocksize]).reshape(qs.shape)
    return (Wf.double() @ x.reshape(-1).double()).float().to(qs.dtype)


def fast_linear_forward(x, packed, qs, A, B, s):
    """kernels/utils.py:1082-1125, bsz == q_len == 1: out = gemv(x) ; temp = A x ; out += s B temp
    (adapters cast to the activation dtype, :1103-1105; fp32 accumulation, one final rounding)."""
    base = gemv_nf4(x, packed, qs).float() if qs is not None else (packed.float() @ x.reshape(-1).float())
    if A is None:
        return base.to(x.dtype)
    t = A.to(x.dtype).float() @ x.reshape(-1).float()
    return (base + s * (B.to(x.dtype).float() @ t)).to(x.dtype)


# --------------------------------------------------------------------------------------
# LoRA projections  (kernels/utils.py matmul_lora, kernels/fast_lora.py)
# --------------------------------------------------------------------------------------


def _mm(a, b):
    """bf16/fp16 GEMM with fp32 accumulation and one rounding at the end (cuBLAS)."""
    return (a.float() @ b.float()).to(a.dtype)


def _addmm_(out, a, b, alpha=1.0, beta=1.0):
    """out = beta*out + alpha*(a@b), fp32 accumulate, rounded to out.dtype (cuBLAS
    addmm_ semantics used throughout kernels/fast_lora.py)."""
    res = alpha * (a.float() @ b.float())
    if beta != 0:
        res = res + beta * out.float()
    return res.to(out.dtype)


def matmul_lora(X, W, W_quant, A, B, s):
    """kernels/utils.py:1128-1170: out = X @ dequant(W).T ; out += (X@A.T) @ (s*B.T)."""
    dtype = X.dtype
    shape = X.shape
    X2 = X.reshape(-1, shape[-1])
    Wd = fast_dequantize(W, W_quant)
    out = _mm(X2, Wd.t().to(dtype))
    if A is not None:
        XA = _mm(X2, A.t().to(dtype))
        out = _addmm_(out, XA, B.t().to(dtype), alpha=s)
    return out.reshape(*shape[:-1], -1)


def _lora_grads(X, dY, A, B, s):
    """kernels/fast_lora.py:172-189 / :476-493 / :639-640 with their bf16 roundings:
    d_A.T = s * X.T @ (dY @ B) ; d_B.T = s * (A @ X.T) @ dY   (A:[r,in], B:[out,r])."""
    dtype = X.dtype
    At, Bt = A.to(dtype).t(), B.to(dtype).t()          # [in,r], [r,out]
    dA_t = _addmm_(torch.empty(At.shape, dtype=dtype), X.t(), _mm(dY, Bt.t()), alpha=s, beta=0)
    dB_t = _addmm_(torch.empty(Bt.shape, dtype=dtype), _mm(At.t(), X.t()), dY, alpha=s, beta=0)
    return dA_t.t(), dB_t.t()


def lora_mlp_fwd(X, gate, up, down, act="swiglu"):
    """kernels/fast_lora.py:69-112.  gate/up/down = (W, W_quant, A, B, s)."""
    fwd = {"swiglu": swiglu_fwd, "geglu_approx": geglu_approx_fwd,
           "geglu_exact": geglu_exact_fwd}[act]
    e = matmul_lora(X, *gate)
    g = matmul_lora(X, *up)
    h = fwd(e, g)
    i = matmul_lora(h, *down)
    return i, e, g


def lora_mlp_bwd(dY, X, e, g, gate, up, down, act="swiglu"):
    """kernels/fast_lora.py:116-229.  Returns dX and (dA,dB) for gate, up, down."""
    bwd = {"swiglu": swiglu_bwd, "geglu_approx": geglu_approx_bwd,
           "geglu_exact": geglu_exact_bwd}[act]
    gW, gQ, gA, gB, gS = gate
    uW, uQ, uA, uB, uS = up
    dW_, dQ_, dA_, dB_, dS_ = down
    dtype = X.dtype
    shape = X.shape
    dY2 = dY.reshape(-1, dY.shape[-1])
    X2 = X.reshape(-1, shape[-1])
    e2 = e.reshape(-1, e.shape[-1])
    g2 = g.reshape(-1, g.shape[-1])
    # DW = dY @ dequant(downW) + s * (dY @ B) @ A   (:155)
    Wd = fast_dequantize(dW_, dQ_).to(dtype)
    DW = _mm(dY2, Wd)
    DW = _addmm_(DW, _mm(dY2, dB_.to(dtype)), dA_.to(dtype), alpha=dS_)
    h, df, de = bwd(DW, e2, g2)
    d_downA, d_downB = _lora_grads(h, dY2, dA_, dB_, dS_)
    d_upA, d_upB = _lora_grads(X2, df, uA, uB, uS)
    d_gateA, d_gateB = _lora_grads(X2, de, gA, gB, gS)
    # dX accumulates in bf16 across the four terms (:193-204)
    dX = _mm(df, fast_dequantize(uW, uQ).to(dtype))
    dX = _addmm_(dX, _mm(df, uB.to(dtype)), uA.to(dtype), alpha=uS)
    dX = _addmm_(dX, de, fast_dequantize(gW, gQ).to(dtype))
    dX = _addmm_(dX, _mm(de, gB.to(dtype)), gA.to(dtype), alpha=gS)
    return dX.reshape(shape), (d_gateA, d_gateB), (d_upA, d_upB), (d_downA, d_downB)


def lora_qkv_fwd(X, q, k, v):
    """kernels/fast_lora.py:368-430."""
    return matmul_lora(X, *q), matmul_lora(X, *k), matmul_lora(X, *v)


def lora_qkv_bwd(dQ, dK, dV, X, q, k, v):
    """kernels/fast_lora.py:432-540."""
    dtype = X.dtype
    shape = X.shape
    X2 = X.reshape(-1, shape[-1])
    outs = []
    dX = None
    for dO, (W, Wq, A, B, s) in ((dQ, q), (dK, k), (dV, v)):
        d2 = dO.reshape(-1, dO.shape[-1])
        outs.append(_lora_grads(X2, d2, A, B, s))
        Wd = fast_dequantize(W, Wq).to(dtype)
        dX = _mm(d2, Wd) if dX is None else _addmm_(dX, d2, Wd)
        dX = _addmm_(dX, _mm(d2, B.to(dtype)), A.to(dtype), alpha=s)
    return dX.reshape(shape), outs[0], outs[1], outs[2]


def lora_w_fwd(X, o):
    """kernels/fast_lora.py:604-615."""
    return matmul_lora(X, *o)


def lora_w_bwd(dY, X, o):
    """kernels/fast_lora.py:617-650."""
    W, Wq, A, B, s = o
    dtype = X.dtype
    shape = X.shape
    d2 = dY.reshape(-1, dY.shape[-1])
    X2 = X.reshape(-1, shape[-1])
    dA, dB = _lora_grads(X2, d2, A, B, s)
    dX = _mm(d2, fast_dequantize(W, Wq).to(dtype))
    dX = _addmm_(dX, _mm(d2, B.to(dtype)), A.to(dtype), alpha=s)
    return dX.reshape(shape), (dA, dB)


# --------------------------------------------------------------------------------------
# fp32 "truth" helpers (no intermediate rounding) used by the err_ours <= err_ref gate
# --------------------------------------------------------------------------------------


def matmul_lora_truth(X, Wd, A, B, s):
    X2 = X.reshape(-1, X.shape[-1]).double()
    out = X2 @ Wd.double().t()
    if A is not None:
        out = out + s * (X2 @ A.double().t()) @ B.double().t()
    return out.reshape(*X.shape[:-1], -1)


# --------------------------------------------------------------------------------------
# attention between RoPE and apply_o  (utils/attention_dispatch.py:298-617; the reference calls
# flash-attn / xformers / SDPA -- all compute the masked softmax below)
# --------------------------------------------------------------------------------------


def attention(Q, K, V, scale, window_left=-1, softcap=0.0, lengths=None):
    """fp32 definition of what the reference's attention backends compute for training:
    causal, optional sliding window as flash-attn's `window_size=(w, w)` under causal masking
    (key j visible to query i iff i - w <= j <= i; mistral.py:112-128, gemma2.py:139-150), optional
    tanh soft-capping of the scaled scores (gemma2.py:152-199), block-diagonal over the documents
    of a packed row (`lengths`, attention_dispatch.py:433-447).
    Q [B,S,Hq,D], K / V [B,S,Hk,D] -> (O [B,S,Hq,D] fp32, lse [B,Hq,S] | [Hq, B*S] for packed rows)."""
    B, S, Hq, D = Q.shape
    rep = Hq // K.shape[2]
    q = Q.float().permute(0, 2, 1, 3)
    k = K.float().permute(0, 2, 1, 3).repeat_interleave(rep, 1)
    v = V.float().permute(0, 2, 1, 3).repeat_interleave(rep, 1)
    if lengths is not None:
        q, k, v = (t.permute(1, 0, 2, 3).reshape(1, Hq, B * S, D) for t in (q, k, v))
    n = q.shape[2]
    s = (q @ k.transpose(-1, -2)) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    i = torch.arange(n, device=Q.device)[:, None]
    j = torch.arange(n, device=Q.device)[None, :]
    mask = j <= i
    if window_left >= 0:
        mask = mask & (j >= i - window_left)
    if lengths is not None:
        doc = torch.repeat_interleave(torch.arange(len(lengths), device=Q.device),
                                      torch.as_tensor(lengths, device=Q.device))
        mask = mask & (doc[:, None] == doc[None, :])
    s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.softmax(s, -1) @ v
    if lengths is not None:
        o = o.reshape(Hq, B, S, D).permute(1, 0, 2, 3)
        return o.permute(0, 2, 1, 3), lse.reshape(Hq, B * S)
    return o.permute(0, 2, 1, 3), lse
