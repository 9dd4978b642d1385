"""TEST INFRASTRUCTURE ONLY: a CPU emulator of the C ABI of include/unsloth_b200.h.

`install(monkeypatch)` replaces `unsloth_b200._lib.call` by a dispatcher that interprets every entry
point on HOST memory (raw pointers + strides, exactly the arguments the shipped Python shims pass)
with the arithmetic of the CPU oracle (oracle/restate.py).  With it the whole host stack above the
C ABI -- kernels/*.py, patch.py, ddp.py -- runs on a CPU-only box, so the `-m "not gpu"` suite can
check the HOST LOGIC (shapes, strides, in-place contracts, label plumbing, rank-block orchestration,
packed batches, the fused add+norm chain) against the reference's golden vectors and against stock
HuggingFace.  It is an executable restatement of the header's semantics; it is never imported by
the product and the GPU suite never uses it.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace as NS

import torch

from oracle import restate as R

_DT = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}


def _addr(p):
    if p is None:
        return 0
    return p.value if isinstance(p, ctypes.c_void_p) else int(p)


def mem(p, dtype, count):
    """1-D tensor aliasing `count` elements of host memory at pointer p."""
    a = _addr(p)
    if a == 0 or count <= 0:
        return None
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    buf = (ctypes.c_char * nbytes).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=count)


def view2d(p, dtype, rows, cols, ld):
    flat = mem(p, dtype, (rows - 1) * ld + cols)
    return flat.as_strided((rows, cols), (ld, 1))


def view4d(p, dtype, sizes, strides):
    span = sum((n - 1) * s for n, s in zip(sizes, strides)) + 1
    return mem(p, dtype, span).as_strided(tuple(sizes), tuple(strides))


# ---------------------------------------------------------------------------------------------
def rms_layernorm_fwd(X, xs, W, wdt, Y, ys, r, n_rows, n_cols, eps, gemma, dt, stream):
    Xv, Wv = view2d(X, _DT[dt], n_rows, n_cols, xs), mem(W, _DT[wdt], n_cols)
    Yv, rv = R.rms_layernorm_fwd(Xv, Wv, eps, bool(gemma))
    view2d(Y, _DT[dt], n_rows, n_cols, ys).copy_(Yv)
    mem(r, torch.float32, n_rows).copy_(rv)


def rms_layernorm_bwd(dY, dys, X, xs, W, wdt, r, dX, dxs, n_rows, n_cols, gemma, dt, stream):
    dYv, Xv = view2d(dY, _DT[dt], n_rows, n_cols, dys), view2d(X, _DT[dt], n_rows, n_cols, xs)
    res = R.rms_layernorm_bwd(dYv, Xv, mem(W, _DT[wdt], n_cols), mem(r, torch.float32, n_rows), bool(gemma))
    view2d(dX, _DT[dt], n_rows, n_cols, dxs).copy_(res)          # dX may alias dY


def add_rms_layernorm_fwd(A, as_, B, bs, W, S, ss, Y, ys, r, n_rows, n_cols, eps, dt, stream):
    d = _DT[dt]
    Sv = view2d(A, d, n_rows, n_cols, as_) + view2d(B, d, n_rows, n_cols, bs)      # one rounding
    Yv, rv = R.rms_layernorm_fwd(Sv, mem(W, d, n_cols), eps, False)
    view2d(S, d, n_rows, n_cols, ss).copy_(Sv)
    view2d(Y, d, n_rows, n_cols, ys).copy_(Yv)
    mem(r, torch.float32, n_rows).copy_(rv)


def rms_layernorm_bwd_acc(dY, dys, X, xs, W, r, dS, dss, n_rows, n_cols, dt, stream):
    d = _DT[dt]
    dx = R.rms_layernorm_bwd(view2d(dY, d, n_rows, n_cols, dys).float(), view2d(X, d, n_rows, n_cols, xs),
                             mem(W, d, n_cols), mem(r, torch.float32, n_rows), False)       # fp32
    dSv = view2d(dS, d, n_rows, n_cols, dss)
    dSv.copy_((dSv.float() + dx).to(d))                          # single rounding of the sum


def rope_qk(Q, qb, qh, qs, K, kb, kh, ks, cos, cos_ld, sin, sin_ld, indices, batch, seqlen, hq, hk,
            D, backward, dt, tdt, cdt, stream):
    Qv = view4d(Q, _DT[dt], (batch, hq, seqlen, D), (qb, qh, qs, 1))
    Kv = view4d(K, _DT[dt], (batch, hk, seqlen, D), (kb, kh, ks, 1)) if _addr(K) and hk else None
    idx = mem(indices, torch.int32, batch * seqlen)
    rows = seqlen if idx is None else int(idx.max().item()) + 1
    cosv = view2d(cos, _DT[tdt], rows, D, cos_ld)
    sinv = view2d(sin, _DT[tdt], rows, D, sin_ld)
    if idx is None and cdt == tdt:
        # no-index form: products / sums rounded in the table dtype (rope_embedding.py:129-158)
        Qo = R.rope_noindex(Qv.transpose(1, 2), cosv, sinv, backward=bool(backward)).transpose(1, 2)
        Ko = None if Kv is None else R.rope_noindex(Kv.transpose(1, 2), cosv, sinv,
                                                    backward=bool(backward)).transpose(1, 2)
    else:
        Kin = Kv if Kv is not None else Qv[:, :1]
        Qo, Ko = R.rope_qk(Qv, Kin, cosv, sinv, idx, backward=bool(backward))
        Ko = None if Kv is None else Ko
    Qv.copy_(Qo.to(Qv.dtype))
    if Kv is not None:
        Kv.copy_(Ko.to(Kv.dtype))


_GLU = {0: (R.swiglu_fwd, R.swiglu_bwd), 1: (R.geglu_approx_fwd, R.geglu_approx_bwd),
        2: (R.geglu_exact_fwd, R.geglu_exact_bwd)}


def glu_fwd(act, e, g, h, n, dt, stream):
    d = _DT[dt]
    mem(h, d, n).copy_(_GLU[act][0](mem(e, d, n).view(1, -1), mem(g, d, n).view(1, -1)).view(-1))


def glu_bwd(act, DW, e, g, n, dt, stream):
    d = _DT[dt]
    DWv, ev, gv = mem(DW, d, n), mem(e, d, n), mem(g, d, n)
    hh, df, de = _GLU[act][1](DWv.view(1, -1), ev.view(1, -1), gv.view(1, -1))
    DWv.copy_(hh.view(-1)); ev.copy_(df.view(-1)); gv.copy_(de.view(-1))


def cross_entropy_fwd(logits, ld, labels, loss, lse, n_rows, vocab, softcap, scale, dt, stream):
    lg = view2d(logits, _DT[dt], n_rows, vocab, ld)
    l, s = R.cross_entropy_fwd(lg, mem(labels, torch.int64, n_rows), softcap, scale)
    mem(loss, torch.float32, n_rows).copy_(l)
    mem(lse, torch.float32, n_rows).copy_(s)


def cross_entropy_bwd(logits, ld, lse, labels, dloss, dloss_stride, n_rows, vocab, softcap, scale, dt,
                      stream):
    lg = view2d(logits, _DT[dt], n_rows, vocab, ld)
    if dloss_stride:
        dl = mem(dloss, torch.float32, (n_rows - 1) * dloss_stride + 1)[::dloss_stride]
    else:
        dl = mem(dloss, torch.float32, 1).expand(n_rows)
    lg.copy_(R.cross_entropy_bwd(lg, mem(lse, torch.float32, n_rows), mem(labels, torch.int64, n_rows),
                                 dl, softcap, scale))


def dequantize_nf4(packed, absmax_q, code2, absmax2, offset, out, n, bs, bs2, odt, stream):
    nb = n // bs
    qs = NS(absmax=mem(absmax_q, torch.uint8, nb), shape=torch.Size((1, n)), dtype=_DT[odt], blocksize=bs,
            offset=(mem(offset, torch.float32, 1)[0] if _addr(offset) else torch.tensor(0.0)),
            state2=NS(absmax=mem(absmax2, torch.float32, (nb + bs2 - 1) // bs2),
                      code=mem(code2, torch.float32, 256), blocksize=bs2))
    mem(out, _DT[odt], n).copy_(R.dequantize_nf4(mem(packed, torch.uint8, n // 2), qs).view(-1))


def quantize_nf4(W, dt, packed, absmax, n, bs, stream):
    flat = mem(W, _DT[dt], n).float().reshape(-1, bs)
    am = flat.abs().amax(dim=1)
    idx = ((flat / am.clamp_min(1e-30)[:, None])[..., None] - R.NF4_CODE).abs().argmin(-1).to(torch.uint8).reshape(-1)
    mem(packed, torch.uint8, n // 2).copy_((idx[0::2] << 4) | idx[1::2])
    mem(absmax, torch.float32, n // bs).copy_(am)


def gemm(M, N, segs, n_segs, a_mn, b_mn, abdt, C, ldc, cdt, alpha, accumulate, split_k, ws, block_n,
         cta_group, stream):
    d = _DT[abdt]
    acc = torch.zeros(M, N, dtype=torch.float64)
    for i in range(n_segs):
        sg = segs[i]
        K = int(sg.k)
        A = view2d(sg.a, d, K, M, sg.lda).t() if a_mn else view2d(sg.a, d, M, K, sg.lda)
        B = view2d(sg.b, d, K, N, sg.ldb).t() if b_mn else view2d(sg.b, d, N, K, sg.ldb)
        acc += A.double() @ B.double().t()
    acc *= alpha
    Cv = view2d(C, _DT[cdt], M, N, ldc)
    if accumulate:
        acc += Cv.double()
    Cv.copy_(acc.to(_DT[cdt]))


def gemm_glu(mode, act, M, N, segs, n_segs, a_mn, b_mn, dt, C, ldc, e, g, ld_eg, alpha, block_n, cta_group,
             stream):
    """ub200_gemm_glu: the GEMM, rounded to `dt`, then the gated activation on the tile (header contract)."""
    bn = block_n or (256 if N > 128 else (128 if N > 64 else 64))
    assert mode in (1, 2) and N % bn == 0 and alpha == 1.0 and ldc % 16 == 0 and ld_eg % 16 == 0 and dt in (1, 2)
    d = _DT[dt]
    tile = torch.empty(M, N, dtype=d)
    gemm(M, N, segs, n_segs, a_mn, b_mn, dt, tile.data_ptr(), N, dt, alpha, 0, 1, None, block_n, cta_group,
         stream)
    Cv, ev, gv = view2d(C, d, M, N, ldc), view2d(e, d, M, N, ld_eg), view2d(g, d, M, N, ld_eg)
    if mode == 1:      # forward: tile = up projection
        gv.copy_(tile)
        Cv.copy_(_GLU[act][0](ev.contiguous().view(1, M, N), tile.view(1, M, N)).view(M, N))
    else:              # backward: tile = DW
        hh, df, de = _GLU[act][1](tile.view(1, M, N), ev.contiguous().view(1, M, N), gv.contiguous().view(1, M, N))
        Cv.copy_(hh.view(M, N)); ev.copy_(df.view(M, N)); gv.copy_(de.view(M, N))


def gemm_grouped(probs, n_probs, abdt, scratch, stream):
    """Sequential interpretation of a grouped launch: problems in list order, so every in-launch
    dependency (wait_problem < own index) is already satisfied.  Checks the documented contract."""
    for i in range(n_probs):
        g = probs[i]
        assert g.n_segs >= 1 and g.block_n in (64, 128, 256)
        assert not (g.b_mn_major and g.block_n < 128)
        if g.wait_problem >= 0:
            assert g.wait_problem < i and probs[g.wait_problem].signals
            assert 0 <= g.wait_segment < g.n_segs
        assert g.split_k <= 1 or g.workspace
        gemm(g.M, g.N, g.segs, g.n_segs, g.a_mn_major, g.b_mn_major, abdt, g.C, g.ldc, g.c_dtype, g.alpha,
             g.accumulate, 1, None, g.block_n, 2, stream)


def cast_pad_2d(src, sdt, sld, rows, cols, dst, ddt, dld, drows, dcols, roff, coff, scale, transpose,
                stream):
    s = view2d(src, _DT[sdt], rows, cols, sld).double() * scale
    s = s.t() if transpose else s
    Dv = view2d(dst, _DT[ddt], drows, dcols, dld)
    Dv.zero_()
    Dv[roff:roff + s.shape[0], coff:coff + s.shape[1]] = s.to(_DT[ddt])


def cast_pad_multi(descs, n, stream):
    for i in range(n):
        d = descs[i]
        cast_pad_2d(d.src, d.src_dtype, d.src_ld, d.rows, d.cols, d.dst, d.dst_dtype, d.dst_ld, d.dst_rows,
                    d.dst_cols, d.row_off, d.col_off, d.scale, d.transpose, stream)


def accumulate_multi(descs, n, stream):
    for i in range(n):
        d = descs[i]
        span = (d.rows - 1) * d.src_rs + (d.cols - 1) * d.src_cs + 1
        src = mem(d.src, torch.float32, span).as_strided((d.rows, d.cols), (d.src_rs, d.src_cs))
        dst = view2d(d.dst, torch.float32, d.rows, d.cols, d.cols)
        dst += src


def gemv_nf4(x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16, out, m, k, bs, bs2, lora_B,
             ldb, lora_t, r, s, dt, stream):
    d = _DT[dt]
    n = m * k
    if _addr(absmax_f32):
        am = mem(absmax_f32, torch.float32, n // bs)
    else:
        am = mem(code2, torch.float32, 256)[mem(absmax_q, torch.uint8, n // bs).long()] * \
            mem(absmax2, torch.float32, (n // bs + bs2 - 1) // bs2)[torch.arange(n // bs) // bs2]
        if _addr(offset):
            am = am + mem(offset, torch.float32, 1)[0]
    code = mem(code16, torch.float32, 16) if _addr(code16) else R.NF4_CODE
    b = mem(packed, torch.uint8, n // 2).long()
    idx = torch.stack([b >> 4, b & 0xF], dim=1).reshape(-1)
    Wf = (code[idx] * am[torch.arange(n) // bs]).reshape(m, k)
    y = Wf.double() @ mem(x, d, k).double()
    if _addr(lora_B):
        y = y + s * (view2d(lora_B, d, m, r, ldb).double() @ mem(lora_t, torch.float32, r).double())
    mem(out, d, m).copy_(y.to(d))


def gemv_dense(x, W, ldw, out, m, k, dt, odt, stream):
    d = _DT[dt]
    mem(out, _DT[odt], m).copy_((view2d(W, d, m, k, ldw).double() @ mem(x, d, k).double()).to(_DT[odt]))


_TABLE = {
    "ub200_rms_layernorm_fwd": rms_layernorm_fwd, "ub200_rms_layernorm_bwd": rms_layernorm_bwd,
    "ub200_add_rms_layernorm_fwd": add_rms_layernorm_fwd, "ub200_rms_layernorm_bwd_acc": rms_layernorm_bwd_acc,
    "ub200_rope_qk": rope_qk, "ub200_glu_fwd": glu_fwd, "ub200_glu_bwd": glu_bwd,
    "ub200_cross_entropy_fwd": cross_entropy_fwd, "ub200_cross_entropy_bwd": cross_entropy_bwd,
    "ub200_dequantize_nf4": dequantize_nf4, "ub200_quantize_nf4": quantize_nf4, "ub200_gemm": gemm, "ub200_gemm_grouped": gemm_grouped,
    "ub200_gemm_glu": gemm_glu,
    "ub200_cast_pad_2d": cast_pad_2d, "ub200_gemv_nf4": gemv_nf4,
    "ub200_cast_pad_multi": cast_pad_multi, "ub200_accumulate_multi": accumulate_multi,
    "ub200_gemv_dense": gemv_dense,
}


def install(monkeypatch):
    """Route unsloth_b200._lib.call to the emulator and let CPU tensors through."""
    import unsloth_b200._lib as L
    calls = []

    def call(name, *args):
        calls.append(name)
        L.launch_count += 1
        args = [a.value if isinstance(a, ctypes.c_void_p) else a for a in args]
        _TABLE[name](*args)

    monkeypatch.setattr(L, "call", call)
    monkeypatch.setattr(L, "stream", lambda: None)
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    return calls
