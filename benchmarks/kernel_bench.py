"""Per-kernel microbenchmarks at BASELINE.json cfg2 sizes (Llama-3-8B, T = 4 x 2048, bf16).

CUDA-event timing on the launching stream, >= 3 warm-ups, L2 flushed between timed
iterations (a 512 MB write), median of N.  Prints one JSON line per kernel with achieved
GB/s (HBM-bound kernels) or TFLOP/s (GEMMs) and the fraction of the measured peaks in
MEASURED_PEAKS.json (fallback 6650 GB/s / 1590 TFLOP/s).  `cublas` rows time torch.matmul on
the same operands for context (library baseline, not the product path).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200 import _lib as L  # noqa: E402
from unsloth_b200.kernels.cross_entropy_loss import _ce_backward_, _ce_forward  # noqa: E402
from unsloth_b200.kernels.rope_embedding import _launch as rope_launch  # noqa: E402
from unsloth_b200.nf4 import quantize_nf4  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    d = json.load(open(pp))
    peaks = {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "src": "measured"}

_flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, bytes_=None, flops=None, **extra):
    row = {"kernel": name, "ms": round(ms, 4)}
    if bytes_ is not None:
        gbs = bytes_ / ms / 1e6
        row.update(GBps=round(gbs, 1), frac_hbm=round(gbs / peaks["hbm_gbs"], 3))
    if flops is not None:
        tf = flops / ms / 1e9
        row.update(TFLOPs=round(tf, 1), frac_tensor=round(tf / peaks["bf16_tflops"], 3))
    row.update(extra)
    print(json.dumps(row), flush=True)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    T, H, I, V, Hq, Hk, D = 8192, 4096, 14336, 128256, 32, 8, 128
    torch.manual_seed(0)
    print(json.dumps({"peaks": peaks, "T": T}), flush=True)

    class N:
        weight = torch.ones(H, device=DEV, dtype=BF)
        variance_epsilon = 1e-5

    if "rms" in only or not only:
        X = torch.randn(T, H, device=DEV, dtype=BF)
        Y = torch.empty_like(X); r = torch.empty(T, device=DEV)
        W = N.weight
        f = lambda: L.call("ub200_rms_layernorm_fwd", L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(Y), H, L.ptr(r), T, H, 1e-5, 0, L.BF16, L.stream())
        report("rms_fwd", timeit(f), bytes_=2 * T * H * 2 + 4 * T)
        dY = torch.randn(T, H, device=DEV, dtype=BF)
        f = lambda: L.call("ub200_rms_layernorm_bwd", L.ptr(dY), H, L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(r), L.ptr(dY), H, T, H, 0, L.BF16, L.stream())
        report("rms_bwd", timeit(f), bytes_=3 * T * H * 2)
    if "rope" in only or not only:
        q = torch.randn(4, 2048, Hq * D, device=DEV, dtype=BF); k = torch.randn(4, 2048, Hk * D, device=DEV, dtype=BF)
        inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D)); fr = torch.outer(torch.arange(2048).float(), inv)
        emb = torch.cat([fr, fr], -1); cos, sin = emb.cos().to(DEV, BF), emb.sin().to(DEV, BF)
        Q, Kk = q.view(4, 2048, Hq, D).transpose(1, 2), k.view(4, 2048, Hk, D).transpose(1, 2)
        f = lambda: rope_launch(Q, Kk, cos, sin, None, False, True)
        report("rope_qk", timeit(f), bytes_=2 * T * (Hq + Hk) * D * 2 + T * D * 2)
    if "glu" in only or not only:
        e = torch.randn(T, I, device=DEV, dtype=BF); g = torch.randn(T, I, device=DEV, dtype=BF); h = torch.empty_like(e)
        f = lambda: L.call("ub200_glu_fwd", 0, L.ptr(e), L.ptr(g), L.ptr(h), e.numel(), L.BF16, L.stream())
        report("swiglu_fwd", timeit(f), bytes_=3 * T * I * 2)
        DW = torch.randn(T, I, device=DEV, dtype=BF)
        f = lambda: L.call("ub200_glu_bwd", 0, L.ptr(DW), L.ptr(e), L.ptr(g), e.numel(), L.BF16, L.stream())
        report("swiglu_bwd", timeit(f), bytes_=6 * T * I * 2)
        del e, g, h, DW
    if "addrms" in only or not only:
        A = torch.randn(T, H, device=DEV, dtype=BF); B = torch.randn(T, H, device=DEV, dtype=BF)
        S = torch.empty_like(A); Y = torch.empty_like(A); r = torch.empty(T, device=DEV)
        W = N.weight
        f = lambda: L.call("ub200_add_rms_layernorm_fwd", L.ptr(A), H, L.ptr(B), H, L.ptr(W), L.ptr(S), H,
                           L.ptr(Y), H, L.ptr(r), T, H, 1e-5, L.BF16, L.stream())
        report("add_rms_fwd", timeit(f), bytes_=4 * T * H * 2 + 4 * T)
        f = lambda: L.call("ub200_rms_layernorm_bwd_acc", L.ptr(A), H, L.ptr(S), H, L.ptr(W), L.ptr(r),
                           L.ptr(B), H, T, H, L.BF16, L.stream())
        report("rms_bwd_acc", timeit(f), bytes_=4 * T * H * 2)
        del A, B, S, Y
    if "gemv" in only or not only:
        # microsecond kernels: a Python launch costs more than the kernel, so time a CUDA graph of
        # `reps` launches cycling over enough distinct weight copies to exceed L2 (126 MB)
        for (m, k) in ((H, H), (I, H), (H, I)):
            Wd = (torch.randn(m, k, device=DEV) * 0.02).to(BF)
            packed, qs = quantize_nf4(Wd)
            n = m * k
            copies = max(2, int(300e6 // (n // 2)) + 1)
            packs = [packed.clone() for _ in range(copies)]
            x = torch.randn(1, 1, k, device=DEV, dtype=BF)
            out = torch.empty(1, 1, m, device=DEV, dtype=BF)
            reps = copies * 2
            K.fast_gemv(x, packs[0], qs, out=out); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(reps):
                    K.fast_gemv(x, packs[i % copies], qs, out=out)
            ms = timeit(g.replay, iters=10, flush=False) / reps
            report("gemv_nf4_%dx%d" % (m, k), ms, bytes_=n * (0.5 + 1 / 64) + 2 * (m + k), copies=copies,
                   timing="cuda graph of %d launches over %d weight copies" % (reps, copies))
            del Wd, packs
        lm = torch.randn(V, H, device=DEV, dtype=BF); x = torch.randn(H, device=DEV, dtype=BF)
        o = torch.empty(V, device=DEV, dtype=BF)
        f = lambda: L.call("ub200_gemv_dense", L.ptr(x), L.ptr(lm), H, L.ptr(o), V, H, L.BF16, L.BF16, L.stream())
        report("gemv_dense_lm_head", timeit(f, iters=20), bytes_=V * H * 2 + 2 * (V + H))
        f = lambda: torch.mv(lm, x, out=o)
        report("torch_mv_lm_head", timeit(f, iters=20), bytes_=V * H * 2 + 2 * (V + H))
        del lm
    if "nf4" in only or not only:
        Wd = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
        packed, qs = quantize_nf4(Wd)
        out = torch.empty_like(Wd)
        f = lambda: K.fast_dequantize(packed, qs, out=out)
        n = Wd.numel()
        report("nf4_dequant", timeit(f), bytes_=n * (0.5 + 1 / 64 + 4 / 16384) + n * 2)
        del Wd, out
    if "ce" in only or not only:
        Tc = 2048
        logits = torch.randn(Tc, V, device=DEV, dtype=BF)
        labels = torch.randint(0, V, (Tc,), device=DEV)
        res = {}
        def f():
            res["l"], res["lse"] = _ce_forward(logits, labels, 0.0, 0.0)
        report("ce_fwd", timeit(f), bytes_=Tc * V * 2)
        dl = torch.full((1,), 1e-3, device=DEV)
        f2 = lambda: _ce_backward_(logits, res["lse"], labels, dl, 0, 0.0, 0.0)
        report("ce_bwd", timeit(f2), bytes_=2 * Tc * V * 2)
        del logits
    if "gemm" in only or not only:
        shapes = [("qo_proj", T, 4096, 4096), ("kv_proj", T, 1024, 4096), ("gate_up", T, I, 4096),
                  ("down", T, 4096, I), ("lm_head_chunk", 2048, V, 4096)]
        for name, M, Nn, Kk in shapes:
            A = torch.randn(M, Kk, device=DEV, dtype=BF)
            B = (torch.randn(Nn, Kk, device=DEV) * 0.02).to(BF)
            C = torch.empty(M, Nn, device=DEV, dtype=BF)
            fl = 2.0 * M * Nn * Kk
            report("gemm_kk_" + name, timeit(lambda: K.gemm(M, Nn, [(A, B, Kk)], C), flush=False), flops=fl, M=M, N=Nn, K=Kk)
            report("gemm_kk_1cta_" + name, timeit(lambda: K.gemm(M, Nn, [(A, B, Kk)], C, cta_group=1), flush=False), flops=fl)
            report("cublas_" + name, timeit(lambda: torch.matmul(A, B.t(), out=C), flush=False), flops=fl)
            Bt = B.t().contiguous()
            report("gemm_kmn_" + name, timeit(lambda: K.gemm(M, Nn, [(A, Bt, Kk)], C, b_mn=True), flush=False), flops=fl)
            del A, B, C, Bt
        # LoRA-grad shape: dA^T[in,64] = X^T @ G  (MN/MN, split-K over tokens)
        X = torch.randn(T, 4096, device=DEV, dtype=BF); G = torch.randn(T, 64, device=DEV, dtype=BF)
        o = torch.empty(4096, 64, device=DEV, dtype=torch.float32)
        report("gemm_mnmn_dA", timeit(lambda: K.gemm(4096, 64, [(X, G, T)], o, a_mn=True, b_mn=True, split_k=4)),
               bytes_=T * 4096 * 2 + T * 64 * 2, flops=2.0 * T * 4096 * 64)
        for sk in (1, 2, 4, 8, 16):
            report("gemm_mnmn_dA_split%d" % sk, timeit(lambda: K.gemm(4096, 64, [(X, G, T)], o, a_mn=True, b_mn=True, split_k=sk)),
                   bytes_=T * 4096 * 2 + T * 64 * 2)
        A64 = torch.randn(64, 4096, device=DEV, dtype=BF); XA = torch.empty(T, 64, device=DEV, dtype=BF)
        for sk in (1, 2, 4):
            report("gemm_xa_skinny_split%d" % sk, timeit(lambda: K.gemm(T, 64, [(X, A64, 4096)], XA, split_k=sk)),
                   bytes_=T * 4096 * 2 + T * 64 * 2)
        big = torch.randn(T, I, device=DEV, dtype=BF); A64b = torch.randn(64, I, device=DEV, dtype=BF)
        for sk in (1, 2, 4):
            report("gemm_xa_skinny_K14336_split%d" % sk, timeit(lambda: K.gemm(T, 64, [(big, A64b, I)], XA, split_k=sk)),
                   bytes_=T * I * 2 + T * 64 * 2)
        o2 = torch.empty(I, 64, device=DEV, dtype=torch.float32)
        for sk in (1, 2, 4):
            report("gemm_mnmn_dB_M14336_split%d" % sk, timeit(lambda: K.gemm(I, 64, [(big, G, T)], o2, a_mn=True, b_mn=True, split_k=sk)),
                   bytes_=T * I * 2 + T * 64 * 2)
        report("gemm_xa_skinny", timeit(lambda: K.gemm(T, 64, [(X, A64, 4096)], XA)), bytes_=T * 4096 * 2 + T * 64 * 2,
               flops=2.0 * T * 4096 * 64)


if __name__ == "__main__":
    main()
