"""How much of the gap to the HBM roofline of the short bandwidth-bound kernels is size (fixed
launch/ramp/tail cost at ~100 MB problems) and how much is the kernel: same kernels at 1x, 4x, 16x
the cfg2 row count, next to a plain device copy of the same byte count."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsloth_b200 import _lib as L  # noqa: E402
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200.kernels.rope_embedding import _launch as rope_launch  # noqa: E402
from unsloth_b200.nf4 import quantize_nf4  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
_flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        _flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def rep(name, ms, nbytes):
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "MB": round(nbytes / 1e6, 1),
                      "GBps": round(nbytes / ms / 1e6, 1), "frac_of_6583": round(nbytes / ms / 1e6 / 6583.5, 3)}), flush=True)


H = 4096
W = torch.ones(H, device=DEV, dtype=BF)
for mult in (1, 4, 16):
    T = 8192 * mult
    X = torch.randn(T, H, device=DEV, dtype=BF); Y = torch.empty_like(X); r = torch.empty(T, device=DEV)
    rep("copy_%dx" % mult, timeit(lambda: Y.copy_(X)), 2 * T * H * 2)
    rep("rms_fwd_%dx" % mult, timeit(lambda: L.call("ub200_rms_layernorm_fwd", L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(Y), H, L.ptr(r), T, H, 1e-5, 0, L.BF16, L.stream())), 2 * T * H * 2 + 4 * T)
    rep("rms_bwd_%dx" % mult, timeit(lambda: L.call("ub200_rms_layernorm_bwd", L.ptr(Y), H, L.ptr(X), H, L.ptr(W), L.BF16, L.ptr(r), L.ptr(Y), H, T, H, 0, L.BF16, L.stream())), 3 * T * H * 2)
    if mult <= 4:
        B, S, Hq, Hk, D = 4 * mult, 2048, 32, 8, 128
        q = torch.randn(B, S, Hq * D, device=DEV, dtype=BF); k = torch.randn(B, S, Hk * D, device=DEV, dtype=BF)
        inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D)); fr = torch.outer(torch.arange(S).float(), inv)
        emb = torch.cat([fr, fr], -1); cos, sin = emb.cos().to(DEV, BF), emb.sin().to(DEV, BF)
        Q, Kk = q.view(B, S, Hq, D).transpose(1, 2), k.view(B, S, Hk, D).transpose(1, 2)
        rep("rope_%dx" % mult, timeit(lambda: rope_launch(Q, Kk, cos, sin, None, False, True)), 2 * B * S * (Hq + Hk) * D * 2)
        del q, k
    del X, Y
for rows in (4096, 14336, 4 * 14336):
    Wd = (torch.randn(rows, 4096, device=DEV) * 0.02).to(BF)
    p, qs = quantize_nf4(Wd); out = torch.empty_like(Wd); n = Wd.numel()
    rep("nf4_dequant_%dx4096" % rows, timeit(lambda: K.fast_dequantize(p, qs, out=out)), n * (0.5 + 1 / 64) + n * 2)
    del Wd, out
