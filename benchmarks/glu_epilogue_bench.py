"""A/B of the gated-activation epilogue (ub200_gemm_glu) at cfg2 sizes (T = 8192, H = 4096, I = 14336, bf16):
  (1) the two kernels alone: DW GEMM + glu_bwd launch vs ONE launch; up GEMM + glu_fwd launch vs ONE launch;
  (2) LoRA_MLP forward / backward (NF4 base, r = 16) under UB200_FUSED_GLU = 0 / fwd / bwd / 1.
CUDA events, L2 flushed before every sample, variants interleaved round-robin, medians of 9."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200 import _lib as L  # noqa: E402
from unsloth_b200.kernels import utils as KU  # noqa: E402
from unsloth_b200.nf4 import quantize_nf4  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
_flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def timed(fn):
    _flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def main():
    torch.manual_seed(0)
    KU.set_keep_dequant(True)
    T_, H, I, r = 8192, 4096, 14336, 16
    # ---- (1) kernels alone
    dY = (torch.randn(T_, H, device=DEV) * 0.1).to(BF)
    Wd = (torch.randn(H, I, device=DEV) * 0.02).to(BF)          # [H, I]: MN-major B operand of dY @ W_down
    G = torch.randn(T_, 64, device=DEV).to(BF)
    Ac = (torch.randn(64, I, device=DEV) * 0.01).to(BF)
    X = torch.randn(T_, H, device=DEV).to(BF)
    Wu = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    XA = torch.randn(T_, 64, device=DEV).to(BF)
    Bp = (torch.randn(I, 64, device=DEV) * 0.01).to(BF)
    e = torch.randn(T_, I, device=DEV).to(BF)
    g = torch.randn(T_, I, device=DEV).to(BF)
    out = torch.empty(T_, I, device=DEV, dtype=BF)
    out2 = torch.empty(T_, I, device=DEV, dtype=BF)
    bsegs = [(dY, Wd, H), (G, Ac, 64)]
    fsegs = [(X, Wu, H), (XA, Bp, 64)]

    def bwd_two():
        KU.gemm(T_, I, bsegs, out, b_mn=True)
        K.swiglu_DWf_DW_dfg_kernel(out, e, g)

    def fwd_two():
        KU.gemm(T_, I, fsegs, out)
        K.swiglu_fg_kernel(e, out)

    variants = {
        "bwd: DW gemm alone": lambda: KU.gemm(T_, I, bsegs, out, b_mn=True),
        "bwd: DW gemm + glu_bwd (2 launches)": bwd_two,
        "bwd: gemm_glu (1 launch)": lambda: KU.gemm_glu(L.GLU_EPI_BWD, 0, T_, I, bsegs, out, e, g, b_mn=True),
        "fwd: up gemm alone": lambda: KU.gemm(T_, I, fsegs, out),
        "fwd: up gemm + glu_fwd (2 launches)": fwd_two,
        "fwd: gemm_glu (1 launch)": lambda: KU.gemm_glu(L.GLU_EPI_FWD, 0, T_, I, fsegs, out2, e, out),
    }
    ts = {k: [] for k in variants}
    for rnd in range(11):
        for k, fn in variants.items():
            t = timed(fn)
            if rnd >= 2:
                ts[k].append(t)
            e.normal_(); g.normal_()          # in-place outputs drift otherwise
    for k in variants:
        print(json.dumps({"variant": k, "ms": round(med(ts[k]), 4)}), flush=True)
    del dY, Wd, G, Ac, X, Wu, XA, Bp, e, g, out, out2

    # ---- (2) LoRA_MLP under the schedules
    Xb = torch.randn(4, 2048, H, device=DEV).to(BF)
    dYb = (torch.randn(4, 2048, H, device=DEV) * 0.1).to(BF)

    def mk(o, i):
        W = (torch.randn(o, i, device=DEV) * 0.02).to(BF)
        p, q = quantize_nf4(W)
        A = torch.nn.Parameter((torch.rand(r, i, device=DEV) * 2 - 1) / i ** 0.5)
        B = torch.nn.Parameter(torch.randn(o, r, device=DEV) * 0.02)
        return p, q, A, B
    gate, up, down = mk(I, H), mk(I, H), mk(H, I)
    acc = {m: ([], []) for m in ("0", "fwd", "bwd", "1")}
    for rnd in range(11):
        for m in acc:
            os.environ["UB200_FUSED_GLU"] = m
            x = Xb.clone().requires_grad_()
            _flush.fill_(1)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            o = K.LoRA_MLP.apply(x, gate[0], gate[1], gate[2], gate[3], 1.0, up[0], up[1], up[2], up[3], 1.0,
                                 down[0], down[1], down[2], down[3], 1.0, K.swiglu_fg_kernel,
                                 K.swiglu_DWf_DW_dfg_kernel, True)
            ev[1].record()
            o.backward(dYb)
            ev[2].record()
            torch.cuda.synchronize()
            if rnd >= 2:
                acc[m][0].append(ev[0].elapsed_time(ev[1])); acc[m][1].append(ev[1].elapsed_time(ev[2]))
    for m in acc:
        print(json.dumps({"UB200_FUSED_GLU": m, "mlp_fwd_ms": round(med(acc[m][0]), 4),
                          "mlp_bwd_ms": round(med(acc[m][1]), 4)}), flush=True)


if __name__ == "__main__":
    main()
