"""A/B of the gated-activation epilogue (ub200_gemm_glu) at cfg2 sizes (T = 8192, H = 4096, I = 14336, bf16):
the two kernels alone: DW GEMM + glu_bwd launch vs ONE launch; up GEMM + glu_fwd launch vs ONE launch (the
whole-step A/B is `UB200_FUSED_GLU=0/1 python bench.py`).
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
    act = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # 0 SwiGLU, 1 GEGLU-tanh, 2 GEGLU-erf
    two_fwd = [K.swiglu_fg_kernel, K.geglu_approx_forward_kernel, K.geglu_exact_forward_kernel][act]
    two_bwd = [K.swiglu_DWf_DW_dfg_kernel, K.geglu_approx_backward_kernel, K.geglu_exact_backward_kernel][act]
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
        two_bwd(out, e, g)

    def fwd_two():
        KU.gemm(T_, I, fsegs, out)
        two_fwd(e.view(1, T_, I), out.view(1, T_, I))

    variants = {
        "bwd: DW gemm alone": lambda: KU.gemm(T_, I, bsegs, out, b_mn=True),
        "bwd: DW gemm + glu_bwd (2 launches)": bwd_two,
        "bwd: gemm_glu (1 launch)": lambda: KU.gemm_glu(L.GLU_EPI_BWD, act, T_, I, bsegs, out, e, g, b_mn=True),
        "fwd: up gemm alone": lambda: KU.gemm(T_, I, fsegs, out),
        "fwd: up gemm + glu_fwd (2 launches)": fwd_two,
        "fwd: gemm_glu (1 launch)": lambda: KU.gemm_glu(L.GLU_EPI_FWD, act, T_, I, fsegs, out2, e, out),
    }
    ts = {k: [] for k in variants}
    for rnd in range(11):
        for k, fn in variants.items():
            t = timed(fn)
            if rnd >= 2:
                ts[k].append(t)
            e.normal_(); g.normal_()          # in-place outputs drift otherwise
    for k in variants:
        print(json.dumps({"act": act, "variant": k, "ms": round(med(ts[k]), 4)}), flush=True)


if __name__ == "__main__":
    main()
