"""Tuning probe: SwiGLU fwd / bwd at cfg2 size for each UB200_GLU_VARIANT (one subprocess each, the
variant is read once per process), CUDA events, L2 flushed, median of 15."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import unsloth_b200.kernels as K
    T_, I = 8192, 14336
    e = torch.randn(1, T_, I, device="cuda").to(torch.bfloat16)
    g = torch.randn(1, T_, I, device="cuda").to(torch.bfloat16)
    DW = torch.randn(T_, I, device="cuda").to(torch.bfloat16)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def t(fn):
        for _ in range(3):
            fn()
        xs = []
        for _ in range(15):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            xs.append(a.elapsed_time(b))
        return sorted(xs)[7]
    f = t(lambda: K.swiglu_fg_kernel(e, g))
    b = t(lambda: K.swiglu_DWf_DW_dfg_kernel(DW, e.view(T_, I), g.view(T_, I)))
    n = T_ * I * 2
    print("GLU", json.dumps({"variant": os.environ.get("UB200_GLU_VARIANT"), "fwd_ms": round(f, 4), "bwd_ms": round(b, 4),
                             "fwd_TBps": round(3 * n / f / 1e9, 2), "bwd_TBps": round(6 * n / b / 1e9, 2)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for v in ("1", "4", "7"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, UB200_GLU_VARIANT=v),
                               capture_output=True, text=True)
            print("\n".join(l for l in r.stdout.splitlines() if l.startswith("GLU")) or r.stderr[-1500:], flush=True)
