"""Attention at the BASELINE shapes: the tcgen05 kernel of this repo (csrc/attention.cu) next to the
libraries the reference's dispatcher can call (flash-attn 2, torch SDPA cuDNN / flash), fwd and
fwd+bwd, on the strided [B,S,H,D] projection-buffer views the model uses.

    python benchmarks/attn_bench.py [cfg2|cfg3|cfg5]    (default: all three)
cfg2: B=4 S=2048 32/8 heads D=128 causal            (Llama-3-8B)
cfg3: B=2 S=4096 32/8 heads D=128 window 2048       (Mistral-7B sliding window)
cfg5: B=1 S=8192 16/8 heads D=256 softcap 50, window 4096 / none   (Gemma-2-9B even / odd layers)"""
import json
import os
import sys

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = "cuda"
CFGS = {"cfg2": dict(B=4, S=2048, Hq=32, Hk=8, D=128, wl=-1, cap=0.0),
        "cfg3": dict(B=2, S=4096, Hq=32, Hk=8, D=128, wl=2048, cap=0.0),
        "cfg5-window": dict(B=1, S=8192, Hq=16, Hk=8, D=256, wl=4096, cap=50.0),
        "cfg5-global": dict(B=1, S=8192, Hq=16, Hk=8, D=256, wl=-1, cap=50.0)}


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def visible_pairs(S, wl):
    if wl < 0 or wl >= S:
        return S * (S + 1) / 2
    return (wl + 1) * (wl + 2) / 2 + (S - wl - 1) * (wl + 1)


def bench(name, c):
    B, S, Hq, Hk, D, wl, cap = c["B"], c["S"], c["Hq"], c["Hk"], c["D"], c["wl"], c["cap"]
    torch.manual_seed(0)
    qkv = [torch.randn(B, S, h * D, device=dev, dtype=torch.bfloat16, requires_grad=True) for h in (Hq, Hk, Hk)]
    scale = D ** -0.5
    flops_f = 4.0 * B * Hq * D * visible_pairs(S, wl)
    flops_fb = flops_f * 3.5

    def views():
        return qkv[0].view(B, S, Hq, D), qkv[1].view(B, S, Hk, D), qkv[2].view(B, S, Hk, D)

    def run(impl, fwd, bwd=True):
        try:
            out = fwd()
            g = torch.randn_like(out)
            with torch.no_grad():
                t_f = timeit(lambda: fwd())
            row = {"cfg": name, "impl": impl, "fwd_ms": round(t_f, 3), "fwd_TFLOPs": round(flops_f / t_f / 1e9, 1)}
            if bwd:
                def fb():
                    fwd().backward(g)
                t_fb = timeit(fb)
                row.update(fwd_bwd_ms=round(t_fb, 3), fwd_bwd_TFLOPs=round(flops_fb / t_fb / 1e9, 1))
            print(json.dumps(row), flush=True)
            return out
        except Exception as ex:
            print(json.dumps({"cfg": name, "impl": impl, "error": repr(ex)[:300]}), flush=True)
            return None

    from flash_attn import flash_attn_func
    ref = run("flash_attn2", lambda: flash_attn_func(*views(), softmax_scale=scale, causal=True,
                                                     window_size=(wl, wl) if wl >= 0 else (-1, -1), softcap=cap))
    if wl < 0 and not cap:
        def f():
            q, k, v = views()
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                   is_causal=True, enable_gqa=True, scale=scale)
            return o.transpose(1, 2)
        run("sdpa_cudnn", f)
    from unsloth_b200.kernels.attention import fast_attention
    for mode in ("own", "library"):
        os.environ["UB200_ATTN_BWD"] = mode
        o = run("unsloth_b200 tcgen05 (bwd: %s)" % mode, lambda: fast_attention(*views(), scale, (wl, wl), cap))
        if o is not None and ref is not None:
            print(json.dumps({"cfg": name, "max_abs_diff_vs_fa2": round((o.float() - ref.float()).abs().max().item(), 5)}),
                  flush=True)
    os.environ.pop("UB200_ATTN_BWD", None)


if __name__ == "__main__":
    which = sys.argv[1:] or list(CFGS)
    for n in which:
        for k, c in CFGS.items():
            if k.startswith(n):
                bench(k, c)
