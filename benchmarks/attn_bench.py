"""Attention library comparison at cfg2 (B=4, S=2048, Hq=32, Hk=8, D=128, causal, bf16):
flash-attn 2 vs torch SDPA backends (cuDNN / flash / efficient), fwd and fwd+bwd, on the strided
[B,S,H,D] projection-buffer views the model uses."""
import json

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

B, S, Hq, Hk, D = 4, 2048, 32, 8, 128
dev = "cuda"
torch.manual_seed(0)
qkv = [torch.randn(B, S, h * D, device=dev, dtype=torch.bfloat16, requires_grad=True) for h in (Hq, Hk, Hk)]
flops_f = 4 * B * Hq * S * S * D / 2       # causal
flops_fb = flops_f * 3.5


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


def views():
    q = qkv[0].view(B, S, Hq, D); k = qkv[1].view(B, S, Hk, D); v = qkv[2].view(B, S, Hk, D)
    return q, k, v


def run(name, fwd):
    try:
        out = fwd()
        g = torch.randn_like(out)
        t_f = timeit(lambda: fwd())
        def fb():
            o = fwd()
            o.backward(g)
        t_fb = timeit(fb)
        print(json.dumps({"impl": name, "fwd_ms": round(t_f, 3), "fwd_TFLOPs": round(flops_f / t_f / 1e9, 1),
                          "fwd_bwd_ms": round(t_fb, 3), "fwd_bwd_TFLOPs": round(flops_fb / t_fb / 1e9, 1)}), flush=True)
        return out
    except Exception as ex:
        print(json.dumps({"impl": name, "error": repr(ex)[:300]}), flush=True)
        return None


from flash_attn import flash_attn_func
ref = run("flash_attn2", lambda: flash_attn_func(*views(), causal=True))
for bk, nm in ((SDPBackend.CUDNN_ATTENTION, "sdpa_cudnn"), (SDPBackend.FLASH_ATTENTION, "sdpa_flash")):
    def f(bk=bk):
        q, k, v = views()
        with sdpa_kernel(bk):
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                               is_causal=True, enable_gqa=True)
        return o.transpose(1, 2)
    o = run(nm, f)
    if o is not None and ref is not None:
        print(json.dumps({"impl": nm, "max_abs_diff_vs_fa2": (o.float() - ref.float()).abs().max().item(),
                          "out_contiguous_bshd": o.is_contiguous()}), flush=True)
