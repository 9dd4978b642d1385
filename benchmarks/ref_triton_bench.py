"""BASELINE.md B1 + B3: the reference's own GPU code timed on the B200 next to ours.

    python benchmarks/ref_triton_bench.py [--ops] [--step] [--layers N]

--ops  : per-op table at cfg2 sizes (Llama-3-8B, T = 4 x 2048, bf16): the reference's Triton kernels
         / LoRA autograd functions (compiled natively, oracle/ref_shim) vs the C-ABI path on
         identical tensors; CUDA events, 3 warm-ups, L2 flushed between iterations, median of 10.
--step : the full QLoRA training step through the reference's kernels (benchmarks/ref_composite.py)
         vs ours (eager and CUDA-graph replay) on the same model object.
One JSON line per row.  Needs the reference install under baseline/_ref.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEV, BF = "cuda", torch.bfloat16
_flush = None


def timeit(fn, iters=10, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        _flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def row(op, ms_ref, ms_ours, **kw):
    print(json.dumps({"op": op, "reference_ms": round(ms_ref, 4), "ours_ms": round(ms_ours, 4),
                      "speedup": round(ms_ref / ms_ours, 2), **kw}), flush=True)


class Norm:
    def __init__(self, w, eps=1e-5):
        self.weight, self.variance_epsilon = w, eps


def ops(refk):
    import unsloth_b200.kernels as K
    from unsloth_b200.kernels.cross_entropy_loss import Fast_CrossEntropyLoss
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(0)
    T_, H, I, V = 8192, 4096, 14336, 128256
    X = torch.randn(4, 2048, H, device=DEV).to(BF)
    W = torch.ones(H, device=DEV, dtype=BF)
    dY = torch.randn(4, 2048, H, device=DEV).to(BF)

    def fb(fn):                                    # forward + backward of a norm-like op
        def run():
            x = X.clone().requires_grad_()
            fn(Norm(W), x).backward(dY.clone())
        return run
    row("rms_layernorm fwd+bwd [8192,4096] (incl. 2 clones)", timeit(fb(refk.rms_layernorm.fast_rms_layernorm)),
        timeit(fb(K.fast_rms_layernorm)))
    with torch.no_grad():
        row("rms_layernorm fwd [8192,4096]", timeit(lambda: refk.rms_layernorm.fast_rms_layernorm(Norm(W), X)),
            timeit(lambda: K.fast_rms_layernorm(Norm(W), X)), bytes=2 * T_ * H * 2)
    # RoPE on the projection buffers (in place)
    D = 128
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(2048).float(), inv)
    cos, sin = torch.cat((fr, fr), -1).cos().to(DEV, BF), torch.cat((fr, fr), -1).sin().to(DEV, BF)
    q = torch.randn(4, 2048, 32 * D, device=DEV).to(BF)
    k = torch.randn(4, 2048, 8 * D, device=DEV).to(BF)
    Q, Kt = q.view(4, 2048, 32, D).transpose(1, 2), k.view(4, 2048, 8, D).transpose(1, 2)
    with torch.no_grad():
        row("rope fwd Q+K [4,32+8,2048,128]", timeit(lambda: refk.rope_embedding.fast_rope_embedding(Q, Kt, cos, sin)),
            timeit(lambda: K.fast_rope_embedding(Q, Kt, cos, sin)), bytes=2 * T_ * 40 * D * 2)
    e = torch.randn(1, T_, I, device=DEV).to(BF)
    g = torch.randn(1, T_, I, device=DEV).to(BF)
    row("swiglu fwd [8192,14336]", timeit(lambda: refk.swiglu.swiglu_fg_kernel(e, g)),
        timeit(lambda: K.swiglu_fg_kernel(e, g)), bytes=3 * T_ * I * 2)
    DW = torch.randn(T_, I, device=DEV).to(BF)
    e2, g2 = e.view(T_, I), g.view(T_, I)
    row("swiglu bwd [8192,14336]", timeit(lambda: refk.swiglu.swiglu_DWf_DW_dfg_kernel(DW, e2, g2)),
        timeit(lambda: K.swiglu_DWf_DW_dfg_kernel(DW, e2, g2)), bytes=6 * T_ * I * 2)
    del e, g, DW, e2, g2
    logits = torch.randn(2048, V, device=DEV).to(BF)
    labels = torch.randint(0, V, (2048,), device=DEV)

    def ce(fn):
        def run():
            lg = logits.clone().requires_grad_()
            fn(lg, labels, 0, 0).sum().backward()
        return run
    row("cross_entropy fwd+bwd [2048,128256] (incl. clone)", timeit(ce(refk.cross_entropy_loss.Fast_CrossEntropyLoss.apply)),
        timeit(ce(Fast_CrossEntropyLoss.apply)))
    del logits
    Wg = (torch.randn(I, H, device=DEV) * 0.02).to(BF)
    packed, qs = quantize_nf4(Wg)
    row("fast_dequantize NF4 [14336,4096]", timeit(lambda: refk.utils.fast_dequantize(packed, qs, use_global_buffer=True)),
        timeit(lambda: K.fast_dequantize(packed, qs, use_global_buffer=True)),
        note="reference host path = 2 launches + torch add on the SAME exported dequant symbols; ours = 1 launch")
    A = (torch.rand(16, H, device=DEV) * 2 - 1) / H ** 0.5
    B = torch.randn(I, 16, device=DEV) * 0.02
    with torch.no_grad():
        row("matmul_lora NF4 gate_proj [8192,4096]->[8192,14336]",
            timeit(lambda: refk.utils.matmul_lora(X, packed, qs, A, B, 1.0)),
            timeit(lambda: K.matmul_lora(X, packed, qs, A, B, 1.0)), flops=2 * T_ * H * I)

    def mk(o, i):
        Wd = (torch.randn(o, i, device=DEV) * 0.02).to(BF)
        p, s_ = quantize_nf4(Wd)
        return p, s_, ((torch.rand(16, i, device=DEV) * 2 - 1) / i ** 0.5).requires_grad_(), \
            (torch.randn(o, 16, device=DEV) * 0.02).requires_grad_()
    gate, up, down = mk(I, H), mk(I, H), mk(H, I)

    def mlp(mod, f, b):
        def run():
            x = X.clone().requires_grad_()
            out = mod.apply(x * 1, gate[0], gate[1], gate[2], gate[3], 1.0, up[0], up[1], up[2], up[3], 1.0,
                            down[0], down[1], down[2], down[3], 1.0, f, b, True)
            out.backward(dY)
        return run
    fl = 3 * 3 * 2 * T_ * H * I
    row("LoRA_MLP NF4 fwd+bwd (T=8192,H=4096,I=14336,r=16)",
        timeit(mlp(refk.fast_lora.LoRA_MLP, refk.swiglu.swiglu_fg_kernel, refk.swiglu.swiglu_DWf_DW_dfg_kernel), iters=5),
        timeit(mlp(K.LoRA_MLP, K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel), iters=5), flops=fl)
    qp, kp, vp = mk(H, H), mk(1024, H), mk(1024, H)

    def qkv(mod):
        dQ, dK = dY, dY[..., :1024].contiguous()

        def run():
            x = X.clone().requires_grad_()
            Q_, K_, V_ = mod.apply(x * 1, qp[0], qp[1], qp[2], qp[3], 1.0, kp[0], kp[1], kp[2], kp[3], 1.0,
                                   vp[0], vp[1], vp[2], vp[3], 1.0, True)
            torch.autograd.backward([Q_, K_, V_], [dQ, dK, dK])
        return run
    row("LoRA_QKV NF4 fwd+bwd (T=8192,H=4096,kv=1024,r=16)", timeit(qkv(refk.fast_lora.LoRA_QKV), iters=5),
        timeit(qkv(K.LoRA_QKV), iters=5), flops=2 * 2 * T_ * H * (H + 2048))


def step(layers, steps, warmup):
    from benchmarks.ref_composite import time_reference_step
    from unsloth_b200.ddp import FlatLoRABucket
    from unsloth_b200.graph import GraphedTrainStep
    from unsloth_b200.patch import build_qlora_model, lora_parameters
    dev = torch.device("cuda", 0)
    model = build_qlora_model("llama-3-8b", r=16, lora_alpha=16, device=dev, seed=3407, num_hidden_layers=layers)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, model.config.vocab_size, (4, 4, 2048), generator=g).to(dev)
    lab = ids.clone()
    lab[torch.rand(lab.shape, generator=g).to(dev) < 0.1] = -100
    for attn in ("flash", "sdpa"):
        r = time_reference_step(model, ids, lab, steps=steps, warmup=warmup, attention=attn)
        print(json.dumps({"step": "reference GPU path (Triton kernels + cuBLAS LoRA schedule, eager)",
                          "layers": model.config.num_hidden_layers, **{k: (round(v, 2) if isinstance(v, float) else v)
                                                                      for k, v in r.items()}}), flush=True)
    bucket = FlatLoRABucket(lora_parameters(model), lr=2e-4, weight_decay=0.01)
    graphed = GraphedTrainStep(model, bucket, 4, 2048, dev)
    for i in range(warmup):
        graphed.step(ids[i % 4], lab[i % 4])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        loss = graphed.step(ids[i % 4], lab[i % 4])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    print(json.dumps({"step": "ours (CUDA-graph replay)", "layers": model.config.num_hidden_layers,
                      "tokens_per_s": round(8192 / ms * 1e3, 1), "ms_per_step": round(ms, 2),
                      "loss_last": round(float(loss.item()), 4)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", action="store_true")
    ap.add_argument("--step", action="store_true")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    from oracle import ref_shim
    refk = ref_shim.load_reference_kernels_native()
    torch.backends.cuda.matmul.allow_tf32 = False
    if a.ops or not a.step:
        ops(refk)
    if a.step:
        step(a.layers, a.steps, a.warmup)


if __name__ == "__main__":
    main()
