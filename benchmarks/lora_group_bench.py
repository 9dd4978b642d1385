"""A/B of the LoRA launch schedules at cfg2 sizes (Llama-3-8B, T = 4 x 2048, r = 16, NF4 + bf16):
    per-gemm          UB200_GROUPED=0                      one launch per GEMM (round 1)
    fwd-fused         UB200_GROUPED_FWD=1, BWD=0           XA produced inside the projection launch
    one-launch        UB200_GROUPED_BWD=1                  every product of a backward phase in one launch
    rank-group+dense  UB200_GROUPED_BWD=2 (default)        rank-block products grouped, dense GEMM alone
CUDA events around fwd and bwd of LoRA_QKV / LoRA_W / LoRA_MLP, L2 flushed, median of 7."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unsloth_b200.kernels as K  # noqa: E402
from unsloth_b200.kernels import utils as KU  # noqa: E402
from unsloth_b200.nf4 import quantize_nf4  # noqa: E402

DEV, BF = "cuda", torch.bfloat16
MODES = {"per-gemm": dict(UB200_GROUPED="0"),
         "fwd-fused": dict(UB200_GROUPED="1", UB200_GROUPED_FWD="1", UB200_GROUPED_BWD="0"),
         "one-launch": dict(UB200_GROUPED="1", UB200_GROUPED_FWD="1", UB200_GROUPED_BWD="1"),
         "rank-group+dense": dict(UB200_GROUPED="1", UB200_GROUPED_FWD="1", UB200_GROUPED_BWD="2"),
         "bwd-rank-group-only": dict(UB200_GROUPED="1", UB200_GROUPED_FWD="0", UB200_GROUPED_BWD="2")}
_flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def time_once(make):
    fn_f, fn_b = make()
    _flush.fill_(1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); out = fn_f(); e[1].record(); fn_b(out); e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])


def main():
    torch.manual_seed(0)
    KU.set_keep_dequant(True)
    T_, H, I, KV, r = 8192, 4096, 14336, 1024, 16
    X = torch.randn(4, 2048, H, device=DEV).to(BF)
    dY = (torch.randn(4, 2048, H, device=DEV) * 0.1).to(BF)
    dKV = dY[..., :KV].contiguous()

    def mk(o, i):
        W = (torch.randn(o, i, device=DEV) * 0.02).to(BF)
        p, q = quantize_nf4(W)
        A = torch.nn.Parameter((torch.rand(r, i, device=DEV) * 2 - 1) / i ** 0.5)
        B = torch.nn.Parameter(torch.randn(o, r, device=DEV) * 0.02)
        return p, q, A, B
    qp, kp, vp, op = mk(H, H), mk(KV, H), mk(KV, H), mk(H, H)
    gate, up, down = mk(I, H), mk(I, H), mk(H, I)

    def qkv():
        x = X.clone().requires_grad_()
        f = lambda: K.LoRA_QKV.apply(x, qp[0], qp[1], qp[2], qp[3], 1.0, kp[0], kp[1], kp[2], kp[3], 1.0,
                                     vp[0], vp[1], vp[2], vp[3], 1.0, True)
        return f, lambda o: torch.autograd.backward(list(o), [dY, dKV, dKV])

    def wo():
        x = X.clone().requires_grad_()
        return (lambda: K.LoRA_W.apply(x, op[0], op[1], op[2], op[3], 1.0)), (lambda o: o.backward(dY))

    def mlp():
        x = X.clone().requires_grad_()
        f = lambda: K.LoRA_MLP.apply(x, gate[0], gate[1], gate[2], gate[3], 1.0, up[0], up[1], up[2], up[3], 1.0,
                                     down[0], down[1], down[2], down[3], 1.0, K.swiglu_fg_kernel,
                                     K.swiglu_DWf_DW_dfg_kernel, True)
        return f, lambda o: o.backward(dY)

    # modes INTERLEAVED round-robin (the part is power-capped: a mode timed later in a sequential
    # sweep runs hotter and slower), 2 warm-up rounds + 9 timed rounds, medians
    acc = {n: {t: ([], []) for t in ("qkv", "o", "mlp")} for n in MODES}
    for rnd in range(11):
        for name, env in MODES.items():
            for k in ("UB200_GROUPED", "UB200_GROUPED_FWD", "UB200_GROUPED_BWD"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for tag, mkfn in (("qkv", qkv), ("o", wo), ("mlp", mlp)):
                f, b = time_once(mkfn)
                if rnd >= 2:
                    acc[name][tag][0].append(f); acc[name][tag][1].append(b)
    for name in MODES:
        row = {"mode": name}
        tot = 0.0
        for tag in ("qkv", "o", "mlp"):
            f, b = med(acc[name][tag][0]), med(acc[name][tag][1])
            row[tag + "_fwd_ms"], row[tag + "_bwd_ms"] = round(f, 4), round(b, 4)
            tot += f + b
        row["layer_lora_total_ms"] = round(tot, 4)
        print(json.dumps(row), flush=True)
    # the dense kernel alone: gemm2_kernel<256> vs the grouped kernel on ONE dense problem
    from unsloth_b200.kernels.utils import Problem, gemm, gemm_grouped
    W = (torch.randn(H, H, device=DEV) * 0.02).to(BF)
    X2 = X.view(T_, H)
    Y = torch.empty(T_, H, device=DEV, dtype=BF)
    A_cat = torch.zeros(64, H, device=DEV, dtype=BF)
    XA = torch.empty(T_, 64, device=DEV, dtype=BF)
    Bp = torch.zeros(H, 64, device=DEV, dtype=BF)
    variants = {
        "gemm2 dense 8192x4096x4096": lambda: gemm(T_, H, [(X2, W, H)], Y),
        "grouped [dense]": lambda: gemm_grouped([Problem(T_, H, [(X2, W, H)], Y)]),
        "gemm2 dense+rank segment": lambda: gemm(T_, H, [(X2, W, H), (XA, Bp, 64)], Y),
        "grouped [dense+rank segment, no wait]": lambda: gemm_grouped([Problem(T_, H, [(X2, W, H), (XA, Bp, 64)], Y)]),
        "grouped [XA | dense waits]": lambda: gemm_grouped([Problem(T_, 64, [(X2, A_cat, H)], XA, signals=True, tag="rank"),
                                                             Problem(T_, H, [(X2, W, H), (XA, Bp, 64)], Y, wait=(0, 1, False))]),
        "grouped [XA] alone": lambda: gemm_grouped([Problem(T_, 64, [(X2, A_cat, H)], XA, signals=True, tag="rank")]),
        "gemm1 XA alone": lambda: gemm(T_, 64, [(X2, A_cat, H)], XA),
    }
    ts = {k: [] for k in variants}
    for rnd in range(12):
        for k, fn in variants.items():
            _flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if rnd >= 2:
                ts[k].append(e0.elapsed_time(e1))
    for k in variants:
        print(json.dumps({"variant": k, "ms": round(med(ts[k]), 4)}), flush=True)


if __name__ == "__main__":
    main()
