"""Numerics probe (GPU box): where do the 16-bit roundings of the reference's RoPE kernel fall when
Triton compiles it natively?  Runs the reference kernel and the C-ABI kernel in each rounding mode
(UB200_ROPE_MODE, csrc/rope.cu) on identical tensors and prints the fraction of differing elements.

    python benchmarks/probe_ref_numerics.py            # spawns one subprocess per (dtype, mode)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(dtype_name, mode):
    import torch
    from oracle import ref_shim
    refk = ref_shim.load_reference_kernels_native()
    from unsloth_b200.kernels import fast_rope_embedding
    dt = getattr(torch, dtype_name)
    torch.manual_seed(7)
    B, S, Hq, Hk, D = 2, 1024, 32, 8, 128
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.outer(torch.arange(S).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos().to("cuda", dt), emb.sin().to("cuda", dt)
    q0 = torch.randn(B, S, Hq * D, device="cuda").to(dt)
    k0 = torch.randn(B, S, Hk * D, device="cuda").to(dt)
    out = {"dtype": dtype_name, "mode": mode}
    for form in ("noindex", "indexed"):
        idx = None
        if form == "indexed":
            idx = torch.arange(S, device="cuda", dtype=torch.int32).repeat(B)

        def run(fn):
            Q = q0.clone().view(B, S, Hq, D).transpose(1, 2)
            K = k0.clone().view(B, S, Hk, D).transpose(1, 2)
            with torch.no_grad():
                Qo, Ko = fn(Q, K, cos, sin, idx)
            return Qo.contiguous(), Ko.contiguous()
        r, o = run(refk.rope_embedding.fast_rope_embedding), run(fast_rope_embedding)
        out[form] = {"Q_differ": (r[0] != o[0]).float().mean().item(),
                     "K_differ": (r[1] != o[1]).float().mean().item()}
    print("PROBE", json.dumps(out), flush=True)


def main():
    if len(sys.argv) == 3:
        return child(sys.argv[1], int(sys.argv[2]))
    for dtype_name in ("bfloat16", "float16"):
        for mode in (0, 1, 2, 3):
            env = dict(os.environ, UB200_ROPE_MODE=str(mode))
            if dtype_name == "bfloat16" and mode == 0:
                env.update(TRITON_KERNEL_DUMP="1", TRITON_DUMP_DIR=os.path.join(ROOT, "gpurun_out", "triton_dump"))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), dtype_name, str(mode)], env=env,
                               capture_output=True, text=True)
            print("\n".join(l for l in r.stdout.splitlines() if l.startswith("PROBE")) or r.stderr[-2000:], flush=True)


if __name__ == "__main__":
    main()
