"""GPU tests of the grouped persistent tcgen05 GEMM (csrc/gemm_grouped.cu, `ub200_gemm_grouped`):
in-launch producer -> consumer dependencies (per-row-block and whole-output waits), split-K with the
last-arriver fixed-order reduction (bit-reproducible), mixed tile widths and layouts, tails, and the
self-cleaning scratch -- against fp32 torch on the same operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16


def _r(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(BF)


def _check(name, out, ref, tol=6e-3):
    err = (out.float() - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item() + 1e-6, (name, err, ref.abs().max().item())


def _scratch_is_clean():
    from unsloth_b200.kernels import utils as KU
    torch.cuda.synchronize()
    for buf in KU._GROUP_SCRATCH.values():
        assert int(buf.abs().sum().item()) == 0, "grouped-GEMM scratch was not left clean"


@pytest.mark.parametrize("T_,in_f,outs", [(1024, 512, (512, 256, 256)), (300, 256, (200,)), (8192, 4096, (4096, 1024, 1024))])
def test_forward_group_with_in_launch_rank_block(T_, in_f, outs):
    """XA = X @ A_cat^T produced INSIDE the launch; each Y_i = X @ W_i^T + XA @ B_i^T waits for the
    row block of XA it needs."""
    from unsloth_b200.kernels.utils import Problem, gemm_grouped
    torch.manual_seed(T_ + in_f)
    X = _r(T_, in_f)
    A_cat = torch.zeros(64, in_f, device=DEV, dtype=BF)
    A_cat[:16 * len(outs)] = _r(16 * len(outs), in_f, scale=in_f ** -0.5)
    XA = torch.empty(T_, 64, device=DEV, dtype=BF)
    probs = [Problem(T_, 64, [(X, A_cat, in_f)], XA, signals=True, tag="rank")]
    Ws, Bs, Ys = [], [], []
    for i, o in enumerate(outs):
        W = _r(o, in_f, scale=0.05)
        Bp = torch.zeros(o, 64, device=DEV, dtype=BF)
        Bp[:, 16 * i:16 * i + 16] = _r(o, 16, scale=0.05)
        Y = torch.empty(T_, o, device=DEV, dtype=BF)
        probs.append(Problem(T_, o, [(X, W, in_f), (XA, Bp, 64, 16)], Y, wait=(0, 1, False)))
        Ws.append(W); Bs.append(Bp); Ys.append(Y)
    gemm_grouped(probs)
    XA_ref = X.float() @ A_cat.float().t()
    _check("XA", XA, XA_ref)
    for i, o in enumerate(outs):
        ref = X.float() @ Ws[i].float().t() + XA.float() @ Bs[i].float().t()
        _check("Y%d" % i, Ys[i], ref)
    first = [y.clone() for y in Ys]
    gemm_grouped(probs)                                    # same scratch again: must be clean + bitwise equal
    for a, b in zip(first, Ys):
        assert torch.equal(a, b)
    _scratch_is_clean()


@pytest.mark.parametrize("T_,in_f,outs", [(2048, 1024, (1024, 256)), (515, 256, (384,)), (8192, 4096, (4096, 1024, 1024))])
def test_backward_group_dA_dB_in_kernel(T_, in_f, outs):
    """G = sum_i dY_i @ sB_i | dB_i = s dY_i^T @ XA (split-K) | dX = sum_i dY_i @ W_i + G @ A_cat
    (waits per row block) | dA^T = X^T @ G (split-K, waits for all of G) -- one launch."""
    from unsloth_b200.kernels.fast_lora import _split_k_grouped
    from unsloth_b200.kernels.utils import Problem, gemm_grouped
    torch.manual_seed(T_ + in_f + 1)
    X = _r(T_, in_f)
    XA = torch.zeros(T_, 64, device=DEV, dtype=BF)
    XA[:, :16 * len(outs)] = _r(T_, 16 * len(outs), scale=0.3)
    A_cat = torch.zeros(64, in_f, device=DEV, dtype=BF)
    A_cat[:16 * len(outs)] = _r(16 * len(outs), in_f, scale=in_f ** -0.5)
    dYs = [_r(T_, o, scale=0.1) for o in outs]
    Ws = [_r(o, in_f, scale=0.05) for o in outs]
    Bps = []
    for i, o in enumerate(outs):
        Bp = torch.zeros(o, 64, device=DEV, dtype=BF)
        Bp[:, 16 * i:16 * i + 16] = _r(o, 16, scale=0.05)
        Bps.append(Bp)
    sk = max(2, _split_k_grouped(T_))
    G = torch.empty(T_, 64, device=DEV, dtype=BF)
    probs = [Problem(T_, 64, [(dY, Bp, o) for dY, Bp, o in zip(dYs, Bps, outs)], G, b_mn=True, signals=True, tag="rank")]
    dBs = []
    for dY, o in zip(dYs, outs):
        dB = torch.empty(o, 64, device=DEV, dtype=torch.float32)
        probs.append(Problem(o, 64, [(dY, XA, T_)], dB, a_mn=True, b_mn=True, alpha=2.0, split_k=sk, tag="rank"))
        dBs.append(dB)
    dX = torch.empty(T_, in_f, device=DEV, dtype=BF)
    segs = [(dY, W, o) for dY, W, o in zip(dYs, Ws, outs)] + [(G, A_cat, 64, 16 * len(outs))]
    probs.append(Problem(T_, in_f, segs, dX, b_mn=True, wait=(0, len(segs) - 1, False)))
    dA = torch.empty(in_f, 64, device=DEV, dtype=torch.float32)
    probs.append(Problem(in_f, 64, [(X, G, T_)], dA, a_mn=True, b_mn=True, split_k=sk, wait=(0, 0, True), tag="rank"))
    gemm_grouped(probs)
    G_ref = sum(dY.float() @ Bp.float() for dY, Bp in zip(dYs, Bps))
    _check("G", G, G_ref)
    for i, (dY, dB) in enumerate(zip(dYs, dBs)):
        _check("dB%d" % i, dB, 2.0 * (dY.float().t() @ XA.float()), tol=2e-3)
    dX_ref = sum(dY.float() @ W.float() for dY, W in zip(dYs, Ws)) + G.float() @ A_cat.float()
    _check("dX", dX, dX_ref)
    _check("dA", dA, X.float().t() @ G.float(), tol=2e-3)
    snap = [t.clone() for t in (G, dX, dA, *dBs)]
    for _ in range(3):
        gemm_grouped(probs)
    for a, b in zip(snap, (G, dX, dA, *dBs)):
        assert torch.equal(a, b), "grouped launch is not bit-reproducible"
    _scratch_is_clean()


def test_grouped_matches_per_gemm_schedule_on_lora_functions(monkeypatch):
    """The LoRA autograd functions through the grouped launches vs the round-1 one-launch-per-GEMM
    schedule (UB200_GROUPED=0): same operands, same products -- outputs and gradients agree to the
    last few ulps (the split-K partition of the token reductions differs)."""
    import unsloth_b200.kernels as K
    from unsloth_b200.nf4 import quantize_nf4
    torch.manual_seed(77)
    T_, H, I, r, s = 1024, 512, 1408, 16, 1.0
    X = _r(2, T_ // 2, H)
    dY = _r(2, T_ // 2, H, scale=0.1)

    def mk(o, i):
        W = _r(o, i, scale=0.03)
        p, q = quantize_nf4(W)
        return p, q, ((torch.rand(r, i, device=DEV) * 2 - 1) / i ** 0.5), torch.randn(o, r, device=DEV) * 0.03
    gate, up, down = mk(I, H), mk(I, H), mk(H, I)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("UB200_GROUPED", mode)
        P = [t.clone().requires_grad_() for p_ in (gate, up, down) for t in (p_[2], p_[3])]
        Xg = X.clone().requires_grad_()
        out = K.LoRA_MLP.apply(Xg * 1, gate[0], gate[1], P[0], P[1], s, up[0], up[1], P[2], P[3], s,
                               down[0], down[1], P[4], P[5], s, K.swiglu_fg_kernel, K.swiglu_DWf_DW_dfg_kernel, True)
        out.backward(dY)
        res[mode] = [out.detach().float(), Xg.grad.float()] + [p_.grad.float() for p_ in P]
    for a, b in zip(res["1"], res["0"]):
        assert (a - b).abs().max().item() <= 4e-3 * b.abs().max().item() + 1e-7
