"""Host orchestration of `unsloth_b200.kernels.fast_lora` (which GEMM is issued with which operands,
layouts, rank-block offsets, scales and output slices) checked ON CPU against the golden vectors the
REFERENCE's LoRA_MLP / LoRA_QKV / LoRA_W produced (tests/golden/lora_*.npz).

The C-ABI launches themselves cannot run without a GPU, so this test -- and only this test --
swaps the thin Python launch wrappers (`gemm`, `cast_pad`) for arithmetic test doubles with the
documented semantics of `ub200_gemm` / `ub200_cast_pad_2d` (include/unsloth_b200.h) and lets
`require_cuda` pass.  Everything above the wrappers is the shipped code.  The GPU suite checks the
same functions through the real library; this one guards refactors of the host side."""
import numpy as np
import pytest
import torch

from oracle import restate as R


def T(x):
    return torch.from_numpy(np.asarray(x))


def gemm_double(M, N, segs, out, a_mn=False, b_mn=False, alpha=1.0, accumulate=False, split_k=1,
                block_n=0, cta_group=0):
    """out[M,N] (+)= alpha * sum_s A_s . B_s^T ; a_mn: A stored [K,M]; b_mn: B stored [K,N]."""
    acc = torch.zeros(M, N, dtype=torch.float64)
    for sg in segs:
        A, B, K = sg[0], sg[1], sg[2]
        assert A.dim() == 2 and B.dim() == 2 and A.stride(-1) == 1 and B.stride(-1) == 1
        Am = A.t() if a_mn else A
        Bm = B.t() if b_mn else B
        assert Am.shape[0] >= M and Bm.shape[0] >= N and Am.shape[1] >= K and Bm.shape[1] >= K, \
            (tuple(Am.shape), tuple(Bm.shape), M, N, K)
        acc += Am[:M, :K].double() @ Bm[:N, :K].double().t()
    acc *= alpha
    if accumulate:
        acc += out.double()
    out.copy_(acc.to(out.dtype))
    return out


def gemm_grouped_double(problems):
    """Documented semantics of ub200_gemm_grouped: the problems in list order (a dependency may only
    point at an EARLIER problem that signals; the dependent operand must be that problem's output)."""
    for i, pr in enumerate(problems):
        if pr.wait is not None:
            dep, seg, whole = pr.wait
            if not isinstance(dep, int):                     # producer given by identity
                dep = next((k for k, q in enumerate(problems) if q is dep), None)
            if dep is not None:                              # None: produced by an EARLIER launch
                assert 0 <= dep < i and problems[dep].signals and 0 <= seg < len(pr.segs)
                operand = pr.segs[seg][1] if whole else pr.segs[seg][0]
                assert operand.data_ptr() == problems[dep].out.data_ptr()
        assert not (pr.b_mn and pr.block_n < 128)
        gemm_double(pr.M, pr.N, pr.segs, pr.out, pr.a_mn, pr.b_mn, pr.alpha, pr.accumulate)
    return [pr.out for pr in problems]


def cast_pad_double(src, dst, row_off=0, col_off=0, scale=1.0, transpose=False):
    blk = (src.t() if transpose else src).to(torch.float64) * scale
    dst.zero_()
    dst[row_off:row_off + blk.shape[0], col_off:col_off + blk.shape[1]] = blk.to(dst.dtype)
    return dst


@pytest.fixture(params=["rank-group+dense", "one-launch", "per-gemm"])
def host_doubles(monkeypatch, request):
    import unsloth_b200._lib as L
    monkeypatch.setenv("UB200_GROUPED", "0" if request.param == "per-gemm" else "1")
    monkeypatch.setenv("UB200_GROUPED_BWD", "1" if request.param == "one-launch" else "2")
    import unsloth_b200.kernels.fast_lora as FL
    import unsloth_b200.kernels.utils as KU
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    for mod in (FL, KU):
        monkeypatch.setattr(mod, "gemm", gemm_double)
        monkeypatch.setattr(mod, "cast_pad", cast_pad_double)
        monkeypatch.setattr(mod, "gemm_grouped", gemm_grouped_double)
    KU.set_keep_dequant(False)
    yield FL
    KU.set_keep_dequant(None)


def _param(x):
    return torch.nn.Parameter(T(x).clone())


def _close(a, b, atol=2e-5):
    torch.testing.assert_close(a.detach().float(), T(b).float(), rtol=1e-5, atol=atol)


@pytest.mark.parametrize("keep", [False, True])
def test_lora_qkv_and_w_orchestration(golden, host_doubles, keep):
    FL = host_doubles
    import unsloth_b200.kernels.utils as KU
    KU.set_keep_dequant(keep)
    g = golden("lora_qkv")
    s = float(g["s"])
    X = T(g["X"]).clone().requires_grad_()
    P = {n: _param(g[n]) for n in ("qA", "qB", "kA", "kB", "vA", "vB")}
    Q, K, V = FL.LoRA_QKV.apply(X, T(g["qW"]), None, P["qA"], P["qB"], s, T(g["kW"]), None, P["kA"], P["kB"], s,
                                T(g["vW"]), None, P["vA"], P["vB"], s, False)
    _close(Q, g["Q"]); _close(K, g["K"]); _close(V, g["V"])
    torch.autograd.backward([Q, K, V], [T(g["dQ"]).clone(), T(g["dK"]).clone(), T(g["dV"]).clone()])
    _close(X.grad, g["dX"])
    for n in ("q", "k", "v"):
        _close(P[n + "A"].grad, g["d_%sA" % n]); _close(P[n + "B"].grad, g["d_%sB" % n])

    g = golden("lora_w")
    X = T(g["X"]).clone().requires_grad_()
    A, B = _param(g["oA"]), _param(g["oB"])
    O = FL.LoRA_W.apply(X, T(g["oW"]), None, A, B, float(g["s"]))
    _close(O, g["O"])
    O.backward(T(g["dY"]).clone())
    _close(X.grad, g["dX"]); _close(A.grad, g["d_oA"]); _close(B.grad, g["d_oB"])


@pytest.mark.parametrize("act", ["swiglu", "geglu_approx"])
def test_lora_mlp_orchestration(golden, host_doubles, act):
    FL = host_doubles
    g = golden("lora_mlp_" + act)
    s = float(g["s"])
    fwd, bwd = {"swiglu": (R.swiglu_fwd, R.swiglu_bwd),
                "geglu_approx": (R.geglu_approx_fwd, R.geglu_approx_bwd)}[act]
    X = T(g["X"]).clone().requires_grad_()
    P = {n: _param(g[n]) for n in ("gA", "gB", "uA", "uB", "dA", "dB")}
    out = FL.LoRA_MLP.apply(X, T(g["gW"]), None, P["gA"], P["gB"], s, T(g["uW"]), None, P["uA"], P["uB"], s,
                            T(g["dW"]), None, P["dA"], P["dB"], s, fwd, bwd, False)
    _close(out, g["out"])
    out.backward(T(g["dY"]).clone())
    _close(X.grad, g["dX"])
    for n in ("g", "u", "d"):
        _close(P[n + "A"].grad, g["d_%sA" % n]); _close(P[n + "B"].grad, g["d_%sB" % n])


def test_lora_without_adapters(golden, host_doubles):
    """A = B = s = None (adapters disabled / merged, kernels/utils.py:365-371) must reduce to the base
    projection with no LoRA gradients."""
    FL = host_doubles
    g = golden("lora_w")
    X = T(g["X"]).clone().requires_grad_()
    O = FL.LoRA_W.apply(X, T(g["oW"]), None, None, None, None)
    _close(O, T(g["X"]) @ T(g["oW"]).t())
    O.backward(T(g["dY"]).clone())
    _close(X.grad, T(g["dY"]) @ T(g["oW"]))
