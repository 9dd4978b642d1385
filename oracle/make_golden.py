"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REFERENCE's
own Triton kernels (imported from /root/reference through oracle/ref_shim.py under
TRITON_INTERPRET=1, fp32, CPU).  Run inside the build container:

    python oracle/make_golden.py

The fixtures hold inputs and reference outputs; tests/test_oracle_golden.py pins
oracle/restate.py against them and the GPU parity tests compare the CUDA path with
them (they travel to the GPU box; /root/reference does not).
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference_kernels  # noqa: E402

import torch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 3407  # the reference's default random_state (models/llama.py:2340)


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
            for k, v in d.items()}


def save(name, **tensors):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(tensors))
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in _np(tensors).items()})


class _Norm:
    def __init__(self, w, eps):
        self.weight = w
        self.variance_epsilon = eps


def gen_rmsnorm(k):
    g = torch.Generator().manual_seed(SEED)
    for name, T, H, gemma, eps in (("rms_llama_512", 7, 512, False, 1e-5),
                                   ("rms_llama_odd", 5, 200, False, 1e-6),
                                   ("rms_gemma_256", 6, 256, True, 1e-6)):
        X = torch.randn(2, T, H, generator=g)
        W = torch.randn(H, generator=g) * 0.5 + (0.0 if gemma else 1.0)
        dY = torch.randn(2, T, H, generator=g)
        Xr = X.clone().requires_grad_()
        Y = k.rms_layernorm.fast_rms_layernorm(_Norm(W, eps), Xr, gemma=gemma)
        Yc = Y.detach().clone()
        Y.backward(dY.clone())
        save(name, X=X, W=W, dY=dY, Y=Yc, dX=Xr.grad, eps=eps, gemma=int(gemma))


def _rope_tables(S, D, g, base=10000.0):
    inv = 1.0 / (base ** (torch.arange(0, D, 2).float() / D))
    t = torch.arange(S).float()
    fr = torch.outer(t, inv)
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def gen_rope(k):
    g = torch.Generator().manual_seed(SEED + 1)
    B, S, Hq, Hk, D = 2, 9, 6, 2, 64
    cos, sin = _rope_tables(16, D, g)
    Q = torch.randn(B, Hq, S, D, generator=g)
    K = torch.randn(B, Hk, S, D, generator=g)
    dQ = torch.randn(B, Hq, S, D, generator=g)
    dK = torch.randn(B, Hk, S, D, generator=g)
    # no-index path (fast_rope_embedding with rope_embedding_indices=None)
    Qr, Kr = Q.clone().requires_grad_(), K.clone().requires_grad_()
    Qo, Ko = k.rope_embedding.fast_rope_embedding(Qr, Kr, cos, sin)
    Qo_c, Ko_c = Qo.detach().clone(), Ko.detach().clone()
    torch.autograd.backward([Qo, Ko], [dQ.clone(), dK.clone()])
    save("rope_noindex", Q=Q, K=K, cos=cos, sin=sin, dQ=dQ, dK=dK,
         Qo=Qo_c, Ko=Ko_c, gQ=Qr.grad, gK=Kr.grad)
    # index path
    idx = torch.randint(0, 16, (B * S,), generator=g).to(torch.int32)
    Qr, Kr = Q.clone().requires_grad_(), K.clone().requires_grad_()
    Qo, Ko = k.rope_embedding.fast_rope_embedding(Qr, Kr, cos, sin, idx)
    Qo_c, Ko_c = Qo.detach().clone(), Ko.detach().clone()
    torch.autograd.backward([Qo, Ko], [dQ.clone(), dK.clone()])
    save("rope_index", Q=Q, K=K, cos=cos, sin=sin, idx=idx, dQ=dQ, dK=dK,
         Qo=Qo_c, Ko=Ko_c, gQ=Qr.grad, gK=Kr.grad)


def gen_ce(k):
    g = torch.Generator().manual_seed(SEED + 2)
    for name, T, V, softcap, scale in (("ce_v1000", 12, 1000, 0.0, 0.0),
                                       ("ce_v70000_chunked", 4, 70000, 0.0, 0.0),
                                       ("ce_softcap30", 8, 1500, 30.0, 0.0),
                                       ("ce_scale", 8, 777, 0.0, 0.0625)):
        logits = torch.randn(1, T, V, generator=g) * 4.0
        labels = torch.randint(0, V, (1, T), generator=g)
        labels[0, 1] = -100
        labels[0, T - 1] = -100
        lr = logits.clone().requires_grad_()
        loss = k.cross_entropy_loss.fast_cross_entropy_loss(lr, labels, softcap, scale)
        loss_c = loss.detach().clone()
        loss.backward()
        # NB: the reference writes the gradient *into* the logits buffer; .grad aliases it
        save(name, logits=logits, labels=labels, loss=loss_c, dlogits=lr.grad.clone(),
             softcap=softcap, scale=scale)


def gen_glu(k):
    g = torch.Generator().manual_seed(SEED + 3)
    B, S, I = 2, 5, 300
    e = torch.randn(B, S, I, generator=g) * 2
    up = torch.randn(B, S, I, generator=g)
    DW = torch.randn(B * S, I, generator=g)
    for name, fwd, bwd in (
            ("swiglu", k.swiglu.swiglu_fg_kernel, k.swiglu.swiglu_DWf_DW_dfg_kernel),
            ("geglu_approx", k.geglu.geglu_approx_forward_kernel, k.geglu.geglu_approx_backward_kernel),
            ("geglu_exact", k.geglu.geglu_exact_forward_kernel, k.geglu.geglu_exact_backward_kernel)):
        h = fwd(e.clone(), up.clone())
        h2, df, de = bwd(DW.clone(), e.reshape(-1, I).clone(), up.reshape(-1, I).clone())
        save(name, e=e, g=up, DW=DW, h=h, bh=h2, bdf=df, bde=de)


def _lora(g, out_f, in_f, r, s):
    W = torch.randn(out_f, in_f, generator=g) * 0.05
    A = (torch.rand(r, in_f, generator=g) * 2 - 1) / (in_f ** 0.5)
    B = torch.randn(out_f, r, generator=g) * 0.02
    return W, A.requires_grad_(), B.requires_grad_(), s


def gen_lora(k):
    g = torch.Generator().manual_seed(SEED + 4)
    Bz, S, H, I, r = 2, 6, 64, 160, 8
    fl = k.fast_lora
    # ---- LoRA_MLP (swiglu and geglu_approx)
    for act, fwd, bwd in (("swiglu", k.swiglu.swiglu_fg_kernel, k.swiglu.swiglu_DWf_DW_dfg_kernel),
                          ("geglu_approx", k.geglu.geglu_approx_forward_kernel,
                           k.geglu.geglu_approx_backward_kernel)):
        X = torch.randn(Bz, S, H, generator=g)
        dY = torch.randn(Bz, S, H, generator=g)
        gW, gA, gB, gs = _lora(g, I, H, r, 2.0)
        uW, uA, uB, us = _lora(g, I, H, r, 2.0)
        dW, dA, dB, ds = _lora(g, H, I, r, 2.0)
        Xr = X.clone().requires_grad_()
        out = fl.LoRA_MLP.apply(Xr, gW, None, gA, gB, gs, uW, None, uA, uB, us,
                                dW, None, dA, dB, ds, fwd, bwd, False)
        out_c = out.detach().clone()
        out.backward(dY.clone())
        save("lora_mlp_" + act, X=X, dY=dY, gW=gW, gA=gA, gB=gB, uW=uW, uA=uA, uB=uB,
             dW=dW, dA=dA, dB=dB, s=2.0, out=out_c, dX=Xr.grad,
             d_gA=gA.grad, d_gB=gB.grad, d_uA=uA.grad, d_uB=uB.grad, d_dA=dA.grad, d_dB=dB.grad)
    # ---- LoRA_QKV
    X = torch.randn(Bz, S, H, generator=g)
    nq, nk = 96, 32
    qW, qA, qB, qs = _lora(g, nq, H, r, 0.5)
    kW, kA, kB, ks = _lora(g, nk, H, r, 0.5)
    vW, vA, vB, vs = _lora(g, nk, H, r, 0.5)
    dQ = torch.randn(Bz, S, nq, generator=g)
    dK = torch.randn(Bz, S, nk, generator=g)
    dV = torch.randn(Bz, S, nk, generator=g)
    Xr = X.clone().requires_grad_()
    Q, K, V = fl.LoRA_QKV.apply(Xr, qW, None, qA, qB, qs, kW, None, kA, kB, ks,
                                vW, None, vA, vB, vs, False)
    Qc, Kc, Vc = Q.detach().clone(), K.detach().clone(), V.detach().clone()
    torch.autograd.backward([Q, K, V], [dQ.clone(), dK.clone(), dV.clone()])
    save("lora_qkv", X=X, dQ=dQ, dK=dK, dV=dV, qW=qW, qA=qA, qB=qB, kW=kW, kA=kA, kB=kB,
         vW=vW, vA=vA, vB=vB, s=0.5, Q=Qc, K=Kc, V=Vc, dX=Xr.grad,
         d_qA=qA.grad, d_qB=qB.grad, d_kA=kA.grad, d_kB=kB.grad, d_vA=vA.grad, d_vB=vB.grad)
    # ---- LoRA_W
    X = torch.randn(Bz, S, H, generator=g)
    oW, oA, oB, os_ = _lora(g, H, H, r, 1.0)
    dY = torch.randn(Bz, S, H, generator=g)
    Xr = X.clone().requires_grad_()
    O = fl.LoRA_W.apply(Xr, oW, None, oA, oB, os_)
    Oc = O.detach().clone()
    O.backward(dY.clone())
    save("lora_w", X=X, dY=dY, oW=oW, oA=oA, oB=oB, s=1.0, O=Oc, dX=Xr.grad,
         d_oA=oA.grad, d_oB=oB.grad)


def gen_reference_tests(k):
    """Known-answer material taken from the reference's OWN tests:
    * RMSNorm self-test shapes/seeds of kernels/rms_layernorm.py:301-342 (scaled down in
      batch/seqlen: the self-test compares against HF LlamaRMSNorm; we store the kernel output);
    * the packed-boundary label vector of tests/utils/test_packing.py:1489-1525.
    """
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
    for dim, seqlen, seed in ((512, 149, 3407), (1024, 61, 42)):
        torch.manual_seed(seed)
        ln = LlamaRMSNorm((dim,), eps=1e-5)
        torch.nn.init.uniform_(ln.weight)
        X = torch.randn((1, seqlen, dim), requires_grad=True)
        Y_hf = ln(X)
        YY = torch.randn((1, seqlen, dim))
        Y_hf.backward(YY)
        hf_grad = X.grad.clone()
        X2 = X.detach().clone().requires_grad_()
        Y = k.rms_layernorm.fast_rms_layernorm(ln, X2)
        Yc = Y.detach().clone()
        Y.backward(YY.clone())
        save("rms_selftest_%d" % dim, X=X.detach(), W=ln.weight.detach(), dY=YY, Y=Yc,
             dX=X2.grad, Y_hf=Y_hf.detach(), dX_hf=hf_grad, eps=1e-5, gemma=0)
    # tests/utils/test_packing.py:1489-1525: packed_seq_lengths=[3,5], labels 0..7 ->
    # after shift+boundary masking the CE sees [-100? ...]; the reference asserts the
    # *unshifted* masked labels equal [-100,1,2,-100,4,5,6,7].
    save("packing_labels", packed_seq_lengths=np.array([3, 5]),
         labels=np.arange(8), expected=np.array([-100, 1, 2, -100, 4, 5, 6, 7]))


def gen_fp16_rounding_points(k):
    """16-bit runs of the reference kernels (fp16: the interpreter has no bf16).  They pin the
    ROUNDING POINTS of SURVEY.md section 9 that fp32 vectors cannot see: `normed -> W.dtype`
    (rms_layernorm.py:57), `f -> g.dtype` (swiglu.py:42, geglu.py:161), the table-dtype products
    of RoPE (rope_embedding.py:129-158) and the in-place 16-bit CE gradient
    (cross_entropy_loss.py:245-276).  The GPU bf16 gates compare against the oracle's emulation
    of exactly these points."""
    g = torch.Generator().manual_seed(SEED + 7)
    H = 256
    X = torch.randn(2, 5, H, generator=g).half()
    W = (torch.randn(H, generator=g) * 0.5 + 1).half()
    dY = torch.randn(2, 5, H, generator=g).half()
    Xr = X.clone().requires_grad_()
    Y = k.rms_layernorm.fast_rms_layernorm(_Norm(W, 1e-5), Xr, gemma=False)
    Yc = Y.detach().clone()
    Y.backward(dY.clone())
    save("fp16_rms_llama", X=X, W=W, dY=dY, Y=Yc, dX=Xr.grad, eps=1e-5)
    I = 300
    e = (torch.randn(2, 5, I, generator=g) * 2).half()
    up = torch.randn(2, 5, I, generator=g).half()
    DW = torch.randn(10, I, generator=g).half()
    for name, fwd, bwd in (
            ("swiglu", k.swiglu.swiglu_fg_kernel, k.swiglu.swiglu_DWf_DW_dfg_kernel),
            ("geglu_approx", k.geglu.geglu_approx_forward_kernel, k.geglu.geglu_approx_backward_kernel),
            ("geglu_exact", k.geglu.geglu_exact_forward_kernel, k.geglu.geglu_exact_backward_kernel)):
        h = fwd(e.clone(), up.clone())
        h2, df, de = bwd(DW.clone(), e.reshape(-1, I).clone(), up.reshape(-1, I).clone())
        save("fp16_" + name, e=e, g=up, DW=DW, h=h, bh=h2, bdf=df, bde=de)
    B, S, Hq, Hk, D = 2, 9, 6, 2, 64
    cos, sin = _rope_tables(16, D, g)
    cos, sin = cos.half(), sin.half()
    Q = torch.randn(B, Hq, S, D, generator=g).half()
    K = torch.randn(B, Hk, S, D, generator=g).half()
    Qo, Ko = k.rope_embedding.fast_rope_embedding(Q.clone(), K.clone(), cos, sin)
    idx = torch.randint(0, 16, (B * S,), generator=g).to(torch.int32)
    Qi, Ki = k.rope_embedding.fast_rope_embedding(Q.clone(), K.clone(), cos, sin, idx)
    save("fp16_rope", Q=Q, K=K, cos=cos, sin=sin, idx=idx, Qo=Qo.detach(), Ko=Ko.detach(),
         Qi=Qi.detach(), Ki=Ki.detach())
    T_, V = 6, 1200
    logits = (torch.randn(1, T_, V, generator=g) * 4).half()
    labels = torch.randint(0, V, (1, T_), generator=g)
    labels[0, 2] = -100
    lr = logits.clone().requires_grad_()
    loss = k.cross_entropy_loss.fast_cross_entropy_loss(lr, labels, 0.0, 0.0)
    loss_c = loss.detach().clone()
    loss.backward()
    save("fp16_ce", logits=logits, labels=labels, loss=loss_c, dlogits=lr.grad.clone())


def gen_fp16_lora(k):
    """fp16 runs of LoRA_W / LoRA_QKV / LoRA_MLP (16-bit X and W, fp32 adapters as PEFT keeps
    them): pin the rounding points of `matmul_lora` (utils.py:1158-1168: base product -> 16 bit,
    XA -> 16 bit, addmm_ -> 16 bit) and of the LoRA gradients (fast_lora.py:172-204, 476-517,
    639-647) that SURVEY section 9 lists."""
    g = torch.Generator().manual_seed(SEED + 8)
    fl = k.fast_lora

    def lora(out_f, in_f, r, s):
        W = (torch.randn(out_f, in_f, generator=g) * 0.05).half()
        A = ((torch.rand(r, in_f, generator=g) * 2 - 1) / (in_f ** 0.5)).requires_grad_()
        B = (torch.randn(out_f, r, generator=g) * 0.02).requires_grad_()
        return W, A, B, s
    Bz, S, H, I, r = 2, 6, 64, 160, 8
    X = torch.randn(Bz, S, H, generator=g).half()
    dY = torch.randn(Bz, S, H, generator=g).half()
    oW, oA, oB, os_ = lora(H, H, r, 1.0)
    Xr = X.clone().requires_grad_()
    O = fl.LoRA_W.apply(Xr, oW, None, oA, oB, os_)
    Oc = O.detach().clone()
    O.backward(dY.clone())
    save("fp16_lora_w", X=X, dY=dY, oW=oW, oA=oA, oB=oB, s=1.0, O=Oc, dX=Xr.grad,
         d_oA=oA.grad, d_oB=oB.grad)
    nq, nk = 96, 32
    qW, qA, qB, qs = lora(nq, H, r, 0.5)
    kW, kA, kB, ks = lora(nk, H, r, 0.5)
    vW, vA, vB, vs = lora(nk, H, r, 0.5)
    dQ = torch.randn(Bz, S, nq, generator=g).half()
    dK = torch.randn(Bz, S, nk, generator=g).half()
    dV = torch.randn(Bz, S, nk, generator=g).half()
    Xr = X.clone().requires_grad_()
    Q, K, V = fl.LoRA_QKV.apply(Xr, qW, None, qA, qB, qs, kW, None, kA, kB, ks,
                                vW, None, vA, vB, vs, False)
    Qc, Kc, Vc = Q.detach().clone(), K.detach().clone(), V.detach().clone()
    torch.autograd.backward([Q, K, V], [dQ.clone(), dK.clone(), dV.clone()])
    save("fp16_lora_qkv", X=X, dQ=dQ, dK=dK, dV=dV, qW=qW, qA=qA, qB=qB, kW=kW, kA=kA, kB=kB,
         vW=vW, vA=vA, vB=vB, s=0.5, Q=Qc, K=Kc, V=Vc, dX=Xr.grad,
         d_qA=qA.grad, d_qB=qB.grad, d_kA=kA.grad, d_kB=kB.grad, d_vA=vA.grad, d_vB=vB.grad)
    gW, gA, gB, gs = lora(I, H, r, 2.0)
    uW, uA, uB, us = lora(I, H, r, 2.0)
    dW, dA, dB, ds = lora(H, I, r, 2.0)
    Xr = X.clone().requires_grad_()
    out = fl.LoRA_MLP.apply(Xr, gW, None, gA, gB, gs, uW, None, uA, uB, us, dW, None, dA, dB, ds,
                            k.swiglu.swiglu_fg_kernel, k.swiglu.swiglu_DWf_DW_dfg_kernel, False)
    out_c = out.detach().clone()
    out.backward(dY.clone())
    save("fp16_lora_mlp_swiglu", X=X, dY=dY, gW=gW, gA=gA, gB=gB, uW=uW, uA=uA, uB=uB,
         dW=dW, dA=dA, dB=dB, s=2.0, out=out_c, dX=Xr.grad,
         d_gA=gA.grad, d_gB=gB.grad, d_uA=uA.grad, d_uB=uB.grad, d_dA=dA.grad, d_dB=dB.grad)


def main():
    k = load_reference_kernels()
    gen_fp16_rounding_points(k)
    gen_fp16_lora(k)
    gen_rmsnorm(k)
    gen_rope(k)
    gen_ce(k)
    gen_glu(k)
    gen_lora(k)
    gen_reference_tests(k)


if __name__ == "__main__":
    main()
