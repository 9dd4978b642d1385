"""LoRA_MLP / LoRA_QKV / LoRA_W -- host-side mirror of unsloth/kernels/fast_lora.py.

Same classes, `apply_*` entry points, argument order and gradient slots as the reference
(fast_lora.py:28-229, 235-332, 335-571, 574-657).  The hand-derived backward is the same
algebra (fast_lora.py:42-62); what changes is how it is executed on a B200:

  * every projection is ONE multi-segment tcgen05 GEMM (csrc/gemm_tcgen05.cu): base weight and
    the rank-r update `(X A^T)(s B^T)` accumulate in the same fp32 TMEM accumulator -- the
    reference issues a cuBLAS GEMM plus an `addmm_` and rounds to bf16 in between;
  * projections that share an input (q/k/v, gate/up) share ONE skinny GEMM `X @ [A_q;A_k;A_v]^T`
    into a zero-padded 64-wide rank block;
  * dX of a group is ONE launch: sum_i dY_i @ W_i + G @ A_cat, with W_i consumed as an MN-major
    operand straight from the [out,in] dequantised buffer (no transpose, no running bf16 sum);
  * dA / dB are split-K reductions over tokens with both operands MN-major (X^T @ G, dY^T @ XA),
    fixed reduction order (deterministic), fp32 results (the reference returns bf16-precision
    grads, fast_lora.py:172-189).
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from . import utils as KU
from .utils import (GradModeAware, Problem, RANK_BLOCK, sink_lora_grads, fused_dequant_enabled, gemm_grouped, gemm_nf4, as_b_operand, cached_cast_pad, cast_pad, dense_weight, gemm,
                    gemm_glu, glu_fusable, fused_glu_enabled,
                    keep_dequant, keep_for_backward, get_lora_parameters, get_lora_parameters_bias,  # noqa: F401
                    matmul_lora)  # noqa: F401
from .swiglu import swiglu_fg_kernel, swiglu_DWf_DW_dfg_kernel
from .geglu import (geglu_exact_forward_kernel, geglu_exact_backward_kernel,
                    geglu_approx_forward_kernel, geglu_approx_backward_kernel)

_SM = 148

# the elementwise functions LoRA_MLP is called with (apply_lora_mlp_*) -> activation id of the GLU epilogue
_GLU_FWD_ACT = {swiglu_fg_kernel: L.ACT_SWIGLU, geglu_approx_forward_kernel: L.ACT_GEGLU_APPROX,
                geglu_exact_forward_kernel: L.ACT_GEGLU_EXACT}
_GLU_BWD_ACT = {swiglu_DWf_DW_dfg_kernel: L.ACT_SWIGLU, geglu_approx_backward_kernel: L.ACT_GEGLU_APPROX,
                geglu_exact_backward_kernel: L.ACT_GEGLU_EXACT}

def _grouped():
    """UB200_GROUPED=0 falls back to one launch per GEMM (round-1 schedule; A/B measurements)."""
    return os.environ.get("UB200_GROUPED", "1") != "0"


# Launch schedule per phase.  "auto" = what the interleaved A/B at cfg2 sizes measured as fastest
# (benchmarks/lora_group_bench.py, profiles/r2_lora_group_bench.log): the rank-block products are
# HBM-bound streams that already run at 0.65-1.0 of the copy bandwidth as separate launches, so
# grouping only removes launch gaps and fill/drain -- worth it where the streams are short (q/k/v/o
# backward: -8 %, MLP forward: -1 %), not where they are long (MLP backward: +4-6 % because the
# streaming tiles evict the dense GEMM's L2-resident operand).
_AUTO_FWD = {"qkv": False, "o": True, "mlp": True}
_AUTO_BWD = {"qkv": 1, "o": 1, "mlp": 0}


def _fwd_grouped(kind="mlp"):
    """Forward: XA = X @ A_cat^T produced inside the launch of the projections that consume it."""
    if not _grouped():
        return False
    v = os.environ.get("UB200_GROUPED_FWD", "auto")
    return _AUTO_FWD.get(kind, True) if v == "auto" else v != "0"


def _bwd_mode(kind="mlp"):
    """Backward schedule of a projection group:
      0  one launch per GEMM (round 1);
      1  everything in one persistent launch (rank-block tiles share the SMs with the dense tiles);
      2  TWO launches: all rank-block products of the phase (G, dB_i, dA) grouped, then the dense dX."""
    if not _grouped():
        return 0
    v = os.environ.get("UB200_GROUPED_BWD", "auto")
    return _AUTO_BWD.get(kind, 0) if v == "auto" else int(v)


def _launch_backward(front, dense, tail, mode):
    """front: rank-block producers / dB reductions; dense: the dX (or DW) GEMM; tail: dA (needs all of G)."""
    if mode == 1:
        probs = front + dense + tail
        if probs:
            gemm_grouped(probs)
        return
    if front or tail:
        gemm_grouped(front + tail)
    if dense:
        gemm_grouped(dense)


def _split_k_grouped(T):
    """Split factor of the token reductions (dA / dB) inside a grouped launch: ~32 k-blocks (2048
    tokens) per work item, so the items are short enough to fill the tail of the launch."""
    kb = (T + 63) // 64
    return max(1, min(16, kb // 32))


def _epoch():
    from . import utils as _u
    return _u.PARAM_EPOCH


def _as2d(t):
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1 or (t2.stride(0) % 8) or (t2.data_ptr() % 16):
        t2 = t2.contiguous()
    return t2


def _split_k(M, N, K):
    tiles = ((M + 127) // 128) * ((N + 63) // 64 if N <= 64 else (N + 127) // 128)
    kb = (K + 63) // 64
    return max(1, min(_SM // max(tiles, 1), kb // 4 if kb >= 4 else 1))


class _Group:
    """Projections (W, W_quant, A, B, s) that share one 2-D input X2 [T, in]."""

    def __init__(self, X2, projs, kind="mlp"):
        self.X2 = X2
        self.projs = projs
        self.kind = kind
        self.T, self.in_f = X2.shape
        self.dtype = X2.dtype
        self.dev = X2.device
        self.dense = None          # dequantised weights kept from forward to backward (keep_dequant)
        self._init_ranks()

    def drop_input(self):
        """Keep only what the backward needs from the parameters (not the activation)."""
        self.X2 = None
        return self

    def _init_ranks(self):
        projs = self.projs
        self.offs, off = [], 0
        for (_, _, A, _, _) in projs:
            self.offs.append(off)
            off += 0 if A is None else A.shape[0]
        self.rank_total = off
        self.has_lora = off > 0
        self.Rp = ((off + RANK_BLOCK - 1) // RANK_BLOCK) * RANK_BLOCK
        self._A_cat = None

    # [Rp, in] : rows offs[i].. hold A_i in the compute dtype, zero elsewhere
    def A_cat(self, refresh=False):
        if self._A_cat is None:
            blocks = [(o, A) for o, (_, _, A, _, _) in zip(self.offs, self.projs) if A is not None]
            if len(blocks) == 1 and blocks[0][0] == 0:
                A = blocks[0][1]
                A = A if A.stride(-1) == 1 else A.contiguous()
                self._A_cat = cached_cast_pad(A, (self.Rp, self.in_f), self.dtype, refresh=refresh)
                return self._A_cat
            # several adapters share the rank block: memoise on the first adapter's Parameter,
            # keyed by the versions of all of them
            first = blocks[0][1]
            ver = tuple((A._version, A.data_ptr()) for _, A in blocks) + (_epoch(),)
            cache = first.__dict__.setdefault("_ub200_acat_cache", {}) if isinstance(first, torch.nn.Parameter) else None
            key = (self.Rp, self.in_f, self.dtype, tuple(o for o, _ in blocks))
            plan = KU.ACTIVE_PLAN
            if plan is not None and plan.in_step:
                hit = plan.lookup(first, ("acat",) + key)     # rebuilt by plan.refresh() at the start of this step
                if hit is not None:
                    self._A_cat = hit
                    return self._A_cat
            if cache is not None and not refresh:
                hit = cache.get(key)
                if hit is not None and hit[0] == ver:
                    self._A_cat = hit[1]
                    return self._A_cat
            A_cat = torch.empty((self.Rp, self.in_f), dtype=self.dtype, device=self.dev)
            parts = []
            for j, (o, A) in enumerate(blocks):
                end = blocks[j + 1][0] if j + 1 < len(blocks) else self.Rp
                A = A if A.stride(-1) == 1 else A.contiguous()
                cast_pad(A, A_cat[o:end])
                parts.append((A, A_cat[o:end]))
            if cache is not None:
                cache[key] = (ver, A_cat)
            if plan is not None and plan.in_step:
                plan.record(first, ("acat",) + key, A_cat, parts)
            self._A_cat = A_cat
        return self._A_cat

    def forward(self, keep=False, glu_act=None):
        """Returns ([Y_i], XA) with XA = X @ A_cat^T  ([T, Rp], unscaled) or None.  `keep`: the
        dequantised weights are private tensors left in `self.dense` for the backward.
        `glu_act` (gate / up group of LoRA_MLP only): the LAST projection's launch applies the gated
        activation in its epilogue -- h = act(Y_0) * Y_1 is left in `self.h` (ub200_gemm_glu)."""
        self.h = None
        if fused_dequant_enabled() and self._fusable():
            return self._forward_fused_dequant()
        if glu_act is not None and not (len(self.projs) == 2 and self.dtype in (torch.bfloat16, torch.float16)
                                        and self.projs[0][0].shape[0] == self.projs[1][0].shape[0]):
            glu_act = None
        if glu_act is None and _fwd_grouped(self.kind):
            return self._forward_grouped(keep)
        X2, T, dt, dev = self.X2, self.T, self.dtype, self.dev
        if keep:
            self.dense = []
        XA = None
        if self.has_lora:
            XA = gemm(T, self.Rp, [(X2, self.A_cat(refresh=True), self.in_f)],
                      torch.empty((T, self.Rp), dtype=dt, device=dev))
        outs = []
        for off, (W, Wq, A, B, s) in zip(self.offs, self.projs):
            Wd = dense_weight(W, Wq, dt, 0, fresh=keep)
            if keep:
                self.dense.append(Wd if Wq is not None else None)
            Bop, b_mn = as_b_operand(Wd)
            N = Wd.shape[0]
            segs = [(X2, Bop, self.in_f)]
            if A is not None:
                Bc = B if B.stride(-1) == 1 else B.contiguous()
                if b_mn:
                    B_pad = cached_cast_pad(Bc, (self.Rp, N), dt, row_off=off, scale=s, transpose=True,
                                            refresh=True)
                else:
                    B_pad = cached_cast_pad(Bc, (N, self.Rp), dt, col_off=off, scale=s, refresh=True)
                segs.append((XA, B_pad, self.Rp))
            Y = torch.empty((T, N), dtype=dt, device=dev)
            if glu_act is not None and len(outs) == 1 and glu_fusable(N, Y, outs[0]):
                # up projection: g -> Y and h = act(e) * g in the same epilogue           (fast_lora.py:84-87)
                self.h = torch.empty((T, N), dtype=dt, device=dev)
                gemm_glu(L.GLU_EPI_FWD, glu_act, T, N, segs, self.h, outs[0], Y, a_mn=False, b_mn=b_mn)
            else:
                gemm(T, N, segs, Y, a_mn=False, b_mn=b_mn)
            outs.append(Y)
        return outs, XA

    # ---- NF4 expansion fused into the GEMM (csrc/gemm_nf4.cu; prototype behind UB200_FUSED_DEQUANT) --
    def _fusable(self):
        for (W, Wq, A, B, s) in self.projs:
            if Wq is None or type(Wq) is list or W.shape[0] == 1:
                return False
            if Wq.blocksize != 64 or Wq.state2.blocksize != 256 or Wq.dtype != self.dtype or self.in_f % 64:
                return False
        return self.dtype in (torch.bfloat16, torch.float16)

    def _forward_fused_dequant(self):
        """No 16-bit copy of W is made: nothing is kept for the backward (it re-dequantises)."""
        X2, T, dt, dev = self.X2, self.T, self.dtype, self.dev
        XA = None
        if self.has_lora:
            XA = gemm(T, self.Rp, [(X2, self.A_cat(refresh=True), self.in_f)],
                      torch.empty((T, self.Rp), dtype=dt, device=dev))
        outs = []
        for off, (W, Wq, A, B, s) in zip(self.offs, self.projs):
            N = Wq.shape[0]
            lora = None
            if A is not None:
                Bc = B if B.stride(-1) == 1 else B.contiguous()
                lora = (XA, cached_cast_pad(Bc, (N, self.Rp), dt, col_off=off, scale=s, refresh=True), self.Rp)
            outs.append(gemm_nf4(X2, W, Wq, torch.empty((T, N), dtype=dt, device=dev), lora))
        return outs, XA

    # ---- grouped execution: ONE persistent launch per phase (csrc/gemm_grouped.cu) ---------------
    def _forward_grouped(self, keep):
        X2, T, dt, dev = self.X2, self.T, self.dtype, self.dev
        if keep:
            self.dense = []
        probs, XA = [], None
        r_true = self.rank_total
        if self.has_lora:
            XA = torch.empty((T, self.Rp), dtype=dt, device=dev)
            probs.append(Problem(T, self.Rp, [(X2, self.A_cat(refresh=True), self.in_f)], XA,
                                 signals=True, tag="rank"))
        outs = []
        for off, (W, Wq, A, B, s) in zip(self.offs, self.projs):
            Wd = dense_weight(W, Wq, dt, len(outs), fresh=keep)
            if keep:
                self.dense.append(Wd if Wq is not None else None)
            Bop, b_mn = as_b_operand(Wd)
            N = Wd.shape[0]
            segs = [(X2, Bop, self.in_f)]
            wait = None
            if A is not None:
                Bc = B if B.stride(-1) == 1 else B.contiguous()
                if b_mn:
                    B_pad = cached_cast_pad(Bc, (self.Rp, N), dt, row_off=off, scale=s, transpose=True,
                                            refresh=True)
                else:
                    B_pad = cached_cast_pad(Bc, (N, self.Rp), dt, col_off=off, scale=s, refresh=True)
                segs.append((XA, B_pad, self.Rp, A.shape[0]))
                wait = (0, 1, False)
            Y = torch.empty((T, N), dtype=dt, device=dev)
            probs.append(Problem(T, N, segs, Y, b_mn=b_mn, wait=wait))
            outs.append(Y)
        gemm_grouped(probs)
        return outs, XA

    def backward_problems(self, dYs, XA, dX_out=None, need_dX=True, extra_front=()):
        """The problems of this group's backward:
          front  G = sum_i dY_i @ (s_i B_i)          [T, Rp]   rank block (signals)
                 dB_i = s_i dY_i^T @ XA              [out_i, Rp] fp32, split-K over tokens
                 (+ extra_front: rank-block problems of a neighbouring group riding along)
          dense  dX = sum_i dY_i @ W_i + G @ A_cat   waits for its row block of G (same launch only)
          tail   dA_cat^T = X^T @ G                  [in, Rp] fp32, split-K, waits for all of G
        Returns (front, dense, tail, finish) with finish() -> (dX or None, [(dA_i, dB_i)...])."""
        X2, T, dt, dev, Rp = self.X2, self.T, self.dtype, self.dev, self.Rp
        front, dense, tail = list(extra_front), [], []
        G = dA_catT = pG = None
        dB_fulls = []
        sk = _split_k_grouped(T)
        if self.has_lora:
            segs = []
            for off, dY, (W, Wq, A, B, s) in zip(self.offs, dYs, self.projs):
                if A is None:
                    continue
                Bc = B if B.stride(-1) == 1 else B.contiguous()
                out_f = Bc.shape[0]
                B_pad = cached_cast_pad(Bc, (out_f, Rp), dt, col_off=off, scale=s)
                segs.append((dY, B_pad, out_f))
            G = torch.empty((T, Rp), dtype=dt, device=dev)
            pG = Problem(T, Rp, segs, G, b_mn=True, signals=True, tag="rank")
            front.insert(0, pG)
            for dY, (W, Wq, A, B, s) in zip(dYs, self.projs):
                if A is None:
                    dB_fulls.append(None)
                    continue
                out_f = dY.shape[1]
                dB_full = torch.empty((out_f, Rp), dtype=torch.float32, device=dev)
                front.append(Problem(out_f, Rp, [(dY, XA, T)], dB_full, a_mn=True, b_mn=True, alpha=s,
                                     split_k=sk, tag="rank"))
                dB_fulls.append(dB_full)
        dX = None
        if need_dX:
            segs = []
            kept = self.dense if self.dense is not None else [None] * len(self.projs)
            for slot, (dY, (W, Wq, A, B, s)) in enumerate(zip(dYs, self.projs)):
                Wd = kept[slot] if kept[slot] is not None else dense_weight(W, Wq, dt, slot)   # [out, in]
                Bop, b_mn = as_b_operand(Wd.t())
                if not b_mn:
                    Bop, b_mn = Bop.t().contiguous(), True
                segs.append((dY, Bop, dY.shape[1]))
            wait = None
            if self.has_lora:
                segs.append((G, self.A_cat(), Rp, self.rank_total))
                wait = (pG, len(segs) - 1, False)
            # dA = X^T @ G may read X while dX tiles are being stored: dX is never written over the
            # saved X buffer when adapters are active (the reference's inplace=True is a memory
            # optimisation, fast_lora.py:194, 498; the 67 MB it saves is transient here)
            inplace_ok = dX_out is not None and not self.has_lora
            dX = dX_out if inplace_ok else torch.empty((T, self.in_f), dtype=dt, device=dev)
            dense.append(Problem(T, self.in_f, segs, dX, b_mn=True, wait=wait))
        if self.has_lora:
            dA_catT = torch.empty((self.in_f, Rp), dtype=torch.float32, device=dev)
            tail.append(Problem(self.in_f, Rp, [(X2, G, T)], dA_catT, a_mn=True, b_mn=True, split_k=sk,
                                wait=(pG, 0, True), tag="rank"))

        def finish():
            grads = []
            for off, dB_full, (W, Wq, A, B, s) in zip(self.offs, dB_fulls or [None] * len(self.projs), self.projs):
                if A is None:
                    grads.append((None, None))
                    continue
                r = A.shape[0]
                grads.append(sink_lora_grads(A, B, dA_catT[:, off:off + r].t(), dB_full[:, off:off + r]))
            return dX, grads
        return front, dense, tail, finish

    def backward(self, dYs, XA, dX_out=None, need_dX=True):
        """dYs: list of [T, out_i].  Returns (dX [T,in] or None, [(dA_i, dB_i) or (None, None)])."""
        mode = _bwd_mode(self.kind)
        if mode:
            front, dense, tail, finish = self.backward_problems(dYs, XA, dX_out, need_dX)
            _launch_backward(front, dense, tail, mode)
            return finish()
        X2, T, dt, dev, Rp = self.X2, self.T, self.dtype, self.dev, self.Rp
        grads = []
        G = None
        if self.has_lora:
            # dB_i [out, Rp] = s_i * dY_i^T @ XA ; keep this adapter's columns.  Independent of G.
            def _dBs():
                res = []
                for dY, (W, Wq, A, B, s) in zip(dYs, self.projs):
                    if A is None:
                        res.append(None)
                        continue
                    out_f = dY.shape[1]
                    dB_full = torch.empty((out_f, Rp), dtype=torch.float32, device=dev)
                    gemm(out_f, Rp, [(dY, XA, T)], dB_full, a_mn=True, b_mn=True, alpha=s,
                         split_k=_split_k(out_f, Rp, T))
                    res.append(dB_full)
                return res
            dB_fulls = None
            # G[T, Rp] = sum_i dY_i @ (s_i B_i) placed at the adapter's rank slot.  The B operand
            # is the SAME zero-padded [out_i, Rp] block the forward used, consumed MN-major
            # ([K=out_i, N=Rp] row-major): no transposed copy of B is ever made.
            segs = []
            for off, dY, (W, Wq, A, B, s) in zip(self.offs, dYs, self.projs):
                if A is None:
                    continue
                Bc = B if B.stride(-1) == 1 else B.contiguous()
                out_f = Bc.shape[0]
                B_pad = cached_cast_pad(Bc, (out_f, Rp), dt, col_off=off, scale=s)
                segs.append((dY, B_pad, out_f))
            G = gemm(T, Rp, segs, torch.empty((T, Rp), dtype=dt, device=dev), b_mn=True)
            # dA_cat^T [in, Rp] = X^T @ G   (both operands MN-major, reduction over tokens)
            dA_catT = torch.empty((self.in_f, Rp), dtype=torch.float32, device=dev)
            gemm(self.in_f, Rp, [(X2, G, T)], dA_catT, a_mn=True, b_mn=True,
                 split_k=_split_k(self.in_f, Rp, T))
            if dB_fulls is None:
                dB_fulls = _dBs()
            for off, dY, dB_full, (W, Wq, A, B, s) in zip(self.offs, dYs, dB_fulls, self.projs):
                if A is None:
                    grads.append((None, None))
                    continue
                r = A.shape[0]
                grads.append(sink_lora_grads(A, B, dA_catT[:, off:off + r].t(), dB_full[:, off:off + r]))
        else:
            grads = [(None, None)] * len(self.projs)
        if not need_dX:
            return None, grads
        # dX = sum_i dY_i @ W_i  +  G @ A_cat      (one launch; W_i as MN-major operands)
        segs = []
        kept = self.dense if self.dense is not None else [None] * len(self.projs)
        for slot, (dY, (W, Wq, A, B, s)) in enumerate(zip(dYs, self.projs)):
            Wd = kept[slot] if kept[slot] is not None else dense_weight(W, Wq, dt, slot)   # [out, in]
            # as the B operand of dY @ Wd we need [N=in, K=out]: that is Wd^T
            Bop, b_mn = as_b_operand(Wd.t())
            if not b_mn:
                # Wd^T has contiguous rows only if Wd was itself a transposed view: fall back to
                # a materialised [in, out] copy so all segments share the MN-major layout
                Bop, b_mn = Bop.t().contiguous(), True
            segs.append((dY, Bop, dY.shape[1]))
        if self.has_lora:
            segs.append((G, self.A_cat(), Rp))
        dX = dX_out if dX_out is not None else torch.empty((T, self.in_f), dtype=dt, device=dev)
        gemm(T, self.in_f, segs, dX, a_mn=False, b_mn=True)
        return dX, grads


class LoRA_MLP(GradModeAware, torch.autograd.Function):
    """fast_lora.py:28-229.  i = down(act(gate(X)) * up(X)), each projection NF4 + LoRA."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, X, gateW, gateW_quant, gateA, gateB, gateS, upW, upW_quant, upA, upB, upS,
                downW, downW_quant, downA, downB, downS, _forward_function, _backward_function,
                inplace=True):
        L.require_cuda(X)
        shape = X.shape
        X2 = _as2d(X)
        grp = _Group(X2, [(gateW, gateW_quant, gateA, gateB, gateS), (upW, upW_quant, upA, upB, upS)])
        keep = keep_for_backward(ctx.needs_input_grad)
        glu_act = _GLU_FWD_ACT.get(_forward_function) if fused_glu_enabled("fwd") else None
        (e, g), XA1 = grp.forward(keep, glu_act=glu_act)
        b_s = shape[:-1]
        if grp.h is not None:
            h2 = grp.h                               # produced by the up projection's epilogue
            grp.h = None
        else:
            h = _forward_function(e.view(*b_s, -1) if X.dim() == 3 else e.view(1, *e.shape),
                                  g.view(*b_s, -1) if X.dim() == 3 else g.view(1, *g.shape))
            h2 = h.reshape(-1, h.shape[-1])
        grp2 = _Group(h2, [(downW, downW_quant, downA, downB, downS)])
        (i,), XA2 = grp2.forward(keep)
        ctx.dense = (grp.dense, grp2.dense)
        ctx.custom_saved_tensors = (gateW, gateW_quant, gateS, upW, upW_quant, upS, downW,
                                    downW_quant, downS, _backward_function)
        # LoRA A/B are kept as the caller's Parameter objects (not through save_for_backward) so
        # that the per-step cast cache, which lives on the Parameter, is hit in backward as well
        ctx.lora = (gateA, gateB, upA, upB, downA, downB)
        ctx.save_for_backward(X2, e, g, XA1, XA2)
        ctx.inplace = inplace
        ctx.shape = shape
        return i.view(*b_s, -1)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dY):
        (gateW, gateW_quant, gateS, upW, upW_quant, upS, downW, downW_quant, downS,
         _backward_function) = ctx.custom_saved_tensors
        gateA, gateB, upA, upB, downA, downB = ctx.lora
        X2, e, g, XA1, XA2 = ctx.saved_tensors
        dY2 = _as2d(dY)
        T = X2.shape[0]
        dt, dev = X2.dtype, X2.device
        if _bwd_mode("mlp"):
            return LoRA_MLP._backward_grouped(ctx, dY2, X2, e, g, XA1, XA2)
        # --- down projection: DW = dY @ W_down + (dY @ s B_down) @ A_down          (:155)
        # `h` is not needed yet: a group over a placeholder input gives DW and G_down
        down = _Group(e, [(downW, downW_quant, downA, downB, downS)])  # in_f = I (e is [T, I])
        G_down = None
        segs = []
        dense_gu, dense_down = ctx.dense
        ctx.dense = None
        Wd = dense_down[0] if dense_down is not None and dense_down[0] is not None else \
            dense_weight(downW, downW_quant, dt, 0)                    # [H, I]
        dense_down = None
        Bop, b_mn = as_b_operand(Wd.t())
        if not b_mn:
            Bop, b_mn = Bop.t().contiguous(), True
        segs.append((dY2, Bop, dY2.shape[1]))
        if downA is not None:
            Bc = downB if downB.stride(-1) == 1 else downB.contiguous()
            B_pad = cached_cast_pad(Bc, (Bc.shape[0], down.Rp), dt, scale=downS)     # as in forward
            G_down = gemm(T, down.Rp, [(dY2, B_pad, Bc.shape[0])],
                          torch.empty((T, down.Rp), dtype=dt, device=dev), b_mn=True)
            segs.append((G_down, down.A_cat(), down.Rp))
        DW = torch.empty((T, e.shape[1]), dtype=dt, device=dev)
        glu_act = _GLU_BWD_ACT.get(_backward_function) if fused_glu_enabled("bwd") else None
        if glu_act is not None and glu_fusable(e.shape[1], DW, e, g) and e.stride(0) == g.stride(0):
            # DW never reaches HBM: the activation backward runs on the accumulator tile, in place over e / g
            gemm_glu(L.GLU_EPI_BWD, glu_act, T, e.shape[1], segs, DW, e, g, a_mn=False, b_mn=True)
            h, df, de = DW, e, g
        else:
            gemm(T, e.shape[1], segs, DW, a_mn=False, b_mn=True)
            # --- activation backward, in place: DW <- h, e <- df, g <- de            (:156-157)
            h, df, de = _backward_function(DW, e, g)
        # --- down LoRA grads                                                      (:171-172)
        d_downA = d_downB = None
        if downA is not None:
            r = downA.shape[0]
            I = h.shape[1]
            dA_T = torch.empty((I, down.Rp), dtype=torch.float32, device=dev)
            gemm(I, down.Rp, [(h, G_down, T)], dA_T, a_mn=True, b_mn=True,
                 split_k=_split_k(I, down.Rp, T))
            Hout = dY2.shape[1]
            dB_full = torch.empty((Hout, down.Rp), dtype=torch.float32, device=dev)
            gemm(Hout, down.Rp, [(dY2, XA2, T)], dB_full, a_mn=True, b_mn=True, alpha=downS,
                 split_k=_split_k(Hout, down.Rp, T))
            d_downA, d_downB = sink_lora_grads(downA, downB, dA_T[:, :r].t(), dB_full[:, :r])
        # --- gate / up: LoRA grads and dX (into the saved X buffer when inplace)   (:178-204)
        grp = _Group(X2, [(gateW, gateW_quant, gateA, gateB, gateS), (upW, upW_quant, upA, upB, upS)])
        grp.dense = dense_gu
        dX, ((d_gateA, d_gateB), (d_upA, d_upB)) = grp.backward(
            [de, df], XA1, dX_out=X2 if ctx.inplace else None)
        return (dX.view(ctx.shape), None, None, d_gateA, d_gateB, None, None, None, d_upA, d_upB,
                None, None, None, d_downA, d_downB, None, None, None, None)


def _mlp_backward_grouped(ctx, dY2, X2, e, g, XA1, XA2):
    """LoRA_MLP.backward (fast_lora.py:116-229) as TWO persistent launches around the in-place
    activation backward:
      A: G_down = dY @ sB_down | DW = dY @ W_down + G_down @ A_down | dB_down = s dY^T @ XA2
      (swiglu / geglu backward in place: DW <- h, e <- df, g <- de)
      B: G = de @ sB_gate + df @ sB_up | dB_gate, dB_up | dA_down^T = h^T @ G_down |
         dX = de @ W_gate + df @ W_up + G @ A_cat | dA_cat^T = X^T @ G"""
    (gateW, gateW_quant, gateS, upW, upW_quant, upS, downW, downW_quant, downS,
     _backward_function) = ctx.custom_saved_tensors
    gateA, gateB, upA, upB, downA, downB = ctx.lora
    T = X2.shape[0]
    dt, dev = X2.dtype, X2.device
    I, Hout = e.shape[1], dY2.shape[1]
    down = _Group(e, [(downW, downW_quant, downA, downB, downS)])
    dense_gu, dense_down = ctx.dense
    ctx.dense = None
    Wd = dense_down[0] if dense_down is not None and dense_down[0] is not None else \
        dense_weight(downW, downW_quant, dt, 0)                    # [H, I]
    dense_down = None
    Bop, b_mn = as_b_operand(Wd.t())
    if not b_mn:
        Bop, b_mn = Bop.t().contiguous(), True
    sk = _split_k_grouped(T)
    front, G_down, dB_full, pG = [], None, None, None
    segs, wait = [(dY2, Bop, Hout)], None
    if downA is not None:
        Bc = downB if downB.stride(-1) == 1 else downB.contiguous()
        B_pad = cached_cast_pad(Bc, (Bc.shape[0], down.Rp), dt, scale=downS)     # as in forward
        G_down = torch.empty((T, down.Rp), dtype=dt, device=dev)
        pG = Problem(T, down.Rp, [(dY2, B_pad, Bc.shape[0])], G_down, b_mn=True, signals=True, tag="rank")
        dB_full = torch.empty((Hout, down.Rp), dtype=torch.float32, device=dev)
        front = [pG, Problem(Hout, down.Rp, [(dY2, XA2, T)], dB_full, a_mn=True, b_mn=True, alpha=downS,
                             split_k=sk, tag="rank")]
        segs.append((G_down, down.A_cat(), down.Rp, downA.shape[0]))
        wait = (pG, 1, False)
    DW = torch.empty((T, I), dtype=dt, device=dev)
    mode = _bwd_mode("mlp")
    _launch_backward(front, [Problem(T, I, segs, DW, b_mn=True, wait=wait)], [], mode)
    h, df, de = _backward_function(DW, e, g)                       # in place            (:156-157)
    extra, dA_T = [], None
    if downA is not None:
        dA_T = torch.empty((I, down.Rp), dtype=torch.float32, device=dev)
        extra = [Problem(I, down.Rp, [(h, G_down, T)], dA_T, a_mn=True, b_mn=True, split_k=sk, tag="rank")]
    grp = _Group(X2, [(gateW, gateW_quant, gateA, gateB, gateS), (upW, upW_quant, upA, upB, upS)])
    grp.dense = dense_gu
    front, dense, tail, finish = grp.backward_problems([de, df], XA1, dX_out=X2 if ctx.inplace else None,
                                                       extra_front=extra)
    _launch_backward(front, dense, tail, mode)
    dX, ((d_gateA, d_gateB), (d_upA, d_upB)) = finish()
    d_downA = d_downB = None
    if downA is not None:
        r = downA.shape[0]
        d_downA, d_downB = sink_lora_grads(downA, downB, dA_T[:, :r].t(), dB_full[:, :r])
    return (dX.view(ctx.shape), None, None, d_gateA, d_gateB, None, None, None, d_upA, d_upB,
            None, None, None, d_downA, d_downB, None, None, None, None)


LoRA_MLP._backward_grouped = staticmethod(_mlp_backward_grouped)


def apply_lora_mlp_swiglu(self, X, inplace=True):
    """fast_lora.py:235-265."""
    gateW, gateW_quant, gateA, gateB, gateS = get_lora_parameters(self.gate_proj)
    upW, upW_quant, upA, upB, upS = get_lora_parameters(self.up_proj)
    downW, downW_quant, downA, downB, downS = get_lora_parameters(self.down_proj)
    return LoRA_MLP.apply(X, gateW, gateW_quant, gateA, gateB, gateS, upW, upW_quant, upA, upB,
                          upS, downW, downW_quant, downA, downB, downS, swiglu_fg_kernel,
                          swiglu_DWf_DW_dfg_kernel, inplace)


def apply_lora_mlp_geglu_exact(self, X, inplace=True):
    """fast_lora.py:271-301."""
    gateW, gateW_quant, gateA, gateB, gateS = get_lora_parameters(self.gate_proj)
    upW, upW_quant, upA, upB, upS = get_lora_parameters(self.up_proj)
    downW, downW_quant, downA, downB, downS = get_lora_parameters(self.down_proj)
    return LoRA_MLP.apply(X, gateW, gateW_quant, gateA, gateB, gateS, upW, upW_quant, upA, upB,
                          upS, downW, downW_quant, downA, downB, downS,
                          geglu_exact_forward_kernel, geglu_exact_backward_kernel, inplace)


def apply_lora_mlp_geglu_approx(self, X):
    """fast_lora.py:307-332 (no `inplace` argument, as in the reference)."""
    gateW, gateW_quant, gateA, gateB, gateS = get_lora_parameters(self.gate_proj)
    upW, upW_quant, upA, upB, upS = get_lora_parameters(self.up_proj)
    downW, downW_quant, downA, downB, downS = get_lora_parameters(self.down_proj)
    return LoRA_MLP.apply(X, gateW, gateW_quant, gateA, gateB, gateS, upW, upW_quant, upA, upB,
                          upS, downW, downW_quant, downA, downB, downS,
                          geglu_approx_forward_kernel, geglu_approx_backward_kernel)


class LoRA_QKV(GradModeAware, torch.autograd.Function):
    """fast_lora.py:335-540."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, X, QW, QW_quant, QA, QB, QS, KW, KW_quant, KA, KB, KS, VW, VW_quant, VA, VB,
                VS, inplace=True):
        L.require_cuda(X)
        shape = X.shape
        X2 = _as2d(X)
        grp = _Group(X2, [(QW, QW_quant, QA, QB, QS), (KW, KW_quant, KA, KB, KS),
                          (VW, VW_quant, VA, VB, VS)], kind="qkv")
        (Q, K, V), XA = grp.forward(keep_for_backward(ctx.needs_input_grad))
        ctx.dense = grp.dense
        if len(shape) == 3:
            Q, K, V = (t.view(shape[0], shape[1], -1) for t in (Q, K, V))
        ctx.custom_saved_tensors = (QW, QW_quant, QS, KW, KW_quant, KS, VW, VW_quant, VS)
        ctx.lora = (QA, QB, KA, KB, VA, VB)
        ctx.save_for_backward(X2, XA)
        ctx.inplace = inplace
        ctx.shape = shape
        return Q, K, V

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dQ, dK, dV):
        QW, QW_quant, QS, KW, KW_quant, KS, VW, VW_quant, VS = ctx.custom_saved_tensors
        QA, QB, KA, KB, VA, VB = ctx.lora
        X2, XA = ctx.saved_tensors
        grp = _Group(X2, [(QW, QW_quant, QA, QB, QS), (KW, KW_quant, KA, KB, KS),
                          (VW, VW_quant, VA, VB, VS)], kind="qkv")
        grp.dense, ctx.dense = ctx.dense, None
        need_dX = ctx.needs_input_grad[0]     # False for the first layer (embedding output)
        dX, ((dQA, dQB), (dKA, dKB), (dVA, dVB)) = grp.backward(
            [_as2d(dQ), _as2d(dK), _as2d(dV)], XA, dX_out=X2 if ctx.inplace else None, need_dX=need_dX)
        return (dX.view(ctx.shape) if need_dX else None, None, None, dQA, dQB, None, None, None, dKA, dKB,
                None, None, None, dVA, dVB, None, None)


def apply_lora_qkv(self, X, inplace=True):
    """fast_lora.py:543-571."""
    QW, QW_quant, QA, QB, QS = get_lora_parameters(self.q_proj)
    KW, KW_quant, KA, KB, KS = get_lora_parameters(self.k_proj)
    VW, VW_quant, VA, VB, VS = get_lora_parameters(self.v_proj)
    return LoRA_QKV.apply(X, QW, QW_quant, QA, QB, QS, KW, KW_quant, KA, KB, KS, VW, VW_quant,
                          VA, VB, VS, inplace)


class LoRA_W(GradModeAware, torch.autograd.Function):
    """fast_lora.py:574-650 (o_proj)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, X, W, W_quant, A, B, S):
        L.require_cuda(X)
        shape = X.shape
        X2 = _as2d(X)
        grp = _Group(X2, [(W, W_quant, A, B, S)], kind="o")
        (XW,), XA = grp.forward(keep_for_backward(ctx.needs_input_grad))
        ctx.dense = grp.dense
        ctx.custom_saved_tensors = (W, W_quant, S)
        ctx.lora = (A, B)
        ctx.save_for_backward(X2, XA)
        ctx.shape = shape
        return XW.view(*shape[:-1], -1)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dY):
        W, W_quant, S = ctx.custom_saved_tensors
        A, B = ctx.lora
        X2, XA = ctx.saved_tensors
        grp = _Group(X2, [(W, W_quant, A, B, S)], kind="o")
        grp.dense, ctx.dense = ctx.dense, None
        dX, ((dA, dB),) = grp.backward([_as2d(dY)], XA)
        return dX.view(ctx.shape), None, None, dA, dB, None


def apply_lora_o(self, X):
    """fast_lora.py:653-657."""
    OW, OW_quant, OA, OB, OS = get_lora_parameters(self.o_proj)
    return LoRA_W.apply(X, OW, OW_quant, OA, OB, OS)
