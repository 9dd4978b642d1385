"""Host-side mirror of unsloth/kernels/utils.py for the hot path:
`fast_dequantize` (:567-679), `matmul_lora` (:1128-1170), `get_lora_parameters[_bias]`
(:335-440), `QUANT_STATE`, plus the thin `gemm` wrapper over ub200_gemm; and the decode-time
consumers of the same NF4 format, `fast_gemv` (:874-973) and `fast_linear_forward` (:1082-1125).
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L

MAX_FUSED_SIZE = 65536
RANK_BLOCK = 64  # the GEMM's K block: LoRA rank blocks are zero-padded to a multiple of it


def QUANT_STATE(W):
    return getattr(W, "quant_state", None)


# ---------------------------------------------------------------------------------------------
# per-device reusable dequantised-weight buffers (the reference's WEIGHT_BUFFERS,
# kernels/utils.py:613-637, generalised to several slots because one multi-segment GEMM
# can consume several dequantised weights at once).  Stream ordering makes reuse safe.
# ---------------------------------------------------------------------------------------------
_WEIGHT_BUFFERS = {}


def _weight_buffer(device, slot, numel, dtype):
    key = (device.index, slot, dtype)
    buf = _WEIGHT_BUFFERS.get(key)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(numel, dtype=dtype, device=device, requires_grad=False)
        _WEIGHT_BUFFERS[key] = buf
    return buf[:numel]


def _unpack_quant_state(quant_state):
    if type(quant_state) is not list:
        s2 = quant_state.state2
        return (quant_state.absmax, quant_state.shape, quant_state.dtype, quant_state.blocksize,
                quant_state.offset, s2.absmax, s2.code, s2.blocksize)
    # old list form (kernels/utils.py:594-598)
    absmax, shape, dtype, blocksize, compressed_stats, _, _ = quant_state
    offset, state2 = compressed_stats
    absmax2, code2, blocksize2, _, _, _, _ = state2
    return absmax, shape, dtype, blocksize, offset, absmax2, code2, blocksize2


@torch.inference_mode
def fast_dequantize(W, quant_state=None, out=None, use_global_buffer=False, _slot=0):
    """NF4 double-quant -> quant_state.dtype.  Same contract as the reference
    (kernels/utils.py:567-679): `quant_state is None` returns W unchanged; `out=` must match
    shape/dtype; returns `out.t()` iff `W.shape[0] == 1` (packed weight passed transposed)."""
    if quant_state is None:
        return W
    L.require_cuda(W)
    absmax, shape, dtype, blocksize, offset, absmax2, code2, blocksize2 = _unpack_quant_state(quant_state)
    device = W.device
    numel = shape[0] * shape[1]
    if use_global_buffer:
        out = _weight_buffer(device, _slot, numel, dtype).view(shape)
    elif out is None:
        out = torch.empty(shape, dtype=dtype, device=device, requires_grad=False)
    else:
        assert out.shape == shape
        assert out.dtype == dtype
    if not torch.is_tensor(offset):
        offset = torch.tensor(float(offset), dtype=torch.float32, device=device)
    L.call("ub200_dequantize_nf4", L.ptr(W), L.ptr(absmax), L.ptr(code2), L.ptr(absmax2),
           L.ptr(offset), L.ptr(out), numel, int(blocksize), int(blocksize2), L.dt(dtype), L.stream())
    is_transposed = True if W.shape[0] == 1 else False
    return out.t() if is_transposed else out


# ---------------------------------------------------------------------------------------------
# decode-time GEMV
# ---------------------------------------------------------------------------------------------
_NF4_CODE = torch.tensor([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453,
                          -0.28444138169288635, -0.18477343022823334, -0.09105003625154495, 0.0,
                          0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
                          0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=torch.float32)


def _gemv_nf4(x, W, quant_state, out, lora_B=None, lora_t=None, s=0.0):
    absmax, shape, dtype, blocksize, offset, absmax2, code2, blocksize2 = _unpack_quant_state(quant_state)
    code16 = quant_state.code if type(quant_state) is not list else quant_state[6]
    if not torch.is_tensor(offset):
        offset = torch.tensor(float(offset), dtype=torch.float32, device=W.device)
    if code16 is not None and (code16.dtype != torch.float32 or code16.numel() != 16):
        raise RuntimeError("unsloth_b200: fast_gemv needs the 16-entry fp32 NF4 code")
    if code16 is not None and type(quant_state) is not list:
        # the standard NF4 table is a constant inside the kernel: passing no pointer takes one global
        # round trip out of the prologue of a ~10 us launch.  Checked once per quant_state (one D2H).
        std = getattr(quant_state, "_ub_std_nf4", None)
        if std is None or std[0] != code16.data_ptr():
            std = (code16.data_ptr(), bool(torch.equal(code16.detach().float().cpu(), _NF4_CODE)))
            try:
                quant_state._ub_std_nf4 = std
            except AttributeError:
                pass
        if std[1]:
            code16 = None
    if x.dtype != dtype or out.dtype != dtype:
        raise RuntimeError("unsloth_b200: fast_gemv needs X and out in quant_state.dtype")
    r = 0 if lora_B is None else lora_B.shape[1]
    L.call("ub200_gemv_nf4", L.ptr(x), L.ptr(W), None, L.ptr(absmax), L.ptr(code2), L.ptr(absmax2),
           L.ptr(offset), None if code16 is None else L.ptr(code16), L.ptr(out), int(shape[0]),
           int(shape[1]), int(blocksize), int(blocksize2),
           None if lora_B is None else L.ptr(lora_B), 0 if lora_B is None else lora_B.stride(0),
           None if lora_t is None else L.ptr(lora_t), r, float(s), L.dt(dtype), L.stream())
    return out


@torch.inference_mode
def fast_gemv(X, W, quant_state, out=None):
    """kernels/utils.py:874-973: X [1, 1, in] against an NF4 weight -> [1, 1, out], the packed
    weight expanded in registers.  ONE launch (the reference: a blockwise fp32 absmax launch, a
    torch `+= offset`, then bitsandbytes' 4-bit GEMV)."""
    if quant_state is None:
        return torch.matmul(X, W, out=out)
    L.require_cuda(X, W)
    _, shape, dtype, *_ = _unpack_quant_state(quant_state)
    if X.numel() != shape[1]:
        raise RuntimeError("unsloth_b200: fast_gemv is the bsz == 1, q_len == 1 path")
    if out is None:
        out = torch.empty((1, 1, shape[0]), dtype=dtype, device=W.device)
    _gemv_nf4(X.reshape(-1), W, quant_state, out)
    return out


def _fast_lora_cast(A, B, dtype):
    """`lora_A._fast_lora` / `lora_B._fast_lora` (kernels/utils.py:1103-1105): adapters cast once to
    the activation dtype for decoding (refreshed when the parameter is updated in place)."""
    for p_ in (A, B):
        c = getattr(p_, "_fast_lora", None)
        ver = (p_._version, PARAM_EPOCH, p_.data_ptr())
        if c is None or c.dtype != dtype or getattr(p_, "_fast_lora_version", None) != ver:
            p_._fast_lora = p_.detach().to(dtype).contiguous()
            p_._fast_lora_version = ver
    return A._fast_lora, B._fast_lora


@torch.inference_mode
def fast_linear_forward(proj, X, temp_lora=None, out=None):
    """kernels/utils.py:1082-1125: the decode-time projection.  q_len != 1 falls through to
    matmul_lora; bsz == q_len == 1 on an NF4 weight is ONE GEMV launch with the LoRA term in its
    epilogue (+ one 16-row GEMV for the LoRA temp A x); other shapes dequantise and use the GEMM."""
    W, W_quant, lora_A, lora_B, lora_S, bias = get_lora_parameters_bias(proj)
    bsz, q_len, in_dim = X.shape
    if q_len != 1:
        return matmul_lora(X, W, W_quant, lora_A, lora_B, lora_S)
    L.require_cuda(X, W)
    dtype = X.dtype
    single = bsz == 1 and dtype in (torch.bfloat16, torch.float16) and in_dim % 32 == 0
    if single and (W_quant is not None or W.dtype == dtype):
        x = X.reshape(-1)
        if not x.is_contiguous():
            x = x.contiguous()
        out_dim = _unpack_quant_state(W_quant)[1][0] if W_quant is not None else W.shape[0]
        if out is None:
            out = torch.empty((1, 1, out_dim), dtype=dtype, device=X.device)
        Bc = t = None
        if lora_A is not None:
            Ac, Bc = _fast_lora_cast(lora_A, lora_B, dtype)
            t = temp_lora if temp_lora is not None and temp_lora.dtype == torch.float32 else \
                torch.empty(Ac.shape[0], dtype=torch.float32, device=X.device)
            L.call("ub200_gemv_dense", L.ptr(x), L.ptr(Ac), Ac.stride(0), L.ptr(t), Ac.shape[0],
                   in_dim, L.dt(dtype), L.F32, L.stream())
        if W_quant is not None:
            _gemv_nf4(x, W, W_quant, out, Bc, t, lora_S or 0.0)
        else:
            Wc = W if W.stride(-1) == 1 else W.contiguous()
            L.call("ub200_gemv_dense", L.ptr(x), L.ptr(Wc), Wc.stride(0), L.ptr(out), out_dim, in_dim,
                   L.dt(dtype), L.dt(dtype), L.stream())
            if lora_A is not None:
                out.view(-1).addmv_(Bc, t.to(dtype), alpha=lora_S)
        if bias is not None:
            out += bias
        return out
    # bsz > 1 (or an odd shape): same arithmetic as the training primitive on [bsz, in] rows
    out2 = matmul_lora(X.reshape(bsz, in_dim), W, W_quant, lora_A, lora_B, lora_S)
    out2 = out2.view(bsz, 1, -1)
    if bias is not None:
        out2 = out2 + bias
    if out is not None:
        out.copy_(out2)
        return out
    return out2


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def _check_operand(t):
    if t.stride(-1) != 1 or t.dim() != 2:
        raise RuntimeError("unsloth_b200.gemm: operands must be 2-D with a contiguous last dim")


# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream
GEMM_EVENTS = None


def gemm(M, N, segs, out, a_mn=False, b_mn=False, alpha=1.0, accumulate=False, split_k=1,
         block_n=0, cta_group=0):
    """out[M,N] (+)= alpha * sum_s A_s . B_s^T on the tcgen05 tensor cores.

    segs: list of (A, B, K).  a_mn=False: A is [M, K]; True: A is [K, M] (row-major).
    b_mn=False: B is [N, K]; True: B is [K, N]."""
    n = len(segs)
    arr = (L.GemmSegment * n)()
    ab_dtype = segs[0][0].dtype
    for i, sg in enumerate(segs):
        A, B, K = sg[0], sg[1], sg[2]
        _check_operand(A); _check_operand(B)
        if A.dtype != ab_dtype or B.dtype != ab_dtype:
            raise RuntimeError("unsloth_b200.gemm: mixed operand dtypes %s/%s" % (A.dtype, B.dtype))
        arr[i].a = A.data_ptr(); arr[i].lda = A.stride(0)
        arr[i].b = B.data_ptr(); arr[i].ldb = B.stride(0)
        arr[i].k = K
    ws = None
    if split_k > 1:
        ws = torch.empty(split_k * M * N, dtype=torch.float32, device=out.device)
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.call("ub200_gemm", M, N, arr, n, int(a_mn), int(b_mn), L.dt(ab_dtype), L.ptr(out),
           out.stride(0), L.dt(out), float(alpha), int(accumulate), int(split_k), L.ptr(ws),
           int(block_n), int(cta_group), L.stream())
    if ev is not None:
        e1.record()
        bn = block_n or (256 if N > 128 else (128 if N > 64 else 64))
        pair = cta_group == 2 or (cta_group == 0 and bn >= 128 and M > 128)
        ev.append((2.0 * M * N * sum(sg[3] if len(sg) > 3 else sg[2] for sg in segs), e0, e1,
                   {"kernel": "gemm2" if pair else "gemm1"}))
    return out


def fused_glu_enabled(direction) -> bool:
    """UB200_FUSED_GLU: "1" (default) = the gated activation of LoRA_MLP runs in the epilogue of the
    projection that feeds it (ub200_gemm_glu); "0" = two launches (round-1 schedule); "fwd" / "bwd" = one
    direction only (A/B measurements)."""
    v = os.environ.get("UB200_FUSED_GLU", "1")
    return v == "1" or v == direction


def glu_fusable(N, *tensors) -> bool:
    """Conditions of ub200_gemm_glu: 16-bit row-major [T, N] tensors with 32-byte aligned rows, whole tiles."""
    if N % (256 if N > 128 else (128 if N > 64 else 64)):
        return False
    for t in tensors:
        if t.dtype not in (torch.bfloat16, torch.float16) or t.dim() != 2 or t.stride(1) != 1 \
                or t.stride(0) % 16 or t.data_ptr() % 32:
            return False
    return True


def gemm_glu(mode, act, M, N, segs, out, e, g, a_mn=False, b_mn=False, alpha=1.0, block_n=0, cta_group=0):
    """ub200_gemm_glu: `gemm` whose epilogue applies the gated activation to the accumulator tile.
    mode GLU_EPI_FWD: tile = up projection -> g, out = act(e) * g.  mode GLU_EPI_BWD: tile = DW;
    out <- h, e <- df, g <- de (in place, the reference's buffer reuse)."""
    n = len(segs)
    arr = (L.GemmSegment * n)()
    ab_dtype = segs[0][0].dtype
    for i, sg in enumerate(segs):
        A, B, K = sg[0], sg[1], sg[2]
        _check_operand(A); _check_operand(B)
        if A.dtype != ab_dtype or B.dtype != ab_dtype:
            raise RuntimeError("unsloth_b200.gemm_glu: mixed operand dtypes %s/%s" % (A.dtype, B.dtype))
        arr[i].a = A.data_ptr(); arr[i].lda = A.stride(0)
        arr[i].b = B.data_ptr(); arr[i].ldb = B.stride(0)
        arr[i].k = K
    if not (out.dtype == e.dtype == g.dtype == ab_dtype) or e.stride(0) != g.stride(0):
        raise RuntimeError("unsloth_b200.gemm_glu: C / e / g must share the operand dtype and e / g one row stride")
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.call("ub200_gemm_glu", int(mode), int(act), M, N, arr, n, int(a_mn), int(b_mn), L.dt(ab_dtype),
           L.ptr(out), out.stride(0), L.ptr(e), L.ptr(g), e.stride(0), float(alpha), int(block_n),
           int(cta_group), L.stream())
    if ev is not None:
        e1.record()
        bn = block_n or (256 if N > 128 else (128 if N > 64 else 64))
        pair = cta_group == 2 or (cta_group == 0 and bn >= 128 and M > 128)
        ev.append((2.0 * M * N * sum(sg[3] if len(sg) > 3 else sg[2] for sg in segs), e0, e1,
                   {"kernel": "gemm2_glu" if pair else "gemm1_glu"}))
    return out


# ---------------------------------------------------------------------------------------------
# grouped GEMM: one persistent launch for a list of dependent problems (csrc/gemm_grouped.cu)
# ---------------------------------------------------------------------------------------------
class Problem:
    """One problem of a grouped launch: out[M,N] (+)= alpha * sum_s A_s . B_s^T.
    segs: (A, B, K) or (A, B, K, K_true) -- K_true is the algorithmic reduction length (the LoRA
    rank for a zero-padded rank block), used only for flop accounting.
    wait: None or (earlier problem whose output this one reads -- its index in the launch list or the
    Problem object itself --, first segment that reads it, whole_output: bool)."""
    __slots__ = ("M", "N", "segs", "out", "a_mn", "b_mn", "alpha", "accumulate", "split_k", "block_n",
                 "signals", "wait", "tag")

    def __init__(self, M, N, segs, out, a_mn=False, b_mn=False, alpha=1.0, accumulate=False, split_k=1,
                 block_n=0, signals=False, wait=None, tag="dense"):
        self.M, self.N, self.segs, self.out = M, N, segs, out
        self.a_mn, self.b_mn, self.alpha, self.accumulate = a_mn, b_mn, alpha, accumulate
        self.split_k, self.signals, self.wait, self.tag = split_k, signals, wait, tag
        if block_n == 0:
            block_n = 256 if N > 128 else (128 if (N > 64 or b_mn) else 64)
        self.block_n = block_n

    def flops(self):
        return 2.0 * self.M * self.N * sum((s[3] if len(s) > 3 else s[2]) for s in self.segs)


_GROUP_SCRATCH = {}
_GROUP_SCRATCH_INTS = 1 << 16


def _group_scratch(device):
    """Zero-initialised int32 scratch per (device, stream): the kernel leaves it zero again."""
    st = L.stream()
    key = (device.type, device.index, getattr(st, "value", None))
    buf = _GROUP_SCRATCH.get(key)
    if buf is None:
        buf = _GROUP_SCRATCH[key] = torch.zeros(_GROUP_SCRATCH_INTS, dtype=torch.int32, device=device)
    return buf


def gemm_grouped(problems):
    """Run `problems` (list of Problem, dependency producers first) in ONE persistent tcgen05 launch."""
    n = len(problems)
    if n > L.GROUPED_MAX_PROBLEMS:
        raise RuntimeError("unsloth_b200.gemm_grouped: at most %d problems per launch" % L.GROUPED_MAX_PROBLEMS)
    arr = (L.GemmProblem * n)()
    keep = []
    ab_dtype = problems[0].segs[0][0].dtype
    dev = problems[0].out.device
    need = 1
    for i, pr in enumerate(problems):
        ns = len(pr.segs)
        if ns > L.GROUPED_MAX_SEGMENTS:
            raise RuntimeError("unsloth_b200.gemm_grouped: at most %d segments per problem" % L.GROUPED_MAX_SEGMENTS)
        sa = (L.GemmSegment * ns)()
        for j, sg in enumerate(pr.segs):
            A, B, K = sg[0], sg[1], sg[2]
            _check_operand(A); _check_operand(B)
            if A.dtype != ab_dtype or B.dtype != ab_dtype:
                raise RuntimeError("unsloth_b200.gemm_grouped: mixed operand dtypes %s/%s" % (A.dtype, B.dtype))
            sa[j].a = A.data_ptr(); sa[j].lda = A.stride(0)
            sa[j].b = B.data_ptr(); sa[j].ldb = B.stride(0)
            sa[j].k = K
        keep.append(sa)
        g = arr[i]
        g.M, g.N, g.segs, g.n_segs = pr.M, pr.N, sa, ns
        g.a_mn_major, g.b_mn_major = int(pr.a_mn), int(pr.b_mn)
        g.C, g.ldc, g.c_dtype = pr.out.data_ptr(), pr.out.stride(0), L.dt(pr.out)
        g.alpha, g.accumulate = float(pr.alpha), int(pr.accumulate)
        kb = sum((sg[2] + 63) // 64 for sg in pr.segs)
        split = max(1, min(int(pr.split_k), kb))
        g.split_k = split
        if split > 1:
            ws = torch.empty(split * pr.M * pr.N, dtype=torch.float32, device=dev)
            keep.append(ws)
            g.workspace = ws.data_ptr()
        g.block_n, g.signals = pr.block_n, int(pr.signals)
        dep = pr.wait
        if dep is not None and isinstance(dep[0], Problem):
            # producer given by identity: a producer that is NOT part of this launch has already
            # completed in an earlier launch on the stream -- nothing to wait for
            idx = next((k for k, q_ in enumerate(problems) if q_ is dep[0]), -1)
            dep = None if idx < 0 else (idx, dep[1], dep[2])
        if dep is None:
            g.wait_problem, g.wait_segment, g.wait_all = -1, 0, 0
        else:
            g.wait_problem, g.wait_segment, g.wait_all = int(dep[0]), int(dep[1]), int(bool(dep[2]))
        m_pairs = (pr.M + 255) // 256
        need += m_pairs + 1 + (m_pairs * ((pr.N + pr.block_n - 1) // pr.block_n) * 8 if split > 1 else 0)
    if need > _GROUP_SCRATCH_INTS:
        raise RuntimeError("unsloth_b200.gemm_grouped: scratch too small (%d ints needed)" % need)
    scratch = _group_scratch(dev)
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.call("ub200_gemm_grouped", arr, n, L.dt(ab_dtype), L.ptr(scratch), L.stream())
    if ev is not None:
        e1.record()
        ev.append((sum(pr.flops() for pr in problems), e0, e1,
                   {"kernel": "grouped", "dense_flops": sum(pr.flops() for pr in problems if pr.tag == "dense"),
                    "rank_flops": sum(pr.flops() for pr in problems if pr.tag != "dense"), "problems": n}))
    return [pr.out for pr in problems]


def fused_dequant_enabled() -> bool:
    """UB200_FUSED_DEQUANT=1: forward projections expand the NF4 weight inside the GEMM's operand
    staging (csrc/gemm_nf4.cu) instead of dequantising to a 16-bit buffer first.  Measured trade-off in
    DESIGN.md section 4.2; default off."""
    return os.environ.get("UB200_FUSED_DEQUANT", "0") == "1"


def gemm_nf4(X2, W, quant_state, out, lora=None):
    """out[T, N] = X2[T, K] @ dequant(W)^T (+ XA @ B_pad^T): NF4 expansion fused into the tcgen05 GEMM.
    lora: None or (XA [T, Rp], B_pad [N, Rp] scaled, Rp)."""
    absmax, shape, dtype, blocksize, offset, absmax2, code2, blocksize2 = _unpack_quant_state(quant_state)
    if not torch.is_tensor(offset):
        offset = torch.tensor(float(offset), dtype=torch.float32, device=W.device)
    T, K = X2.shape
    N = shape[0]
    xa, bp, rk = (None, None, 0) if lora is None else lora
    ev = GEMM_EVENTS
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.call("ub200_gemm_nf4", T, N, K, L.ptr(X2), X2.stride(0), L.ptr(W), L.ptr(absmax), L.ptr(code2),
           L.ptr(absmax2), L.ptr(offset), int(blocksize), int(blocksize2), L.ptr(xa),
           0 if xa is None else xa.stride(0), L.ptr(bp), 0 if bp is None else bp.stride(0), int(rk),
           L.ptr(out), out.stride(0), L.dt(out), L.stream())
    if ev is not None:
        e1.record()
        ev.append((2.0 * T * N * K, e0, e1, {"kernel": "gemm_nf4"}))
    return out


def cast_pad(src, dst, row_off=0, col_off=0, scale=1.0, transpose=False):
    """dst (2-D, fully overwritten) <- zeros with scale*src (optionally transposed) placed at
    (row_off, col_off)."""
    L.call("ub200_cast_pad_2d", L.ptr(src), L.dt(src), src.stride(0), src.shape[0], src.shape[1],
           L.ptr(dst), L.dt(dst), dst.stride(0), dst.shape[0], dst.shape[1], int(row_off),
           int(col_off), float(scale), int(transpose), L.stream())
    return dst


# Cached 16-bit casts of the LoRA parameters.  Validity rules:
#   * every FORWARD of a projection rebuilds its casts (`refresh=True`) -- a cast is only ever
#     REUSED by the backward of the same forward (and by a later segment of the same step), so an
#     optimiser that writes parameters behind autograd's back (`p.data.add_()`, raw pointers, 8-bit
#     / fused optimisers: none of them bump `tensor._version`) can never leave a stale adapter in a
#     training step;
#   * lookups outside a forward (backward, decode-time `_fast_lora`) are keyed by
#     (`_version`, PARAM_EPOCH, `data_ptr`); PARAM_EPOCH is bumped by ub200_adamw_flat's wrapper
#     and by a global `torch.optim.Optimizer` post-step hook, and `bump_param_epoch()` is public
#     for anything else that edits parameters in place between decode calls.
PARAM_EPOCH = 0


def bump_param_epoch():
    global PARAM_EPOCH
    PARAM_EPOCH += 1


def _optimizer_post_step(optimizer, args, kwargs):
    bump_param_epoch()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_optimizer_post_step)
except Exception:  # pragma: no cover - very old torch
    pass


class StepPlan:
    """Batched housekeeping of a bucketed training step (opt-in: `FlatLoRABucket.begin_step()` /
    `end_backward()`, used by `GraphedTrainStep`, `bench.py` and `smoke()`).  A Llama-3-8B step makes 448
    LoRA casts and hands 448 LoRA gradients to autograd's AccumulateGrad, each a 4-5 us launch (1.8 % of
    the step, profiles/r2_launches_bench_full_summary.txt).  With a plan:
      * casts: every cast `cached_cast_pad` makes of a LoRA Parameter during the first step is
        remembered with its destination; `refresh()` rebuilds ALL of them at the start of each later step
        (ub200_cast_pad_multi: one launch per 40 tensors) and the calls inside that step return the
        refreshed buffers.  Outside `in_step` nothing changes: every forward still rebuilds its casts.
      * gradients: d_A / d_B of a parameter whose `.grad` is a preallocated contiguous fp32 tensor (the
        flat bucket) are collected instead of being returned to autograd and added by `flush_grads()`
        (ub200_accumulate_multi) -- the same `grad += view` arithmetic, bit for bit."""

    def __init__(self, params):
        self.param_ids = {id(p) for p in params}
        self.entries = {}            # (id(src), key) -> [src, key, dst, stamp, parts]
        self.stamp = 0
        self.in_step = False
        self.recorded = False
        self._cast_arr = None
        self.grad_pairs = []

    # -- casts ---------------------------------------------------------------------------------
    def lookup(self, src, key):
        ent = self.entries.get((id(src), key))
        if ent is not None and ent[3] == self.stamp and ent[0] is src:
            return ent[2]
        return None

    def record(self, src, key, dst, parts=None):
        """dst was just built from `src` (key = the cast's arguments); parts: [(src_j, dst_view_j)] when several
        adapters are cast into row slices of one shared block (plain casts, no offset / scale)."""
        if parts is None:
            shape, dtype, row_off, col_off, scale, transpose = key
            parts = [(src, dst, row_off, col_off, scale, transpose)]
        else:
            parts = [(a, v, 0, 0, 1.0, False) for a, v in parts]
        if all(id(a) in self.param_ids and a.dim() == 2 and a.stride(1) == 1 for a, *_ in parts):
            self.entries[(id(src), key)] = [src, key, dst, self.stamp, parts]
            self._cast_arr = None

    def refresh(self):
        """Start of a step: rebuild every recorded cast from the current parameter values."""
        self.stamp += 1
        if not self.entries:
            return
        if self._cast_arr is None:
            ents = list(self.entries.values())
            parts = [pt for e in ents for pt in e[4]]
            arr = (L.CastDesc * len(parts))()
            for d, (src, dst, row_off, col_off, scale, transpose) in zip(arr, parts):
                d.src, d.dst = src.data_ptr(), dst.data_ptr()
                d.src_ld, d.dst_ld = src.stride(0), dst.stride(0)
                d.src_dtype, d.dst_dtype = L.dt(src), L.dt(dst)
                d.rows, d.cols = src.shape[0], src.shape[1]
                d.dst_rows, d.dst_cols = dst.shape[0], dst.shape[1]
                d.row_off, d.col_off, d.scale, d.transpose = int(row_off), int(col_off), float(scale), int(transpose)
            self._cast_arr = (arr, ents, parts)
        arr, ents, parts = self._cast_arr
        for d, (src, dst, *_r) in zip(arr, parts):
            if src.data_ptr() != d.src or src.stride(0) != d.src_ld:
                raise RuntimeError("unsloth_b200.StepPlan: a recorded LoRA parameter moved or changed layout")
        L.call("ub200_cast_pad_multi", arr, len(parts), L.stream())
        L.launch_count += (len(parts) - 1) // 40
        for e in ents:
            e[3] = self.stamp

    # -- gradients -----------------------------------------------------------------------------
    def sink(self, param, view):
        g = getattr(param, "grad", None)
        if (not self.in_step or view is None or id(param) not in self.param_ids or g is None
                or g.dtype != torch.float32 or view.dtype != torch.float32 or not g.is_contiguous()
                or g.shape != view.shape or view.dim() != 2):
            return False
        self.grad_pairs.append((view, g))
        return True

    def flush_grads(self):
        n = len(self.grad_pairs)
        if not n:
            return
        arr = (L.AccDesc * n)()
        for d, (view, g) in zip(arr, self.grad_pairs):
            d.src, d.dst = view.data_ptr(), g.data_ptr()
            d.src_rs, d.src_cs = view.stride(0), view.stride(1)
            d.rows, d.cols = g.shape[0], g.shape[1]
        L.call("ub200_accumulate_multi", arr, n, L.stream())
        L.launch_count += (n - 1) // 40
        self.grad_pairs = []


ACTIVE_PLAN = None


def sink_lora_grads(A, B, dA, dB):
    """(dA, dB) as autograd outputs -- or (None, None) when the active StepPlan took them."""
    plan = ACTIVE_PLAN
    if plan is not None and plan.in_step and plan.sink(A, dA):
        if plan.sink(B, dB):
            return None, None
        return None, dB
    return dA, dB


def cached_cast_pad(src, shape, dtype, row_off=0, col_off=0, scale=1.0, transpose=False,
                    refresh=False):
    """`cast_pad` into a fresh [shape] tensor, memoised ON the source nn.Parameter so that the
    backward reuses the cast its forward made.  `refresh=True` (every forward) rebuilds it
    unconditionally; see PARAM_EPOCH above.  Non-Parameter sources are never cached."""
    if not isinstance(src, torch.nn.Parameter):
        return cast_pad(src, torch.empty(shape, dtype=dtype, device=src.device), row_off, col_off,
                        scale, transpose)
    cache = src.__dict__.setdefault("_ub200_cast_cache", {})
    key = (tuple(shape), dtype, row_off, col_off, float(scale), bool(transpose))
    plan = ACTIVE_PLAN
    if plan is not None and plan.in_step:
        dst = plan.lookup(src, key)          # rebuilt by plan.refresh() at the start of THIS step
        if dst is not None:
            return dst
    ver = (src._version, PARAM_EPOCH, src.data_ptr())
    hit = None if refresh else cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    dst = cast_pad(src, torch.empty(shape, dtype=dtype, device=src.device), row_off, col_off, scale,
                   transpose)
    cache[key] = (ver, dst)
    if plan is not None and plan.in_step and src.dim() == 2 and src.stride(1) == 1:
        plan.record(src, key, dst)
    return dst


def keep_dequant() -> bool:
    """Memory-for-bandwidth policy of the training step (B200: 180 GB of HBM3e).  When on, the
    16-bit expansion of an NF4 weight made for the forward GEMM stays resident until that layer's
    backward instead of being rebuilt there (the reference re-dequantises, fast_lora.py:193-204):
    half of the step's dequant launches disappear for +2 bytes per base parameter of peak memory
    (13.96 GB on Llama-3-8B).  UB200_KEEP_DEQUANT=1/0 forces it; the default turns it on for
    devices with at least 128 GiB."""
    global _KEEP_DEQUANT
    if _KEEP_DEQUANT is None:
        env = os.environ.get("UB200_KEEP_DEQUANT", "auto").lower()
        if env in ("0", "off", "false"):
            _KEEP_DEQUANT = False
        elif env in ("1", "on", "true"):
            _KEEP_DEQUANT = True
        else:
            _KEEP_DEQUANT = (torch.cuda.is_available()
                             and torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
                             >= 128 * 2 ** 30)
    return _KEEP_DEQUANT


# set by patch.install(gradient_checkpointing=True): with per-layer recompute the forward's 16-bit
# expansions must not stay alive across layers (that is the memory the recompute is meant to save)
KEEP_DEQUANT_BLOCKED = False


import threading as _threading

_OUTER = _threading.local()


class GradModeAware:
    """Mixin for autograd Functions: records the caller's grad mode around `apply`.  Inside
    `Function.forward` grad mode is always off and `ctx.needs_input_grad` stays True under
    `torch.no_grad()`, so this is the only way a forward can tell that no backward will follow."""

    @classmethod
    def apply(cls, *args, **kwargs):
        prev = getattr(_OUTER, "grad", None)
        _OUTER.grad = torch.is_grad_enabled()
        try:
            return super().apply(*args, **kwargs)
        finally:
            _OUTER.grad = prev


def outer_grad_enabled() -> bool:
    g = getattr(_OUTER, "grad", None)
    return True if g is None else g


def keep_for_backward(needs_input_grad) -> bool:
    """Should this forward keep its dequantised weights for the backward?  Only when a backward
    will actually run (caller in grad mode, some input needs a gradient) and per-layer recompute
    is not in charge of the memory."""
    return (keep_dequant() and not KEEP_DEQUANT_BLOCKED and outer_grad_enabled()
            and any(needs_input_grad))


def set_keep_dequant(value):
    """True / False / None (= re-read UB200_KEEP_DEQUANT / the device-size default)."""
    global _KEEP_DEQUANT
    _KEEP_DEQUANT = value


_KEEP_DEQUANT = None


def dense_weight(W, W_quant, dtype, slot=0, fresh=False):
    """Logical [N_out, K_in] weight in the compute dtype, exactly what the reference's
    `fast_dequantize(W, W_quant)` hands to `torch.matmul(X, W.t())`: a reusable-slot dequantised
    buffer (a transposed VIEW of it when the packed weight was passed as `W.t()`), or W itself
    for 16-bit LoRA.  `fresh`: a private tensor the caller may keep (see keep_dequant)."""
    if W_quant is None:
        Wd = W
    elif fresh:
        Wd = fast_dequantize(W, W_quant)
    else:
        Wd = fast_dequantize(W, W_quant, use_global_buffer=True, _slot=slot)
    if Wd.dtype != dtype:
        Wd = Wd.to(dtype)
    return Wd


def as_b_operand(Wd):
    """(tensor, b_mn) for a logical [N, K] matrix: K-major if rows are contiguous, MN-major
    (the [K, N] row-major buffer underneath) if it is a transposed view."""
    if Wd.stride(-1) == 1:
        return Wd, False
    if Wd.stride(0) == 1:
        return Wd.t(), True
    return Wd.contiguous(), False


def matmul_lora(X, W, W_quant, A, B, s, out=None):
    """out = X @ dequant(W).T + (X @ A.T) @ (s * B.T)   (kernels/utils.py:1128-1170).
    A: [r, K_in], B: [N_out, r] (possibly transposed views, as the reference's backward passes
    them).  One tcgen05 launch for the sum: the LoRA update rides as an extra K block of the
    same fp32 accumulator (single rounding; the reference rounds the base product to bf16
    first, then `addmm_`)."""
    L.require_cuda(X)
    dtype = X.dtype
    reshape = X.dim() == 3
    if reshape:
        batch, seq_len, _ = X.shape
    X2 = X.reshape(-1, X.shape[-1])
    if X2.stride(-1) != 1:
        X2 = X2.contiguous()
    T, K = X2.shape
    Wd = dense_weight(W, W_quant, dtype, 0)
    N = Wd.shape[0]
    assert Wd.shape[1] == K, (tuple(Wd.shape), tuple(X2.shape))
    Bop, b_mn = as_b_operand(Wd)
    out2 = torch.empty((T, N), dtype=dtype, device=X.device) if out is None else out.reshape(T, N)
    segs = [(X2, Bop, K)]
    if A is not None:
        A = A if A.stride(-1) == 1 else A.contiguous()
        B = B if B.stride(-1) == 1 else B.contiguous()
        r = A.shape[0]
        Rp = ((r + RANK_BLOCK - 1) // RANK_BLOCK) * RANK_BLOCK
        A_pad = cast_pad(A, torch.empty((Rp, K), dtype=dtype, device=X.device))
        XA = gemm(T, Rp, [(X2, A_pad, K)], torch.empty((T, Rp), dtype=dtype, device=X.device))
        if b_mn:
            B_pad = cast_pad(B, torch.empty((Rp, N), dtype=dtype, device=X.device), scale=s,
                             transpose=True)
        else:
            B_pad = cast_pad(B, torch.empty((N, Rp), dtype=dtype, device=X.device), scale=s)
        segs.append((XA, B_pad, Rp))
    gemm(T, N, segs, out2, a_mn=False, b_mn=b_mn)
    return out2.view(batch, seq_len, -1) if reshape else out2


# ---------------------------------------------------------------------------------------------
# PEFT attribute scraper (host only) -- kernels/utils.py:335-440
# ---------------------------------------------------------------------------------------------
def get_lora_parameters(proj):
    """Return (W, W_quant, A, B, scaling); A = B = scaling = None when adapters are disabled
    or merged (kernels/utils.py:335-397)."""
    base_layer = getattr(proj, "base_layer", proj)
    W = base_layer.weight
    W_quant = getattr(W, "quant_state", None)
    if getattr(proj, "disable_adapters", True) or proj.merged:
        return W, W_quant, None, None, None
    adapter = getattr(proj, "active_adapters", None)
    if adapter is None:
        adapter = getattr(proj, "active_adapter", ("default"))
    adapter = adapter[0]
    return (W, W_quant, proj.lora_A[adapter].weight, proj.lora_B[adapter].weight,
            proj.scaling[adapter])


def get_lora_parameters_bias(proj):
    """kernels/utils.py:400-440."""
    base_layer = getattr(proj, "base_layer", proj)
    W = base_layer.weight
    W_quant = getattr(W, "quant_state", None)
    if getattr(proj, "disable_adapters", True) or proj.merged:
        return W, W_quant, None, None, None, base_layer.bias
    adapter = getattr(proj, "active_adapters", None)
    if adapter is None:
        adapter = getattr(proj, "active_adapter", ("default"))
    adapter = adapter[0]
    return (W, W_quant, proj.lora_A[adapter].weight, proj.lora_B[adapter].weight,
            proj.scaling[adapter], base_layer.bias)
