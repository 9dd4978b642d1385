// Rotary position embedding, in place, Q and K in ONE launch, forward and backward.
//
// Replaces the reference's Triton kernels
//   unsloth/kernels/rope_embedding.py:104-166 (_rope_embedding, no position ids)
//   unsloth/kernels/rope_embedding.py:23-98   (_rope_embedding_QK, strided + indices)
// Both forms are one strided kernel here: element (b, h, s, d) lives at
//   base + b*batch_stride + h*head_stride + s*seq_stride + d.
// The no-index form on the contiguous [B,S,H*D] projection buffer is just
// (batch_stride, head_stride, seq_stride) = (S*H*D, D, H*D); no clone, no host sync
// (the reference synchronises the stream per layer when DEVICE_COUNT>1, :278-279).
//
// HBM-bound: each CTA handles one token row (b,s); every thread owns one 16-byte vector of
// the first half of a head and its partner vector in the second half, so all global
// accesses are 16-byte and coalesced over d.  cos/sin are read once per row into
// registers and reused across all Q and K heads.  Algorithmic bytes per token:
//   2*(Hq+Hk)*D*bytes (read+write) + D*table_bytes (first halves of cos and sin).
//
// `compute_dtype` reproduces the reference's rounding: the no-index
// kernel evaluates in the TABLE dtype, the QK kernel in the promoted dtype.  Within that dtype
// the roundings fall where the reference's natively compiled Triton kernel puts them (one product
// rounded, the other fused into the sum by fma) -- measured on a B200, not assumed.
#include <cstdlib>

#include "common.cuh"

// default rounding modes of the packed 16-bit kernel (see rope_packed_kernel)
#ifndef UB200_ROPE_MODE_BF16
#define UB200_ROPE_MODE_BF16 4
#endif
#ifndef UB200_ROPE_MODE_F16
#define UB200_ROPE_MODE_F16 4
#endif

namespace ub {

template <typename T>
__global__ void __launch_bounds__(512) rope_kernel(
    T* Q, int64_t q_bs, int64_t q_hs, int64_t q_ss, T* K, int64_t k_bs, int64_t k_hs,
    int64_t k_ss, const void* __restrict__ cos, int64_t cos_rs, const void* __restrict__ sin,
    int64_t sin_rs, const int32_t* __restrict__ indices, int seqlen, int n_heads_q,
    int n_heads_k, int head_dim, int backward, int table_dt, int comp_dt, int64_t n_rows) {
  constexpr int V = DT<T>::VEC;
  const int half = head_dim >> 1;
  const int vec_per_half = half / V;                 // threads per head
  const int heads_per_pass = blockDim.x / vec_per_half;
  const int lane_v = threadIdx.x % vec_per_half;     // which vector inside the half
  const int head_in_pass = threadIdx.x / vec_per_half;
  const int d0 = lane_v * V;
  const int total_heads = n_heads_q + n_heads_k;

  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int b = (int)(row / seqlen);
    const int s = (int)(row - (int64_t)b * seqlen);
    const int pos = indices ? indices[row] : s;
    float c[V], sn[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      c[i] = load_as_f(cos, table_dt, (int64_t)pos * cos_rs + d0 + i);
      float sv = load_as_f(sin, table_dt, (int64_t)pos * sin_rs + d0 + i);
      sn[i] = backward ? -sv : sv;
    }
    if (head_in_pass >= heads_per_pass) continue;
    for (int h = head_in_pass; h < total_heads; h += heads_per_pass) {
      T* p = (h < n_heads_q)
                 ? Q + (int64_t)b * q_bs + (int64_t)h * q_hs + (int64_t)s * q_ss
                 : K + (int64_t)b * k_bs + (int64_t)(h - n_heads_q) * k_hs + (int64_t)s * k_ss;
      float x1[V], x2[V], o1[V], o2[V];
      load_vec<T>(p + d0, x1);
      load_vec<T>(p + half + d0, x2);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float a1 = round_to(comp_dt, x1[i]), a2 = round_to(comp_dt, x2[i]);
        // q1*cos - q2*sin ; q2*cos + q1*sin, each op rounded to the compute dtype
        // the contraction LLVM applies to the reference kernel (see rope_packed_kernel, MODE 4):
        // second product rounded, first one fused into the sum
        const float m22 = round_to(comp_dt, __fmul_rn(a2, sn[i]));
        const float m21 = round_to(comp_dt, __fmul_rn(a2, c[i]));
        o1[i] = round_to(comp_dt, __fmaf_rn(a1, c[i], -m22));
        o2[i] = round_to(comp_dt, __fmaf_rn(a1, sn[i], m21));
      }
      store_vec<T>(p + d0, o1);
      store_vec<T>(p + half + d0, o2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lean 16-bit specialisation: activations, tables and compute dtype all bf16 (or all fp16) -- the
// Llama / Mistral case.  The reference evaluates q1*cos - q2*sin in the table dtype, i.e. every
// product and the sum round to bf16; that is exactly one packed HMUL2 / HSUB2 / HADD2 per two
// elements (the `_rn` forms forbid contraction into HFMA2).  ncu on the generic kernel showed
// 68 % SM-pipe utilisation from scalar rounding emulation on a kernel that should wait on HBM.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Pk2;
template <> struct Pk2<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ float2 to_f2(T2 v) { return __bfloat1622float2(v); }
  __device__ static __forceinline__ T2 from_f2(float2 v) { return __float22bfloat162_rn(v); }
};
template <> struct Pk2<__half> {
  using T2 = __half2;
  __device__ static __forceinline__ float2 to_f2(T2 v) { return __half22float2(v); }
  __device__ static __forceinline__ T2 from_f2(float2 v) { return __float22half2_rn(v); }
};

// MODE selects where the 16-bit roundings fall (measured against the reference's Triton kernel run
// natively on a B200, tests/test_gpu_vs_reference.py / benchmarks/probe_ref_numerics.py):
//   0  every product and the sum rounded (what TRITON_INTERPRET=1 / numpy does: fp16 goldens)
//   1  LLVM's fp-contract=fast form of `a*c - b*s`:  fma(a, c, -rn(b*s)) ; fma(b, c, rn(a*s))
//   2  the other contraction:                        fma(-b, s, rn(a*c)) ; fma(a, s, rn(b*c))
//   3  fp32 evaluation, one rounding at the store
//   4  what Triton 3.6 / LLVM emits for the reference kernel on sm_100 (read off its PTX,
//      profiles/r2_triton_rope_ptx.txt):  o1 = fma(a, c, -rn(b*s)) ; o2 = fma(a, s, rn(b*c))
//      -- bit-identical to the reference's native bf16 AND fp16 output (the DEFAULT)
template <typename T, int HP, int MODE>
__global__ void __launch_bounds__(512) rope_packed_kernel(
    T* Q, int64_t q_bs, int64_t q_hs, int64_t q_ss, T* K, int64_t k_bs, int64_t k_hs,
    int64_t k_ss, const T* __restrict__ cos, int64_t cos_rs, const T* __restrict__ sin,
    int64_t sin_rs, const int32_t* __restrict__ indices, int seqlen, int n_heads_q,
    int n_heads_k, int head_dim, int backward, int64_t n_rows) {
  using T2 = typename Pk2<T>::T2;
  union V16 { int4 q; T2 h[4]; };
  const int half = head_dim >> 1;
  const int vec_per_half = half / 8;
  const int heads_per_pass = blockDim.x / vec_per_half;
  const int lane_v = threadIdx.x % vec_per_half;
  const int head_in_pass = threadIdx.x / vec_per_half;
  const int d0 = lane_v * 8;
  const int total_heads = n_heads_q + n_heads_k;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int b = (int)(row / seqlen);
    const int s = (int)(row - (int64_t)b * seqlen);
    const int pos = indices ? indices[row] : s;
    V16 c, sn;
    c.q = *reinterpret_cast<const int4*>(cos + (int64_t)pos * cos_rs + d0);
    sn.q = *reinterpret_cast<const int4*>(sin + (int64_t)pos * sin_rs + d0);
    if (backward) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sn.h[i] = __hneg2(sn.h[i]);
    }
    if (head_in_pass >= heads_per_pass) continue;
    for (int h0 = head_in_pass; h0 < total_heads; h0 += HP * heads_per_pass) {
      T* p[HP];
      V16 x1[HP], x2[HP];
#pragma unroll
      for (int u = 0; u < HP; ++u) {
        const int h = h0 + u * heads_per_pass;
        p[u] = nullptr;
        if (h < total_heads) {
          p[u] = (h < n_heads_q)
                     ? Q + (int64_t)b * q_bs + (int64_t)h * q_hs + (int64_t)s * q_ss
                     : K + (int64_t)b * k_bs + (int64_t)(h - n_heads_q) * k_hs + (int64_t)s * k_ss;
          x1[u].q = *reinterpret_cast<const int4*>(p[u] + d0);
          x2[u].q = *reinterpret_cast<const int4*>(p[u] + half + d0);
        }
      }
#pragma unroll
      for (int u = 0; u < HP; ++u) {
        if (p[u] == nullptr) continue;
        V16 o1, o2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (MODE == 0) {
            o1.h[i] = __hsub2_rn(__hmul2_rn(x1[u].h[i], c.h[i]), __hmul2_rn(x2[u].h[i], sn.h[i]));
            o2.h[i] = __hadd2_rn(__hmul2_rn(x2[u].h[i], c.h[i]), __hmul2_rn(x1[u].h[i], sn.h[i]));
          } else if (MODE == 1) {
            o1.h[i] = __hfma2(x1[u].h[i], c.h[i], __hneg2(__hmul2_rn(x2[u].h[i], sn.h[i])));
            o2.h[i] = __hfma2(x2[u].h[i], c.h[i], __hmul2_rn(x1[u].h[i], sn.h[i]));
          } else if (MODE == 2) {
            o1.h[i] = __hfma2(__hneg2(x2[u].h[i]), sn.h[i], __hmul2_rn(x1[u].h[i], c.h[i]));
            o2.h[i] = __hfma2(x1[u].h[i], sn.h[i], __hmul2_rn(x2[u].h[i], c.h[i]));
          } else if (MODE == 4) {
            o1.h[i] = __hfma2(x1[u].h[i], c.h[i], __hneg2(__hmul2_rn(x2[u].h[i], sn.h[i])));
            o2.h[i] = __hfma2(x1[u].h[i], sn.h[i], __hmul2_rn(x2[u].h[i], c.h[i]));
          } else {
            const float2 a = Pk2<T>::to_f2(x1[u].h[i]), b = Pk2<T>::to_f2(x2[u].h[i]);
            const float2 cc = Pk2<T>::to_f2(c.h[i]), ss = Pk2<T>::to_f2(sn.h[i]);
            o1.h[i] = Pk2<T>::from_f2(make_float2(__fmaf_rn(a.x, cc.x, -__fmul_rn(b.x, ss.x)), __fmaf_rn(a.y, cc.y, -__fmul_rn(b.y, ss.y))));
            o2.h[i] = Pk2<T>::from_f2(make_float2(__fmaf_rn(b.x, cc.x, __fmul_rn(a.x, ss.x)), __fmaf_rn(b.y, cc.y, __fmul_rn(a.y, ss.y))));
          }
        }
        *reinterpret_cast<int4*>(p[u] + d0) = o1.q;
        *reinterpret_cast<int4*>(p[u] + half + d0) = o2.q;
      }
    }
  }
}

}  // namespace ub

extern "C" int ub200_rope_qk(void* Q, int64_t q_batch_stride, int64_t q_head_stride,
                             int64_t q_seq_stride, void* K, int64_t k_batch_stride,
                             int64_t k_head_stride, int64_t k_seq_stride, const void* cos,
                             int64_t cos_row_stride, const void* sin, int64_t sin_row_stride,
                             const int32_t* indices, int batch, int seqlen, int n_heads_q,
                             int n_heads_k, int head_dim, int backward, int dtype,
                             int table_dtype, int compute_dtype, cudaStream_t stream) {
  using namespace ub;
  const int64_t n_rows = (int64_t)batch * seqlen;
  if (n_rows <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  const int half = head_dim / 2;
  if (head_dim % 2 || half % V) return UB200_ERR_BAD_ARG;
  if (q_batch_stride % V || q_head_stride % V || q_seq_stride % V) return UB200_ERR_BAD_ARG;
  if (K == nullptr) n_heads_k = 0;
  if (n_heads_k && (k_batch_stride % V || k_head_stride % V || k_seq_stride % V))
    return UB200_ERR_BAD_ARG;
  const int vec_per_half = half / V;
  if (vec_per_half > 512) return UB200_ERR_UNSUPPORTED;
  // one pass over all heads if it fits in 256 threads
  int heads = n_heads_q + n_heads_k;
  int threads = vec_per_half * heads;
  if (threads > 512) threads = (512 / vec_per_half) * vec_per_half;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  int64_t g = (int64_t)UB_SM_COUNT * 8;
  const int grid = (int)(n_rows < g ? n_rows : g);
  // lean packed path: activations, tables and compute dtype all the same 16-bit type
  if (dtype != UB200_F32 && table_dtype == dtype && compute_dtype == dtype && (cos_row_stride % 8) == 0 &&
      (sin_row_stride % 8) == 0 && ((reinterpret_cast<uintptr_t>(cos) | reinterpret_cast<uintptr_t>(sin) |
                                     reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K)) & 15) == 0) {
    constexpr int HP = 2;
    int pthreads = vec_per_half * ((heads + HP - 1) / HP);
    if (pthreads > 512) pthreads = (512 / vec_per_half) * vec_per_half;
    pthreads = ((pthreads + 31) / 32) * 32;
    if (pthreads > 512) pthreads = 512;
    const int64_t pg = (int64_t)UB_SM_COUNT * 12;
    const int pgrid = (int)(n_rows < pg ? n_rows : pg);
    // rounding mode: see rope_packed_kernel.  UB200_ROPE_MODE overrides (numerics probes only).
    static const int env_mode = [] { const char* e = getenv("UB200_ROPE_MODE"); return e ? atoi(e) : -1; }();
    const int mode = env_mode >= 0 ? env_mode : (dtype == UB200_BF16 ? UB200_ROPE_MODE_BF16 : UB200_ROPE_MODE_F16);
#define GOP(T, M)                                                                                   \
  rope_packed_kernel<T, HP, M><<<pgrid, pthreads, 0, stream>>>(                                     \
      (T*)Q, q_batch_stride, q_head_stride, q_seq_stride, (T*)K, k_batch_stride, k_head_stride,     \
      k_seq_stride, (const T*)cos, cos_row_stride, (const T*)sin, sin_row_stride, indices, seqlen, \
      n_heads_q, n_heads_k, head_dim, backward, n_rows)
#define GOM(T)                                                                                      \
  do {                                                                                              \
    if (mode == 1) GOP(T, 1); else if (mode == 2) GOP(T, 2); else if (mode == 3) GOP(T, 3); else if (mode == 0) GOP(T, 0); else GOP(T, 4); \
  } while (0)
    if (dtype == UB200_BF16) GOM(__nv_bfloat16);
    else GOM(__half);
#undef GOM
#undef GOP
    UB_RETURN_LAST();
  }
#define GO(T)                                                                                  \
  rope_kernel<T><<<grid, threads, 0, stream>>>(                                                \
      (T*)Q, q_batch_stride, q_head_stride, q_seq_stride, (T*)K, k_batch_stride, k_head_stride, \
      k_seq_stride, cos, cos_row_stride, sin, sin_row_stride, indices, seqlen, n_heads_q,       \
      n_heads_k, head_dim, backward, table_dtype, compute_dtype, n_rows)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GO(__half); }
  else if (dtype == UB200_F32) { GO(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
  UB_RETURN_LAST();
}
