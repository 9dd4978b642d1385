// Cross entropy on materialised logits: forward (loss + logsumexp) and in-place backward.
//
// Replaces the reference's Triton kernels
//   unsloth/kernels/cross_entropy_loss.py:35-111  (_cross_entropy_forward)
//   unsloth/kernels/cross_entropy_loss.py:114-199 (_chunked_cross_entropy_forward + the three
//                                                  torch launches at :368-370)
//   unsloth/kernels/cross_entropy_loss.py:202-285 (_cross_entropy_backward)
// The reference needs a separate "chunked" path for vocabularies above 65,536 (Llama-3's
// 128,256 included) because a Triton program holds the whole row; here ONE CTA streams a
// row of any length once with an online (max, sum-exp) pair per thread, so there is a
// single path and no host-side logsumexp combine.
//
// HBM-bound.  Algorithmic bytes per row: fwd V*b (+16), bwd 2*V*b.
#include "common.cuh"

#include <math_constants.h>

namespace ub {

__device__ __forceinline__ float ce_transform(float x, float softcap, float scale) {
  if (scale != 0.f) x = scale * x;
  if (softcap != 0.f) x = softcap * tanhf(x / softcap);
  return x;
}

// exp for the streaming loops: 16-bit logits use the MUFU exp2 path (relative error ~1e-6, the
// inputs themselves carry 2^-9); fp32 logits keep the accurate expf for the 1e-5 gate.
template <bool FAST>
__device__ __forceinline__ float ce_exp(float x) { return FAST ? __expf(x) : expf(x); }

// online logsumexp state merge
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float nm = fmaxf(m, m2);
  if (nm == -CUDART_INF_F) { m = nm; s = 0.f; return; }
  s = s * expf(m - nm) + s2 * expf(m2 - nm);
  m = nm;
}

template <typename T>
__global__ void __launch_bounds__(512, 2) ce_fwd_kernel(
    const T* __restrict__ logits, int64_t row_stride, const int64_t* __restrict__ labels,
    float* __restrict__ loss, float* __restrict__ lse_out, int64_t n_rows, int vocab,
    float softcap, float scale, int vec_ok) {
  constexpr int V = DT<T>::VEC;
  __shared__ float sm_m[32], sm_s[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* x = logits + row * row_stride;
    float m = -CUDART_INF_F, s = 0.f;
    if (vec_ok) {
      const int nvec = vocab / V;
      // software-pipelined: the two 16-byte loads of iteration k+1 are issued before the exps of
      // iteration k, so each thread always has HBM requests in flight (ncu: the unpipelined loop
      // was latency-bound at 0.50 of peak)
      const int step = 2 * blockDim.x;
      int4 raw[2], nxt[2];
      auto fetch = [&](int i, int4 (&dst)[2]) {
        const int i2 = i + blockDim.x;
        dst[0] = (i < nvec) ? __ldcs(reinterpret_cast<const int4*>(x + (int64_t)i * V)) : make_int4(0, 0, 0, 0);
        dst[1] = (i2 < nvec) ? __ldcs(reinterpret_cast<const int4*>(x + (int64_t)i2 * V)) : make_int4(0, 0, 0, 0);
      };
      fetch(tid, raw);
      for (int i = tid; i < nvec; i += step) {
        fetch(i + step, nxt);
        float v[2][V];
        const bool has2 = i + blockDim.x < nvec;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const Vec16<T> r = *reinterpret_cast<const Vec16<T>*>(&raw[u]);
#pragma unroll
          for (int k = 0; k < V; ++k) v[u][k] = DT<T>::to_f(r.v[k]);
        }
        float lm = -CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int k = 0; k < V; ++k) {
            v[u][k] = (u == 0 || has2) ? ce_transform(v[u][k], softcap, scale) : -CUDART_INF_F;
            lm = fmaxf(lm, v[u][k]);
          }
        }
        const float nm = fmaxf(m, lm);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          acc0 += ce_exp<sizeof(T) == 2>(v[0][k] - nm);
          acc1 += ce_exp<sizeof(T) == 2>(v[1][k] - nm);
        }
        s = s * ce_exp<sizeof(T) == 2>(m - nm) + (acc0 + acc1);
        m = nm;
        raw[0] = nxt[0];
        raw[1] = nxt[1];
      }
      for (int i = nvec * V + tid; i < vocab; i += blockDim.x) {
        const float v = ce_transform(DT<T>::to_f(x[i]), softcap, scale);
        lse_merge(m, s, v, 1.f);
      }
    } else {
      for (int i = tid; i < vocab; i += blockDim.x) {
        const float v = ce_transform(DT<T>::to_f(x[i]), softcap, scale);
        lse_merge(m, s, v, 1.f);
      }
    }
    // warp then block merge
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, m2, s2);
    }
    __syncthreads();
    if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
    __syncthreads();
    if (warp == 0) {
      m = lane < nw ? sm_m[lane] : -CUDART_INF_F;
      s = lane < nw ? sm_s[lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
      }
      if (lane == 0) {
        const float lse = m + logf(s);
        lse_out[row] = lse;
        const int64_t lab = labels[row];
        float l = 0.f;
        if (lab != -100) {
          const float xl = ce_transform(DT<T>::to_f(x[lab]), softcap, scale);
          l = lse - xl;
        }
        loss[row] = l;
      }
    }
  }
}

// grid: (row, column-chunk).  In place: logits <- dloss * d(loss)/d(logits)
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(
    T* logits, int64_t row_stride, const float* __restrict__ lse, const int64_t* __restrict__ labels,
    const float* __restrict__ dloss, int64_t dloss_stride, int vocab, float softcap, float scale,
    int vec_ok, int cols_per_block) {
  constexpr int V = DT<T>::VEC;
  const int64_t row = blockIdx.x;
  const int c0 = blockIdx.y * cols_per_block;
  const int c1 = min(vocab, c0 + cols_per_block);
  T* x = logits + row * row_stride;
  const int64_t lab = labels[row];
  const float dl = (lab != -100) ? dloss[row * dloss_stride] : 0.f;
  const float l = lse[row];
  auto grad = [&](float xv, int col) {
    if (scale != 0.f) xv = xv * scale;
    float partial = xv;
    if (softcap != 0.f) { partial = tanhf(xv / softcap); xv = softcap * partial; }
    float y = ce_exp<sizeof(T) == 2>(xv - l);
    if (col == lab) y -= 1.0f;
    if (scale != 0.f) y *= scale;
    if (softcap != 0.f) y *= (1.0f - partial * partial);
    return dl * y;
  };
  if (vec_ok) {
    for (int c = c0 + threadIdx.x * V; c < c1; c += blockDim.x * V) {
      float v[V];
      load_vec_cs<T>(x + c, v);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = grad(v[k], c + k);
      store_vec<T>(x + c, v);
    }
  } else {
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x)
      x[c] = DT<T>::from_f(grad(DT<T>::to_f(x[c]), c));
  }
}

// ---------------------------------------------------------------------------------------------
// Lean 16-bit paths without softcap / scaling (Llama, Mistral: the headline configuration).
// ncu on the generic kernels: 0.80-0.88 issued warps per scheduler -- issue-bound, not HBM-bound.
// Here the exponent is one FFMA + one MUFU.EX2 per element (exp(x - c) = 2^(x*log2e - c*log2e)),
// the running maximum is taken with packed HMNMX2 on the raw bf16 pairs, and the label column is
// patched once per vector instead of being compared per element.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Pk;
template <> struct Pk<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ float2 up(T2 v) { return __bfloat1622float2(v); }
  __device__ static __forceinline__ T2 down(float a, float b) { return __floats2bfloat162_rn(a, b); }
  __device__ static __forceinline__ T2 ninf() { return __float2bfloat162_rn(-CUDART_INF_F); }
};
template <> struct Pk<__half> {
  using T2 = __half2;
  __device__ static __forceinline__ float2 up(T2 v) { return __half22float2(v); }
  __device__ static __forceinline__ T2 down(float a, float b) { return __floats2half2_rn(a, b); }
  __device__ static __forceinline__ T2 ninf() { return __float2half2_rn(-CUDART_INF_F); }
};
// single MUFU.EX2 (exp2f() without -use_fast_math adds denormal range handling around it)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define UB_LOG2E 1.4426950408889634f
#define UB_LN2 0.6931471805599453f

template <typename T>
__global__ void __launch_bounds__(512, 2) ce_fwd_lean_kernel(
    const T* __restrict__ logits, int64_t row_stride, const int64_t* __restrict__ labels,
    float* __restrict__ loss, float* __restrict__ lse_out, int64_t n_rows, int vocab) {
  using T2 = typename Pk<T>::T2;
  union V16 { int4 q; T2 h[4]; };
  __shared__ float sm_m[32], sm_s[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int nvec = vocab / 8;
  const int step = 2 * blockDim.x;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* x = logits + row * row_stride;
    // state in the log2 domain: m2 = max(x)*log2e, s = sum 2^(x*log2e - m2)
    float m2 = -CUDART_INF_F, s = 0.f;
    V16 cur[2], nxt[2];
    auto fetch = [&](int i, V16 (&dst)[2]) {
      const int i2 = i + blockDim.x;
      if (i < nvec) dst[0].q = __ldcs(reinterpret_cast<const int4*>(x + (int64_t)i * 8));
      else { dst[0].h[0] = dst[0].h[1] = dst[0].h[2] = dst[0].h[3] = Pk<T>::ninf(); }
      if (i2 < nvec) dst[1].q = __ldcs(reinterpret_cast<const int4*>(x + (int64_t)i2 * 8));
      else { dst[1].h[0] = dst[1].h[1] = dst[1].h[2] = dst[1].h[3] = Pk<T>::ninf(); }
    };
    fetch(tid, cur);
    for (int i = tid; i < nvec; i += step) {
      fetch(i + step, nxt);
      T2 mx = __hmax2(__hmax2(cur[0].h[0], cur[0].h[1]), __hmax2(cur[0].h[2], cur[0].h[3]));
      mx = __hmax2(mx, __hmax2(__hmax2(cur[1].h[0], cur[1].h[1]), __hmax2(cur[1].h[2], cur[1].h[3])));
      const float2 mf = Pk<T>::up(mx);
      const float nm2 = fmaxf(m2, fmaxf(mf.x, mf.y) * UB_LOG2E);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = Pk<T>::up(cur[u].h[k]);
          a0 += fast_ex2(fmaf(f.x, UB_LOG2E, -nm2));
          a1 += fast_ex2(fmaf(f.y, UB_LOG2E, -nm2));
        }
      }
      s = s * fast_ex2(m2 - nm2) + (a0 + a1);
      m2 = nm2;
      cur[0].q = nxt[0].q;
      cur[1].q = nxt[1].q;
    }
    for (int i = nvec * 8 + tid; i < vocab; i += blockDim.x) {      // ragged tail
      const float t = DT<T>::to_f(x[i]) * UB_LOG2E;
      const float nm2 = fmaxf(m2, t);
      s = s * fast_ex2(m2 - nm2) + fast_ex2(t - nm2);
      m2 = nm2;
    }
    auto merge = [&](float om, float os) {
      const float nm = fmaxf(m2, om);
      if (nm == -CUDART_INF_F) { s = 0.f; return; }
      s = s * fast_ex2(m2 - nm) + os * fast_ex2(om - nm);
      m2 = nm;
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m2, o), os = __shfl_xor_sync(0xffffffffu, s, o);
      merge(om, os);
    }
    __syncthreads();
    if (lane == 0) { sm_m[warp] = m2; sm_s[warp] = s; }
    __syncthreads();
    if (warp == 0) {
      m2 = lane < nw ? sm_m[lane] : -CUDART_INF_F;
      s = lane < nw ? sm_s[lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m2, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        merge(om, os);
      }
      if (lane == 0) {
        const float lse = m2 * UB_LN2 + logf(s);
        lse_out[row] = lse;
        const int64_t lab = labels[row];
        loss[row] = (lab != -100) ? lse - DT<T>::to_f(x[lab]) : 0.f;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_lean_kernel(
    T* logits, int64_t row_stride, const float* __restrict__ lse, const int64_t* __restrict__ labels,
    const float* __restrict__ dloss, int64_t dloss_stride, int vocab, int cols_per_block) {
  using T2 = typename Pk<T>::T2;
  union V16 { int4 q; T2 h[4]; };
  const int64_t row = blockIdx.x;
  const int c0 = blockIdx.y * cols_per_block;
  const int c1 = min(vocab, c0 + cols_per_block);
  T* x = logits + row * row_stride;
  const int64_t lab = labels[row];
  const float dl = (lab != -100) ? dloss[row * dloss_stride] : 0.f;
  const float l2 = lse[row] * UB_LOG2E;
  for (int c = c0 + threadIdx.x * 8; c < c1; c += blockDim.x * 8) {
    V16 v, o;
    v.q = __ldcs(reinterpret_cast<const int4*>(x + c));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Pk<T>::up(v.h[k]);
      o.h[k] = Pk<T>::down(dl * fast_ex2(fmaf(f.x, UB_LOG2E, -l2)), dl * fast_ex2(fmaf(f.y, UB_LOG2E, -l2)));
    }
    const bool has_label = lab >= c && lab < c + 8;      // (softmax - 1) at the label column
    T fix = DT<T>::from_f(0.f);
    if (has_label)                                        // re-read the original logit (L1 hit)
      fix = DT<T>::from_f(dl * (fast_ex2(fmaf(DT<T>::to_f(x[lab]), UB_LOG2E, -l2)) - 1.0f));
    *reinterpret_cast<int4*>(x + c) = o.q;
    if (has_label) x[lab] = fix;                          // same thread, program order: lands last
  }
}

}  // namespace ub

extern "C" int ub200_cross_entropy_fwd(const void* logits, int64_t row_stride,
                                       const int64_t* labels, float* loss, float* lse,
                                       int64_t n_rows, int vocab, float softcap, float scale,
                                       int dtype, cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  const int vec_ok = (row_stride % V == 0) && (((uintptr_t)logits) % 16 == 0);
  // 512-thread CTAs, three resident per SM: while one CTA is in its end-of-row reduction the
  // others keep streaming (a single 1024-thread CTA per SM idled the SM at every row end)
  const int threads = vocab >= 8192 ? 512 : (vocab >= 2048 ? 256 : 128);
  int64_t g = (int64_t)UB_SM_COUNT * 6;
  const int grid = (int)(n_rows < g ? n_rows : g);
  if (dtype != UB200_F32 && softcap == 0.f && scale == 0.f && vec_ok && vocab >= 8192) {
    int64_t lg = (int64_t)UB_SM_COUNT * 2;               // one resident wave of 512-thread CTAs
    const int lgrid = (int)(n_rows < lg ? n_rows : lg);
    if (dtype == UB200_BF16)
      ce_fwd_lean_kernel<__nv_bfloat16><<<lgrid, 512, 0, stream>>>((const __nv_bfloat16*)logits, row_stride, labels, loss, lse, n_rows, vocab);
    else
      ce_fwd_lean_kernel<__half><<<lgrid, 512, 0, stream>>>((const __half*)logits, row_stride, labels, loss, lse, n_rows, vocab);
    UB_RETURN_LAST();
  }
#define GO(T)                                                                                  \
  ce_fwd_kernel<T><<<grid, threads, 0, stream>>>((const T*)logits, row_stride, labels, loss, lse, \
                                                 n_rows, vocab, softcap, scale, vec_ok)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GO(__half); }
  else if (dtype == UB200_F32) { GO(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
  UB_RETURN_LAST();
}

extern "C" int ub200_cross_entropy_bwd(void* logits, int64_t row_stride, const float* lse,
                                       const int64_t* labels, const float* dloss,
                                       int64_t dloss_stride, int64_t n_rows, int vocab,
                                       float softcap, float scale, int dtype,
                                       cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  if (n_rows > 2147483647LL) return UB200_ERR_UNSUPPORTED;
  const int V = dtype == UB200_F32 ? 4 : 8;
  const int vec_ok = (row_stride % V == 0) && (vocab % V == 0) && (((uintptr_t)logits) % 16 == 0);
  const int cols_per_block = 256 * V * 4;  // 4 vectors per thread
  const int chunks = (vocab + cols_per_block - 1) / cols_per_block;
  if (chunks > 65535) return UB200_ERR_UNSUPPORTED;
  dim3 grid((unsigned)n_rows, (unsigned)chunks);
  if (dtype != UB200_F32 && softcap == 0.f && scale == 0.f && vec_ok) {
    if (dtype == UB200_BF16)
      ce_bwd_lean_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)logits, row_stride, lse, labels, dloss, dloss_stride, vocab, cols_per_block);
    else
      ce_bwd_lean_kernel<__half><<<grid, 256, 0, stream>>>((__half*)logits, row_stride, lse, labels, dloss, dloss_stride, vocab, cols_per_block);
    UB_RETURN_LAST();
  }
#define GO(T)                                                                                  \
  ce_bwd_kernel<T><<<grid, 256, 0, stream>>>((T*)logits, row_stride, lse, labels, dloss,          \
                                             dloss_stride, vocab, softcap, scale, vec_ok,          \
                                             cols_per_block)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GO(__half); }
  else if (dtype == UB200_F32) { GO(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
  UB_RETURN_LAST();
}
