// SwiGLU / GEGLU (tanh and erf) elementwise forward and in-place backward.
//
// Replaces the reference's Triton kernels
//   unsloth/kernels/swiglu.py:27-47, 67-109      (_fg_kernel, _DWf_DW_dfg_kernel)
//   unsloth/kernels/geglu.py:142-167, 188-244    (tanh approximation, Gemma / Gemma-2)
//   unsloth/kernels/geglu.py:31-53, 74-123       (exact erf form)
// HBM-bound streaming: 16-byte loads/stores, grid-stride, int64 indexing throughout.
// Algorithmic bytes per element: fwd 3*b (e,g -> h), bwd 6*b (DW,e,g -> h,df,de in place).
// Rounding points mirror the reference (f rounded to the tensor dtype before *g, all
// products rounded to the tensor dtype; de evaluated in fp32) -- SURVEY.md section 9.
#include <cstdlib>

#include "common.cuh"
#include "glu_math.cuh"

#ifndef UB200_GLU_DEFAULT_VARIANT
#define UB200_GLU_DEFAULT_VARIANT 4
#endif

namespace ub {

// Launch shape (measured, profiles/r2_glu_variants.log): the default is ONE-SHOT CTAs of 128 threads, one
// 16-byte vector of every operand per thread -- the grid covers the tensor (114,688 CTAs at cfg2) and the
// hardware block scheduler balances the tail; that is the shape of the reference's Triton launch and reaches
// the same 6.7 TB/s on the backward (0.211 ms, forward 0.115 ms = 6.1 TB/s).  The grid-stride form (variant 1:
// U vectors per thread per iteration, 32 CTAs per SM) ran 7 % behind: 6.05 rounds of the stride leave a 7th,
// 5 %-full round that one CTA in twenty pays for in full.
template <typename T, int ACT, int U, int MINB, int BS = 256, bool ONE = false, bool CS = true, bool STCS = false>
__global__ void __launch_bounds__(BS, MINB) glu_fwd_kernel(const T* __restrict__ e,
                                                     const T* __restrict__ g,
                                                     T* __restrict__ h, int64_t n_vec) {
  constexpr int V = DT<T>::VEC;
  // ONE: one-shot CTAs (the grid covers the tensor, the hardware block scheduler balances the tail)
  const int64_t step = ONE ? n_vec : (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n_vec; base += step) {
    int4 er[U], gr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n_vec) {
        er[u] = CS ? __ldcs(reinterpret_cast<const int4*>(e + i * V)) : __ldg(reinterpret_cast<const int4*>(e + i * V));
        gr[u] = CS ? __ldcs(reinterpret_cast<const int4*>(g + i * V)) : __ldg(reinterpret_cast<const int4*>(g + i * V));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i >= n_vec) break;
      const T* ev = reinterpret_cast<const T*>(&er[u]);
      const T* gv = reinterpret_cast<const T*>(&gr[u]);
      float o[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float f, d;
        act_eval<ACT, sizeof(T) == 2>(DT<T>::to_f(ev[k]), f, d);
        o[k] = DT<T>::rnd(f) * DT<T>::to_f(gv[k]);
      }
      if (STCS) store_vec_cs<T>(h + i * V, o); else store_vec<T>(h + i * V, o);
    }
  }
}

template <typename T, int ACT, int U, int MINB, int BS = 256, bool ONE = false, bool CS = true>
__global__ void __launch_bounds__(BS, MINB) glu_bwd_kernel(T* DW, T* e, T* g, int64_t n_vec) {
  constexpr int V = DT<T>::VEC;
  const int64_t step = ONE ? n_vec : (int64_t)gridDim.x * blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n_vec; base += step) {
    int4 dr[U], er[U], gr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n_vec) {
        dr[u] = CS ? __ldcs(reinterpret_cast<const int4*>(DW + i * V)) : *reinterpret_cast<const int4*>(DW + i * V);
        er[u] = CS ? __ldcs(reinterpret_cast<const int4*>(e + i * V)) : *reinterpret_cast<const int4*>(e + i * V);
        gr[u] = CS ? __ldcs(reinterpret_cast<const int4*>(g + i * V)) : *reinterpret_cast<const int4*>(g + i * V);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i >= n_vec) break;
      const T* dw = reinterpret_cast<const T*>(&dr[u]);
      const T* ev = reinterpret_cast<const T*>(&er[u]);
      const T* gv = reinterpret_cast<const T*>(&gr[u]);
      float oh[V], odf[V], ode[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float f, d;
        const float dwk = DT<T>::to_f(dw[k]), gk = DT<T>::to_f(gv[k]);
        act_eval<ACT, sizeof(T) == 2>(DT<T>::to_f(ev[k]), f, d);
        const float fr = DT<T>::rnd(f);
        oh[k] = fr * gk;                        // h  = f * g
        odf[k] = dwk * fr;                      // df = DW * f
        const float dg = DT<T>::rnd(dwk * gk);  // dg = DW * g (tensor dtype)
        ode[k] = dg * d;                        // de = dg.float() * df/de
      }
      store_vec<T>(DW + i * V, oh);
      store_vec<T>(e + i * V, odf);
      store_vec<T>(g + i * V, ode);
    }
  }
}

// UB200_GLU_VARIANT (numerics-neutral tuning probe): 4 (default) = one-shot CTAs, 1 = grid-stride, 7 = one-shot
// with streaming stores on the forward
static int glu_variant() {
  static const int v = [] { const char* e = getenv("UB200_GLU_VARIANT"); return e ? atoi(e) : UB200_GLU_DEFAULT_VARIANT; }();
  return v;
}

static inline int ew_grid(int64_t n_vec, int threads) {
  int64_t b = (n_vec + threads - 1) / threads;
  int64_t cap = (int64_t)UB_SM_COUNT * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace ub

extern "C" int ub200_glu_fwd(int act, const void* e, const void* g, void* h, int64_t n,
                             int dtype, cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  if (n % V) return UB200_ERR_BAD_ARG;
  const int64_t nv = n / V;
  const int var = glu_variant();
  const int grid = ew_grid(nv, 256 * 2);
  const unsigned one = (unsigned)((nv + 127) / 128);
#define GO(T, A)                                                                                                            \
  do {                                                                                                                      \
    if (var == 1) glu_fwd_kernel<T, A, 2, 6><<<grid, 256, 0, stream>>>((const T*)e, (const T*)g, (T*)h, nv);                  \
    else if (var == 7) glu_fwd_kernel<T, A, 1, 16, 128, true, true, true><<<one, 128, 0, stream>>>((const T*)e, (const T*)g, (T*)h, nv); \
    else glu_fwd_kernel<T, A, 1, 16, 128, true><<<one, 128, 0, stream>>>((const T*)e, (const T*)g, (T*)h, nv);                 \
  } while (0)
#define GOA(T)                                                  \
  if (act == ACT_SWIGLU) GO(T, ACT_SWIGLU);                     \
  else if (act == ACT_GEGLU_APPROX) GO(T, ACT_GEGLU_APPROX);    \
  else if (act == ACT_GEGLU_EXACT) GO(T, ACT_GEGLU_EXACT);      \
  else return UB200_ERR_BAD_ARG
  if (dtype == UB200_BF16) { GOA(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GOA(__half); }
  else if (dtype == UB200_F32) { GOA(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
#undef GOA
  UB_RETURN_LAST();
}

extern "C" int ub200_glu_bwd(int act, void* DW, void* e, void* g, int64_t n, int dtype,
                             cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  if (n % V) return UB200_ERR_BAD_ARG;
  const int64_t nv = n / V;
  const int var = glu_variant();
  const int grid = ew_grid(nv, 256);
  const unsigned one = (unsigned)((nv + 127) / 128);
#define GO(T, A)                                                                                           \
  do {                                                                                                     \
    if (var == 1) glu_bwd_kernel<T, A, 1, 5><<<grid, 256, 0, stream>>>((T*)DW, (T*)e, (T*)g, nv);            \
    else glu_bwd_kernel<T, A, 1, 10, 128, true><<<one, 128, 0, stream>>>((T*)DW, (T*)e, (T*)g, nv);          \
  } while (0)
#define GOA(T)                                                  \
  if (act == ACT_SWIGLU) GO(T, ACT_SWIGLU);                     \
  else if (act == ACT_GEGLU_APPROX) GO(T, ACT_GEGLU_APPROX);    \
  else if (act == ACT_GEGLU_EXACT) GO(T, ACT_GEGLU_EXACT);      \
  else return UB200_ERR_BAD_ARG
  if (dtype == UB200_BF16) { GOA(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GOA(__half); }
  else if (dtype == UB200_F32) { GOA(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
#undef GOA
  UB_RETURN_LAST();
}
