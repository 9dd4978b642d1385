// RMSNorm forward / backward for sm_100a.
//
// Replaces the Triton kernels of the reference:
//   unsloth/kernels/rms_layernorm.py:21-59   (_rms_layernorm_forward)
//   unsloth/kernels/rms_layernorm.py:123-159 (_gemma_rms_layernorm_forward)
//   unsloth/kernels/rms_layernorm.py:62-120  (_rms_layernorm_backward)
//
// HBM-bound streaming kernels.  One CTA walks rows grid-stride; a row is read ONCE into
// registers with 16-byte loads (VPT vectors per thread), reduced with warp shuffles +
// one smem exchange, and written once.  Algorithmic bytes per row:
//   fwd: H*(in+out bytes) + 4 (r)      bwd: 3*H*bytes
// Rounding points mirror the reference.
#include <type_traits>

#include "common.cuh"

namespace ub {

template <typename T, int VPT, bool GEMMA>
__global__ void __launch_bounds__(256) rms_fwd_kernel(
    const T* __restrict__ X, int64_t xs, const void* __restrict__ W, int wdt,
    T* __restrict__ Y, int64_t ys, float* __restrict__ r, int64_t n_rows, int n_cols, float eps) {
  constexpr int V = DT<T>::VEC;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  // weight slice for this thread is row-invariant: keep it in registers
  float w[VPT][V];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float wv = (c + i < n_cols) ? load_as_f(W, wdt, c + i) : 0.f;
      w[j][i] = GEMMA ? wv + 1.0f : wv;
    }
  }
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* x = X + row * xs;
    float xv[VPT][V];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * V;
      if (c < n_cols) {
        load_vec_cs<T>(x + c, xv[j]);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) xv[j][i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) ss += xv[j][i] * xv[j][i];
    }
    ss = block_sum(ss, red);
    const float inv = rsqrtf(ss / (float)n_cols + eps);
    if (tid == 0) r[row] = inv;
    T* y = Y + row * ys;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * V;
      if (c < n_cols) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float normed = xv[j][i] * inv;
          if (!GEMMA) normed = round_to(wdt, normed);  // `normed.to(W.dtype)` (:57)
          float prod = normed * w[j][i];
          if (!GEMMA) prod = round_to(wdt, prod);       // product is a W.dtype value
          o[i] = prod;
        }
        store_vec<T>(y + c, o);
      }
    }
  }
}

template <typename T, int VPT, bool GEMMA>
__global__ void __launch_bounds__(256) rms_bwd_kernel(
    const T* dY, int64_t dys, const T* __restrict__ X, int64_t xs,
    const void* __restrict__ W, int wdt, const float* __restrict__ r, T* dX,  // dX may alias dY
   
    int64_t dxs, int64_t n_rows, int n_cols) {
  constexpr int V = DT<T>::VEC;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  float w[VPT][V];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float wv = (c + i < n_cols) ? load_as_f(W, wdt, c + i) : 0.f;
      w[j][i] = GEMMA ? wv + 1.0f : wv;
    }
  }
  const float n = (float)n_cols;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* dy = dY + row * dys;
    const T* x = X + row * xs;
    const float inv = r[row];
    float dyw[VPT][V], nrm[VPT][V];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * V;
      if (c < n_cols) {
        float a[V], b[V];
        load_vec_cs<T>(dy + c, a);
        load_vec_cs<T>(x + c, b);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          dyw[j][i] = a[i] * w[j][i];
          nrm[j][i] = b[i] * inv;
          acc += dyw[j][i] * nrm[j][i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { dyw[j][i] = 0.f; nrm[j][i] = 0.f; }
      }
    }
    acc = block_sum(acc, red);
    const float k = inv / n;
    T* dx = dX + row * dxs;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * V;
      if (c < n_cols) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = k * (n * dyw[j][i] - nrm[j][i] * acc);
        store_vec<T>(dx + c, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lean 16-bit specialisations (the hot case: bf16/fp16 activations with weights of the same
// dtype).  The generic kernels above spend most of their issue slots emulating the reference's
// rounding points with scalar converts and a run-time dtype switch (ncu: 0.54 IPC on a kernel
// that should be waiting on HBM); here `normed.to(W.dtype) * W` is one packed HMUL2 per two
// elements (a bf16 x bf16 product rounded once -- exactly the reference's semantics), rows are
// held as raw 16-byte vectors, and the weight slice is preloaded once per CTA.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Pk;
template <> struct Pk<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ float2 up(T2 v) { return __bfloat1622float2(v); }
  __device__ static __forceinline__ T2 down(float a, float b) { return __floats2bfloat162_rn(a, b); }
};
template <> struct Pk<__half> {
  using T2 = __half2;
  __device__ static __forceinline__ float2 up(T2 v) { return __half22float2(v); }
  __device__ static __forceinline__ T2 down(float a, float b) { return __floats2half2_rn(a, b); }
};

template <typename T, int VPT>
__global__ void __launch_bounds__(256) rms_fwd_packed_kernel(
    const T* __restrict__ X, int64_t xs, const T* __restrict__ W, T* __restrict__ Y, int64_t ys,
    float* __restrict__ r, int64_t n_rows, int n_cols, float eps) {
  using T2 = typename Pk<T>::T2;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  union V16 { int4 q; T2 h[4]; };
  V16 w[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * 8;
    w[j].q = (c < n_cols) ? *reinterpret_cast<const int4*>(W + c) : make_int4(0, 0, 0, 0);
  }
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* x = X + row * xs;
    V16 xr[VPT];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      xr[j].q = (c < n_cols) ? __ldcs(reinterpret_cast<const int4*>(x + c)) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = Pk<T>::up(xr[j].h[i]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    ss = block_sum(ss, red);
    const float inv = rsqrtf(ss / (float)n_cols + eps);
    if (tid == 0) r[row] = inv;
    T* y = Y + row * ys;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        V16 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = Pk<T>::up(xr[j].h[i]);
          o.h[i] = __hmul2(Pk<T>::down(f.x * inv, f.y * inv), w[j].h[i]);   // normed.to(W.dtype) * W
        }
        *reinterpret_cast<int4*>(y + c) = o.q;
      }
    }
  }
}

// backward, 16-bit activations and weights of the same dtype: raw vectors in registers (no float
// copies of the row), weights preloaded, two passes over the registers.
template <typename T, int VPT, bool GEMMA>
__global__ void __launch_bounds__(256) rms_bwd_packed_kernel(
    const T* dY, int64_t dys, const T* __restrict__ X, int64_t xs, const T* __restrict__ W,
    const float* __restrict__ r, T* dX, int64_t dxs, int64_t n_rows, int n_cols) {
  using T2 = typename Pk<T>::T2;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  union V16 { int4 q; T2 h[4]; };
  V16 w[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * 8;
    w[j].q = (c < n_cols) ? *reinterpret_cast<const int4*>(W + c) : make_int4(0, 0, 0, 0);
  }
  const float n = (float)n_cols;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* dy = dY + row * dys;
    const T* x = X + row * xs;
    V16 a[VPT], b[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        a[j].q = __ldcs(reinterpret_cast<const int4*>(dy + c));
        b[j].q = __ldcs(reinterpret_cast<const int4*>(x + c));
      } else {
        a[j].q = make_int4(0, 0, 0, 0);
        b[j].q = make_int4(0, 0, 0, 0);
      }
    }
    const float inv = r[row];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fy = Pk<T>::up(a[j].h[i]), fx = Pk<T>::up(b[j].h[i]);
        float2 fw = Pk<T>::up(w[j].h[i]);
        if (GEMMA) { fw.x += 1.0f; fw.y += 1.0f; }
        acc = fmaf(fy.x * fw.x, fx.x * inv, acc);
        acc = fmaf(fy.y * fw.y, fx.y * inv, acc);
      }
    }
    acc = block_sum(acc, red);
    const float k = inv / n;
    T* dx = dX + row * dxs;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        V16 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fy = Pk<T>::up(a[j].h[i]), fx = Pk<T>::up(b[j].h[i]);
          float2 fw = Pk<T>::up(w[j].h[i]);
          if (GEMMA) { fw.x += 1.0f; fw.y += 1.0f; }
          const float o0 = k * (n * (fy.x * fw.x) - (fx.x * inv) * acc);
          const float o1 = k * (n * (fy.y * fw.y) - (fx.y * inv) * acc);
          o.h[i] = Pk<T>::down(o0, o1);
        }
        *reinterpret_cast<int4*>(dx + c) = o.q;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Residual add fused with the NEXT RMSNorm (Llama / Mistral decoder layers):
//   S = A + B (one bf16 rounding, exactly torch's bf16 add) ; Y = RMSNorm(S) ; r = rsqrt(...)
// and, for the backward, the norm gradient accumulated straight into the residual-stream gradient:
//   dS += rms_bwd(dY, S, W, r)
// The reference does these as separate passes (models/llama.py:838-844: `residual + hidden_states`
// then fast_rms_layernorm); fusing removes one full read+write of the residual stream per norm
// in each direction (4 instead of 5 / 6 tensor passes).  16-bit activations, same-dtype weights.
// ---------------------------------------------------------------------------------------------
template <typename T, int VPT>
__global__ void __launch_bounds__(256) add_rms_fwd_packed_kernel(
    const T* __restrict__ A, int64_t as, const T* __restrict__ B, int64_t bs,
    const T* __restrict__ W, T* __restrict__ S, int64_t ss_, T* __restrict__ Y, int64_t ys,
    float* __restrict__ r, int64_t n_rows, int n_cols, float eps) {
  using T2 = typename Pk<T>::T2;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  union V16 { int4 q; T2 h[4]; };
  V16 w[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * 8;
    w[j].q = (c < n_cols) ? *reinterpret_cast<const int4*>(W + c) : make_int4(0, 0, 0, 0);
  }
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* a = A + row * as;
    const T* b = B + row * bs;
    V16 xs[VPT];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        V16 va, vb;
        va.q = __ldcs(reinterpret_cast<const int4*>(a + c));
        vb.q = __ldcs(reinterpret_cast<const int4*>(b + c));
#pragma unroll
        for (int i = 0; i < 4; ++i) xs[j].h[i] = __hadd2(va.h[i], vb.h[i]);
        *reinterpret_cast<int4*>(S + row * ss_ + c) = xs[j].q;
      } else {
        xs[j].q = make_int4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = Pk<T>::up(xs[j].h[i]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
    ss = block_sum(ss, red);
    const float inv = rsqrtf(ss / (float)n_cols + eps);
    if (tid == 0) r[row] = inv;
    T* y = Y + row * ys;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        V16 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = Pk<T>::up(xs[j].h[i]);
          o.h[i] = __hmul2(Pk<T>::down(f.x * inv, f.y * inv), w[j].h[i]);
        }
        *reinterpret_cast<int4*>(y + c) = o.q;
      }
    }
  }
}

template <typename T, int VPT>
__global__ void __launch_bounds__(256) rms_bwd_acc_packed_kernel(
    const T* __restrict__ dY, int64_t dys, const T* __restrict__ X, int64_t xs,
    const T* __restrict__ W, const float* __restrict__ r, T* dS, int64_t dss, int64_t n_rows,
    int n_cols) {
  using T2 = typename Pk<T>::T2;
  __shared__ float red[32];
  const int tid = threadIdx.x;
  union V16 { int4 q; T2 h[4]; };
  V16 w[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = (j * blockDim.x + tid) * 8;
    w[j].q = (c < n_cols) ? *reinterpret_cast<const int4*>(W + c) : make_int4(0, 0, 0, 0);
  }
  const float n = (float)n_cols;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* dy = dY + row * dys;
    const T* x = X + row * xs;
    T* ds = dS + row * dss;
    V16 a[VPT], b[VPT], g[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        a[j].q = __ldcs(reinterpret_cast<const int4*>(dy + c));
        b[j].q = __ldcs(reinterpret_cast<const int4*>(x + c));
        g[j].q = *reinterpret_cast<const int4*>(ds + c);
      } else {
        a[j].q = b[j].q = g[j].q = make_int4(0, 0, 0, 0);
      }
    }
    const float inv = r[row];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fy = Pk<T>::up(a[j].h[i]), fx = Pk<T>::up(b[j].h[i]), fw = Pk<T>::up(w[j].h[i]);
        acc = fmaf(fy.x * fw.x, fx.x * inv, acc);
        acc = fmaf(fy.y * fw.y, fx.y * inv, acc);
      }
    }
    acc = block_sum(acc, red);
    const float k = inv / n;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * blockDim.x + tid) * 8;
      if (c < n_cols) {
        V16 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fy = Pk<T>::up(a[j].h[i]), fx = Pk<T>::up(b[j].h[i]), fw = Pk<T>::up(w[j].h[i]);
          const float2 fg = Pk<T>::up(g[j].h[i]);
          o.h[i] = Pk<T>::down(fg.x + k * (n * (fy.x * fw.x) - (fx.x * inv) * acc),
                               fg.y + k * (n * (fy.y * fw.y) - (fx.y * inv) * acc));
        }
        *reinterpret_cast<int4*>(ds + c) = o.q;
      }
    }
  }
}

template <typename T, bool GEMMA, typename F>
static int dispatch_vpt(int n_cols, F&& launch) {
  constexpr int V = DT<T>::VEC;
  // threads chosen so that VPT in {1,2,4,8} covers n_cols
  int threads = 256;
  if (n_cols <= 128 * V) threads = 128;
  int vpt = (n_cols + threads * V - 1) / (threads * V);
  if (vpt <= 1) return launch(std::integral_constant<int, 1>{}, threads);
  if (vpt <= 2) return launch(std::integral_constant<int, 2>{}, threads);
  if (vpt <= 4) return launch(std::integral_constant<int, 4>{}, threads);
  if (vpt <= 8) return launch(std::integral_constant<int, 8>{}, threads);
  return UB200_ERR_UNSUPPORTED;
}


static inline int rows_grid(int64_t n_rows) {
  int64_t g = (int64_t)UB_SM_COUNT * 8;
  return (int)(n_rows < g ? n_rows : g);
}

template <typename T> struct is16 { static constexpr bool v = false; static constexpr int code = UB200_F32; };
template <> struct is16<__nv_bfloat16> { static constexpr bool v = true; static constexpr int code = UB200_BF16; };
template <> struct is16<__half> { static constexpr bool v = true; static constexpr int code = UB200_F16; };

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, bool GEMMA>
static int rms_fwd_t(const void* X, int64_t xs, const void* W, int wdt, void* Y, int64_t ys,
                     float* r, int64_t n_rows, int n_cols, float eps, cudaStream_t st) {
  if constexpr (is16<T>::v && !GEMMA) {
    if (wdt == is16<T>::code && aligned16(W) && aligned16(X) && aligned16(Y)) {
      return dispatch_vpt<T, GEMMA>(n_cols, [&](auto vpt, int threads) {
        static int occ = 0;   // resident CTAs per SM for this instantiation: size the grid to one wave
        if (!occ) {
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rms_fwd_packed_kernel<T, decltype(vpt)::value>, threads, 0);
          if (occ < 1) occ = 1;
        }
        const int64_t cap = (int64_t)UB_SM_COUNT * occ;
        rms_fwd_packed_kernel<T, decltype(vpt)::value><<<(int)(n_rows < cap ? n_rows : cap), threads, 0, st>>>(
            (const T*)X, xs, (const T*)W, (T*)Y, ys, r, n_rows, n_cols, eps);
        return UB200_OK;
      });
    }
  }
  return dispatch_vpt<T, GEMMA>(n_cols, [&](auto vpt, int threads) {
    rms_fwd_kernel<T, decltype(vpt)::value, GEMMA><<<rows_grid(n_rows), threads, 0, st>>>(
        (const T*)X, xs, W, wdt, (T*)Y, ys, r, n_rows, n_cols, eps);
    return UB200_OK;
  });
}
template <typename T, bool GEMMA>
static int rms_bwd_t(const void* dY, int64_t dys, const void* X, int64_t xs, const void* W,
                     int wdt, const float* r, void* dX, int64_t dxs, int64_t n_rows, int n_cols,
                     cudaStream_t st) {
  if constexpr (is16<T>::v) {
    if (wdt == is16<T>::code && aligned16(W) && aligned16(X) && aligned16(dY) && aligned16(dX)) {
      return dispatch_vpt<T, GEMMA>(n_cols, [&](auto vpt, int threads) {
        static int occ = 0;
        if (!occ) {
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rms_bwd_packed_kernel<T, decltype(vpt)::value, GEMMA>, threads, 0);
          if (occ < 1) occ = 1;
        }
        const int64_t cap = (int64_t)UB_SM_COUNT * occ;
        rms_bwd_packed_kernel<T, decltype(vpt)::value, GEMMA><<<(int)(n_rows < cap ? n_rows : cap), threads, 0, st>>>(
            (const T*)dY, dys, (const T*)X, xs, (const T*)W, r, (T*)dX, dxs, n_rows, n_cols);
        return UB200_OK;
      });
    }
  }
  return dispatch_vpt<T, GEMMA>(n_cols, [&](auto vpt, int threads) {
    rms_bwd_kernel<T, decltype(vpt)::value, GEMMA><<<rows_grid(n_rows), threads, 0, st>>>(
        (const T*)dY, dys, (const T*)X, xs, W, wdt, r, (T*)dX, dxs, n_rows, n_cols);
    return UB200_OK;
  });
}

}  // namespace ub

extern "C" int ub200_rms_layernorm_fwd(const void* X, int64_t x_row_stride, const void* W,
                                       int w_dtype, void* Y, int64_t y_row_stride, float* r,
                                       int64_t n_rows, int n_cols, float eps, int gemma,
                                       int dtype, cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  if (n_cols % V || x_row_stride % V || y_row_stride % V) return UB200_ERR_BAD_ARG;
  int rc;
#define GO(T)                                                                                   \
  rc = gemma ? rms_fwd_t<T, true>(X, x_row_stride, W, w_dtype, Y, y_row_stride, r, n_rows,      \
                                  n_cols, eps, stream)                                          \
             : rms_fwd_t<T, false>(X, x_row_stride, W, w_dtype, Y, y_row_stride, r, n_rows,     \
                                   n_cols, eps, stream)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GO(__half); }
  else if (dtype == UB200_F32) { GO(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
  if (rc) return rc;
  UB_RETURN_LAST();
}

extern "C" int ub200_rms_layernorm_bwd(const void* dY, int64_t dy_row_stride, const void* X,
                                       int64_t x_row_stride, const void* W, int w_dtype,
                                       const float* r, void* dX, int64_t dx_row_stride,
                                       int64_t n_rows, int n_cols, int gemma, int dtype,
                                       cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  const int V = dtype == UB200_F32 ? 4 : 8;
  if (n_cols % V || x_row_stride % V || dy_row_stride % V || dx_row_stride % V)
    return UB200_ERR_BAD_ARG;
  int rc;
#define GO(T)                                                                                   \
  rc = gemma ? rms_bwd_t<T, true>(dY, dy_row_stride, X, x_row_stride, W, w_dtype, r, dX,        \
                                  dx_row_stride, n_rows, n_cols, stream)                        \
             : rms_bwd_t<T, false>(dY, dy_row_stride, X, x_row_stride, W, w_dtype, r, dX,       \
                                   dx_row_stride, n_rows, n_cols, stream)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); }
  else if (dtype == UB200_F16) { GO(__half); }
  else if (dtype == UB200_F32) { GO(float); }
  else return UB200_ERR_BAD_ARG;
#undef GO
  if (rc) return rc;
  UB_RETURN_LAST();
}

extern "C" int ub200_add_rms_layernorm_fwd(const void* A, int64_t a_row_stride, const void* B,
                                           int64_t b_row_stride, const void* W, void* S,
                                           int64_t s_row_stride, void* Y, int64_t y_row_stride,
                                           float* r, int64_t n_rows, int n_cols, float eps, int dtype,
                                           cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_UNSUPPORTED;
  if (n_cols % 8 || a_row_stride % 8 || b_row_stride % 8 || s_row_stride % 8 || y_row_stride % 8)
    return UB200_ERR_BAD_ARG;
  if (!aligned16(A) || !aligned16(B) || !aligned16(W) || !aligned16(S) || !aligned16(Y)) return UB200_ERR_BAD_ARG;
  int rc;
#define GO(T)                                                                                        \
  rc = dispatch_vpt<T, false>(n_cols, [&](auto vpt, int threads) {                                   \
    static int occ = 0;                                                                              \
    if (!occ) {                                                                                      \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                                 \
          &occ, add_rms_fwd_packed_kernel<T, decltype(vpt)::value>, threads, 0);                     \
      if (occ < 1) occ = 1;                                                                          \
    }                                                                                                \
    const int64_t cap = (int64_t)UB_SM_COUNT * occ;                                                  \
    add_rms_fwd_packed_kernel<T, decltype(vpt)::value>                                               \
        <<<(int)(n_rows < cap ? n_rows : cap), threads, 0, stream>>>(                                \
            (const T*)A, a_row_stride, (const T*)B, b_row_stride, (const T*)W, (T*)S, s_row_stride,  \
            (T*)Y, y_row_stride, r, n_rows, n_cols, eps);                                            \
    return UB200_OK;                                                                                 \
  })
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); } else { GO(__half); }
#undef GO
  if (rc) return rc;
  UB_RETURN_LAST();
}

extern "C" int ub200_rms_layernorm_bwd_acc(const void* dY, int64_t dy_row_stride, const void* X,
                                           int64_t x_row_stride, const void* W, const float* r,
                                           void* dS, int64_t ds_row_stride, int64_t n_rows,
                                           int n_cols, int dtype, cudaStream_t stream) {
  using namespace ub;
  if (n_rows <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_UNSUPPORTED;
  if (n_cols % 8 || dy_row_stride % 8 || x_row_stride % 8 || ds_row_stride % 8) return UB200_ERR_BAD_ARG;
  if (!aligned16(dY) || !aligned16(X) || !aligned16(W) || !aligned16(dS)) return UB200_ERR_BAD_ARG;
  int rc;
#define GO(T)                                                                                        \
  rc = dispatch_vpt<T, false>(n_cols, [&](auto vpt, int threads) {                                   \
    static int occ = 0;                                                                              \
    if (!occ) {                                                                                      \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                                 \
          &occ, rms_bwd_acc_packed_kernel<T, decltype(vpt)::value>, threads, 0);                     \
      if (occ < 1) occ = 1;                                                                          \
    }                                                                                                \
    const int64_t cap = (int64_t)UB_SM_COUNT * occ;                                                  \
    rms_bwd_acc_packed_kernel<T, decltype(vpt)::value>                                               \
        <<<(int)(n_rows < cap ? n_rows : cap), threads, 0, stream>>>(                                \
            (const T*)dY, dy_row_stride, (const T*)X, x_row_stride, (const T*)W, r, (T*)dS,          \
            ds_row_stride, n_rows, n_cols);                                                          \
    return UB200_OK;                                                                                 \
  })
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); } else { GO(__half); }
#undef GO
  if (rc) return rc;
  UB_RETURN_LAST();
}
