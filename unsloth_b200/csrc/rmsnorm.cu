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

template <typename T, bool GEMMA>
static int rms_fwd_t(const void* X, int64_t xs, const void* W, int wdt, void* Y, int64_t ys,
                     float* r, int64_t n_rows, int n_cols, float eps, cudaStream_t st) {
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
