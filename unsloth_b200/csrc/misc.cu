// Small helpers of the LoRA path: dtype cast with zero padding (+ optional transpose), and a
// flat AdamW over the contiguous LoRA parameter bucket.
//
//  * ub200_cast_pad_2d replaces the per-call `A.to(dtype)`, `B.to(dtype)` casts and the
//    `alpha = s` scaling of unsloth/kernels/utils.py:1163-1167 and fast_lora.py:138-153; the
//    padding brings the rank dimension up to the 64-wide K block of the tcgen05 GEMM.
//  * ub200_adamw_flat: the reference leaves the optimiser to torch (AdamW over 448 LoRA
//    tensors); here the LoRA parameters live in ONE flat fp32 bucket (shared with the DDP
//    all-reduce) so the update is a single streaming launch.  HBM-bound: 28 B / parameter.
#include "common.cuh"

namespace ub {

__global__ void __launch_bounds__(256) cast_pad_kernel(const void* __restrict__ src, int sdt,
                                                       int64_t sld, int rows, int cols,
                                                       void* __restrict__ dst, int ddt, int64_t dld,
                                                       int drows, int dcols, int roff,
                                                       int coff, float scale, int transpose) {
  const int64_t total = (int64_t)drows * dcols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int dr = (int)(i / dcols), dc = (int)(i - (int64_t)dr * dcols);
    const int lr = dr - roff, lc = dc - coff;   // position inside the placed block
    const int sr = transpose ? lc : lr, sc = transpose ? lr : lc;
    float v = 0.f;
    if (sr >= 0 && sc >= 0 && sr < rows && sc < cols) v = scale * load_as_f(src, sdt, (int64_t)sr * sld + sc);
    const int64_t o = (int64_t)dr * dld + dc;
    if (ddt == UB200_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
    else if (ddt == UB200_F16) reinterpret_cast<__half*>(dst)[o] = __float2half_rn(v);
    else reinterpret_cast<float*>(dst)[o] = v;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    int64_t n4, int64_t n, float lr, float b1,
                                                    float b2, float eps, float wd, float bc1,
                                                    float bc2, float gs) {
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gs;
    pp *= (1.0f - lr * wd);
    mm = b1 * mm + (1.0f - b1) * gg;
    vv = b2 * vv + (1.0f - b2) * gg * gg;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp -= step * (mm / denom);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  // tail
  const int64_t t0 = n4 * 4;
  for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    upd(p[i], g[i], m[i], v[i]);
}

// ---- batched forms: one launch for MANY small tensors ------------------------------------------
// A QLoRA step refreshes 448 LoRA casts and (with autograd) accumulates 448 LoRA gradients, each a
// 4-5 us launch moving a few hundred KB: 1.8 % of the step as ~900 launches (profiles/
// r2_launches_bench_full_summary.txt).  The descriptors travel BY VALUE in the kernel parameters
// (a chunk of 40 per launch, < 4 KB), so nothing has to be staged on the device and the launches
// are CUDA-graph capturable; blockIdx.y = descriptor, blockIdx.x strides over its elements.
constexpr int MULTI_CHUNK = 40;
// CTAs per descriptor.  8 (first cut) left the batched casts at ~40 G elements/s -- 3.7 ms of a cfg2 step for
// 0.45 GB of traffic (profiles/r2_launches_bench_full_summary.txt): 2048 threads per descriptor, a 64-bit
// division and a 2-byte store per element.  Now 32 CTAs, 32-bit index arithmetic and, where the destination
// allows it (16-bit, 8-column granularity), one 16-byte store per thread per iteration.
constexpr int MULTI_CTAS = 32;
struct CastBatch { ub200_cast_desc d[MULTI_CHUNK]; };
struct AccBatch { ub200_acc_desc d[MULTI_CHUNK]; };

// value of destination element (dr, dc): scale * src at the placed block, zero in the padding
__device__ __forceinline__ float cast_pad_value(const ub200_cast_desc& d, int dr, int dc) {
  const int lr = dr - d.row_off, lc = dc - d.col_off;
  const int sr = d.transpose ? lc : lr, sc = d.transpose ? lr : lc;
  if (sr >= 0 && sc >= 0 && sr < d.rows && sc < d.cols)
    return d.scale * load_as_f(d.src, d.src_dtype, (int64_t)sr * d.src_ld + sc);
  return 0.f;
}

__global__ void __launch_bounds__(256) cast_pad_multi_kernel(const __grid_constant__ CastBatch b) {
  const ub200_cast_desc& d = b.d[blockIdx.y];
  const int64_t total = (int64_t)d.dst_rows * d.dst_cols;
  const bool vec = d.dst_dtype != UB200_F32 && (d.dst_cols % 8) == 0 && (d.dst_ld % 8) == 0 &&
                   (reinterpret_cast<uintptr_t>(d.dst) & 15) == 0 && total < (1ll << 31);
  if (vec) {
    const uint32_t cv = (uint32_t)d.dst_cols / 8u, nv = (uint32_t)(total / 8);
    const bool bf = d.dst_dtype == UB200_BF16;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
      const uint32_t dr = i / cv, dc0 = (i - dr * cv) * 8u;
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = cast_pad_value(d, (int)dr, (int)dc0 + 2 * k);
        const float hi = cast_pad_value(d, (int)dr, (int)dc0 + 2 * k + 1);
        if (bf) {
          const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
          w[k] = *reinterpret_cast<const uint32_t*>(&t);
        } else {
          const __half2 t = __floats2half2_rn(lo, hi);
          w[k] = *reinterpret_cast<const uint32_t*>(&t);
        }
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(d.dst) + (int64_t)dr * d.dst_ld + dc0) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int dr = (int)(i / d.dst_cols), dc = (int)(i - (int64_t)dr * d.dst_cols);
    const float v = cast_pad_value(d, dr, dc);
    const int64_t o = (int64_t)dr * d.dst_ld + dc;
    if (d.dst_dtype == UB200_BF16) reinterpret_cast<__nv_bfloat16*>(d.dst)[o] = __float2bfloat16_rn(v);
    else if (d.dst_dtype == UB200_F16) reinterpret_cast<__half*>(d.dst)[o] = __float2half_rn(v);
    else reinterpret_cast<float*>(d.dst)[o] = v;
  }
}

// dst[r, c] (contiguous fp32) += src[r * src_rs + c * src_cs]
template <typename I>
__device__ __forceinline__ void accumulate_one(const ub200_acc_desc& d, I total) {
  // walk the SOURCE's fast axis with consecutive threads when it is the transposed one
  const bool src_row_fast = d.src_rs == 1 && d.src_cs != 1;
  const I rows = (I)d.rows, cols = (I)d.cols;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    I r, c;
    if (src_row_fast) { c = i / rows; r = i - c * rows; }
    else { r = i / cols; c = i - r * cols; }
    d.dst[(int64_t)r * d.cols + (int64_t)c] += d.src[(int64_t)r * d.src_rs + (int64_t)c * d.src_cs];
  }
}

__global__ void __launch_bounds__(256) accumulate_multi_kernel(const __grid_constant__ AccBatch b) {
  const ub200_acc_desc& d = b.d[blockIdx.y];
  const int64_t total = (int64_t)d.rows * d.cols;
  if (total < (1ll << 31)) accumulate_one<uint32_t>(d, (uint32_t)total);
  else accumulate_one<int64_t>(d, total);
}

}  // namespace ub

extern "C" int ub200_abi_version(void) { return 1; }

extern "C" int ub200_cast_pad_multi(const ub200_cast_desc* descs, int n, cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  if (!descs) return UB200_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n; i0 += MULTI_CHUNK) {
    CastBatch b;
    const int m = n - i0 < MULTI_CHUNK ? n - i0 : MULTI_CHUNK;
    for (int i = 0; i < m; ++i) {
      b.d[i] = descs[i0 + i];
      if (!b.d[i].src || !b.d[i].dst || b.d[i].dst_rows <= 0 || b.d[i].dst_cols <= 0) return UB200_ERR_BAD_ARG;
    }
    cast_pad_multi_kernel<<<dim3(MULTI_CTAS, m), 256, 0, stream>>>(b);
  }
  UB_RETURN_LAST();
}

extern "C" int ub200_accumulate_multi(const ub200_acc_desc* descs, int n, cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  if (!descs) return UB200_ERR_BAD_ARG;
  for (int i0 = 0; i0 < n; i0 += MULTI_CHUNK) {
    AccBatch b;
    const int m = n - i0 < MULTI_CHUNK ? n - i0 : MULTI_CHUNK;
    for (int i = 0; i < m; ++i) {
      b.d[i] = descs[i0 + i];
      if (!b.d[i].src || !b.d[i].dst || b.d[i].rows <= 0 || b.d[i].cols <= 0) return UB200_ERR_BAD_ARG;
    }
    accumulate_multi_kernel<<<dim3(MULTI_CTAS, m), 256, 0, stream>>>(b);
  }
  UB_RETURN_LAST();
}

extern "C" int ub200_cast_pad_2d(const void* src, int src_dtype, int64_t src_ld, int rows, int cols,
                                 void* dst, int dst_dtype, int64_t dst_ld, int dst_rows,
                                 int dst_cols, int dst_row_off, int dst_col_off, float scale,
                                 int transpose, cudaStream_t stream) {
  using namespace ub;
  if (dst_rows <= 0 || dst_cols <= 0) return UB200_OK;
  const int64_t total = (int64_t)dst_rows * dst_cols;
  int64_t b = (total + 255) / 256;
  if (b > UB_SM_COUNT * 8) b = UB_SM_COUNT * 8;
  cast_pad_kernel<<<(int)b, 256, 0, stream>>>(src, src_dtype, src_ld, rows, cols, dst, dst_dtype,
                                              dst_ld, dst_rows, dst_cols, dst_row_off, dst_col_off, scale,
                                              transpose);
  UB_RETURN_LAST();
}

extern "C" int ub200_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay,
                                float bias_corr1, float bias_corr2, float grad_scale,
                                cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  int64_t b = ((n4 > 0 ? n4 : n) + 255) / 256;
  if (b > UB_SM_COUNT * 8) b = UB_SM_COUNT * 8;
  adamw_kernel<<<(int)b, 256, 0, stream>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay,
                                           bias_corr1, bias_corr2, grad_scale);
  UB_RETURN_LAST();
}
