// Decode-time matrix-vector products: y[m] = W[m,k] . x[k]
//
//  * ub200_gemv_nf4 / cgemm_4bit_inference_naive_{bf16,fp16}: W is NF4-packed, expanded in
//    registers, never written to HBM.  Replaces the bitsandbytes launch behind the reference's
//    `fast_gemv` (unsloth/kernels/utils.py:874-973) and the q_len == 1 branch of
//    `fast_linear_forward` (:1082-1125); the two C symbols are the ones the reference binds through
//    ctypes (:283-284).  ub200_gemv_nf4 additionally folds the double-quantised absmax
//    reconstruction (the reference's separate cdequantize_blockwise_fp32 launch + `+= offset`,
//    :938-948) and the LoRA epilogue `+ s * B[row,:r] . t[:r]` (:1108-1112) into the same launch.
//  * ub200_gemv_dense: 16-bit dense rows (the `torch.mv(lm_head, h)` of models/llama.py:1460 and
//    the `A x` LoRA temp).
//
// bitsandbytes is not vendored in the reference: the arithmetic here (fp32 code * fp32 absmax,
// fp32 accumulation over k) restates the published NF4 layout and is at least as accurate as the
// 16-bit products of the original; parity against bitsandbytes itself is unpinned.
//
// HBM-bound by bytes.  Algorithmic bytes per call: m*k*(0.5 + 1/blocksize) + k*2 + m*2 (NF4);
// m*k*2 + k*2 + m*out_bytes (dense).  One warp owns a row (NF4) or two rows (dense) and walks k
// in 1024- / 256-column steps, 16 bytes per lane per step.
#include "common.cuh"

namespace ub {

__constant__ float kNF4g[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

template <typename T> struct P2;
template <> struct P2<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  // bf16 is the upper half of an fp32: one shift / one mask per element
  __device__ static __forceinline__ float2 up(T2 v) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&v);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
  }
  __device__ static __forceinline__ __nv_bfloat16 down(float v) { return __float2bfloat16_rn(v); }
};
template <> struct P2<__half> {
  using T2 = __half2;
  __device__ static __forceinline__ float2 up(T2 v) { return __half22float2(v); }
  __device__ static __forceinline__ __half down(float v) { return __float2half_rn(v); }
};

// One row per warp; the 16-entry fp32 code table is replicated per LANE in shared memory (2 KB:
// entry n of lane l at n*32 + l), so every lookup of a warp is conflict-free and costs one
// wavefront whatever the nibbles are; 40 registers per thread => 48 resident warps per SM.
// Round-1 status (profiles/r1_gemv_nf4_ncu.txt, profiles/r1_kernel_bench_gemv_addrms.log): this
// kernel and a 32 KB byte-pair-table variant with 4 rows per warp both land at 24-28 us for a
// 14336 x 4096 weight (1.1-1.3 TB/s): ~6 issued instructions per weight (shift, mask, LDS, FFMA,
// x unpack) at the power-capped ~1.4 GHz SM clock is an ISSUE bound of ~14 us plus a 2.02-wave
// tail, not a DRAM bound.  Reaching the HBM roofline needs <= 2 instructions per weight, i.e. a
// 16-bit pair table feeding HFMA2 (an `mma.sync` fragment variant was built and measured in round 2:
// not faster, profiles/r2_gemv_bench.log, removed) -- DESIGN.md section 8.
template <typename T>
__global__ void __launch_bounds__(256) gemv_nf4_lite_kernel(
    const T* __restrict__ x, const uint8_t* __restrict__ packed,
    const float* __restrict__ absmax_f32, const uint8_t* __restrict__ absmax_q,
    const float* __restrict__ code2, const float* __restrict__ absmax2,
    const float* __restrict__ offset, const float* __restrict__ code16, T* __restrict__ out, int m,
    int k, int bs_shift, int bs2_shift, const T* __restrict__ lora_B, int ldb,
    const float* __restrict__ lora_t, int r, float s) {
  using T2 = typename P2<T>::T2;
  __shared__ float lut[16 * 32];
  for (int e = threadIdx.x; e < 16 * 32; e += 256) lut[e] = code16 ? code16[e >> 5] : kNF4g[e >> 5];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= m) return;
  const float off = offset ? *offset : 0.f;
  const uint8_t* wrow = packed + (int64_t)row * (k / 2);
  const int64_t ebase = (int64_t)row * k;
  const uint32_t lane_bits = (uint32_t)lane << 2;
  const char* lut_bytes = reinterpret_cast<const char*>(lut);
  auto scale = [&](int c0) {
    const int64_t blk = (ebase + c0) >> bs_shift;
    return absmax_f32 ? absmax_f32[blk] : fmaf(code2[absmax_q[blk]], absmax2[blk >> bs2_shift], off);
  };
  float acc = 0.f;
  int c0 = lane * 32;
  int4 wc = make_int4(0, 0, 0, 0), wn;
  float amc = 0.f, amn;
  if (c0 < k) { wc = __ldcs(reinterpret_cast<const int4*>(wrow + c0 / 2)); amc = scale(c0); }
#pragma unroll 1
  for (; c0 < k; c0 += 1024) {
    const bool more = c0 + 1024 < k;
    if (more) { wn = __ldcs(reinterpret_cast<const int4*>(wrow + (c0 + 1024) / 2)); amn = scale(c0 + 1024); }
    const uint32_t u[4] = {(uint32_t)wc.x, (uint32_t)wc.y, (uint32_t)wc.z, (uint32_t)wc.w};
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      union { int4 v; T2 h[4]; } xv;
      xv.v = __ldg(reinterpret_cast<const int4*>(x + c0) + q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = P2<T>::up(xv.h[j]);
        // byte j of the word: high nibble = even element; nibble moved to bits 7..10 (n * 128)
        const int sh_hi = 8 * j + 4 - 7, sh_lo = 8 * j - 7;
        const uint32_t o_hi = ((sh_hi < 0 ? u[q] << -sh_hi : u[q] >> sh_hi) & 0x780u) | lane_bits;
        const uint32_t o_lo = ((sh_lo < 0 ? u[q] << -sh_lo : u[q] >> sh_lo) & 0x780u) | lane_bits;
        p0 = fmaf(*reinterpret_cast<const float*>(lut_bytes + o_hi), f.x, p0);
        p1 = fmaf(*reinterpret_cast<const float*>(lut_bytes + o_lo), f.y, p1);
      }
    }
    acc = fmaf(amc, p0 + p1, acc);
    if (more) { wc = wn; amc = amn; }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc;
    if (lora_B) {
      float d = 0.f;
      for (int j = 0; j < r; ++j) d = fmaf(DT<T>::to_f(lora_B[(int64_t)row * ldb + j]), lora_t[j], d);
      v = fmaf(s, d, v);
    }
    out[row] = P2<T>::down(v);
  }
}

template <typename T, typename O, int ROWS>
__global__ void __launch_bounds__(256) gemv_dense_kernel(const T* __restrict__ x,
                                                         const T* __restrict__ W, int64_t ldw,
                                                         O* __restrict__ out, int m, int k) {
  using T2 = typename P2<T>::T2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = (blockIdx.x * 8 + warp) * ROWS;
  if (row0 >= m) return;
  float acc[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) acc[i] = 0.f;
  for (int c0 = lane * 8; c0 < k; c0 += 256) {
    union V { int4 v; T2 h[4]; };
    V xv, wv[ROWS];
    xv.v = __ldg(reinterpret_cast<const int4*>(x + c0));
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int row = row0 + i < m ? row0 + i : m - 1;
      wv[i].v = __ldcs(reinterpret_cast<const int4*>(W + row * ldw + c0));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fx = P2<T>::up(xv.h[j]);
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const float2 fw = P2<T>::up(wv[i].h[j]);
        acc[i] = fmaf(fw.x, fx.x, acc[i]);
        acc[i] = fmaf(fw.y, fx.y, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < ROWS; ++i)
      if (row0 + i < m) {
        if constexpr (sizeof(O) == 4) out[row0 + i] = acc[i];
        else out[row0 + i] = P2<T>::down(acc[i]);
      }
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T>
static int launch_gemv_nf4(const void* x, const uint8_t* packed, const float* absmax_f32,
                           const uint8_t* absmax_q, const float* code2, const float* absmax2,
                           const float* offset, const float* code16, void* out, int m, int k,
                           int blocksize, int blocksize2, const void* lora_B, int ldb,
                           const float* lora_t, int r, float s, cudaStream_t st) {
  int bs_shift = 0, bs2_shift = 0;
  while ((1 << bs_shift) < blocksize) ++bs_shift;
  while ((1 << bs2_shift) < blocksize2) ++bs2_shift;
  gemv_nf4_lite_kernel<T><<<(m + 7) / 8, 256, 0, st>>>(
      (const T*)x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16, (T*)out, m, k,
      bs_shift, bs2_shift, (const T*)lora_B, ldb, lora_t, r, s);
  UB_RETURN_LAST();
}

}  // namespace ub

extern "C" int ub200_gemv_nf4(const void* x, const uint8_t* packed, const float* absmax_f32,
                              const uint8_t* absmax_q, const float* code2, const float* absmax2,
                              const float* offset, const float* code16, void* out, int m, int k,
                              int blocksize, int blocksize2, const void* lora_B, int ldb,
                              const float* lora_t, int r, float s, int dtype,
                              cudaStream_t stream) {
  using namespace ub;
  if (m <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_UNSUPPORTED;
  if (k <= 0 || k % 32 || blocksize < 32 || (blocksize & (blocksize - 1))) return UB200_ERR_UNSUPPORTED;
  if (!absmax_f32 && (!absmax_q || !code2 || !absmax2 || blocksize2 <= 0)) return UB200_ERR_BAD_ARG;
  if (!absmax_f32 && (blocksize2 & (blocksize2 - 1))) return UB200_ERR_UNSUPPORTED;
  if (lora_B && (!lora_t || r <= 0 || ldb < r)) return UB200_ERR_BAD_ARG;
  if (!x || !packed || !out || !al16(x) || !al16(packed)) return UB200_ERR_BAD_ARG;
  if (dtype == UB200_BF16)
    return launch_gemv_nf4<__nv_bfloat16>(x, packed, absmax_f32, absmax_q, code2, absmax2, offset,
                                          code16, out, m, k, blocksize, blocksize2, lora_B, ldb,
                                          lora_t, r, s, stream);
  return launch_gemv_nf4<__half>(x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16,
                                 out, m, k, blocksize, blocksize2, lora_B, ldb, lora_t, r, s, stream);
}

extern "C" void cgemm_4bit_inference_naive_bf16(int m, int n, int k, __nv_bfloat16* A,
                                                unsigned char* B, float* absmax, float* datatype,
                                                __nv_bfloat16* out, int lda, int ldb, int ldc,
                                                int blocksize, cudaStream_t stream) {
  (void)n; (void)lda; (void)ldb; (void)ldc;
  ub200_gemv_nf4(A, B, absmax, nullptr, nullptr, nullptr, nullptr, datatype, out, m, k, blocksize,
                 0, nullptr, 0, nullptr, 0, 0.f, UB200_BF16, stream);
}
extern "C" void cgemm_4bit_inference_naive_fp16(int m, int n, int k, __half* A, unsigned char* B,
                                                float* absmax, float* datatype, __half* out,
                                                int lda, int ldb, int ldc, int blocksize,
                                                cudaStream_t stream) {
  (void)n; (void)lda; (void)ldb; (void)ldc;
  ub200_gemv_nf4(A, B, absmax, nullptr, nullptr, nullptr, nullptr, datatype, out, m, k, blocksize,
                 0, nullptr, 0, nullptr, 0, 0.f, UB200_F16, stream);
}

extern "C" int ub200_gemv_dense(const void* x, const void* W, int64_t ldw, void* out, int m, int k,
                                int dtype, int out_dtype, cudaStream_t stream) {
  using namespace ub;
  if (m <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_UNSUPPORTED;
  if (out_dtype != UB200_F32 && out_dtype != dtype) return UB200_ERR_UNSUPPORTED;
  if (k <= 0 || k % 8 || ldw % 8) return UB200_ERR_BAD_ARG;
  if (!x || !W || !out || !al16(x) || !al16(W)) return UB200_ERR_BAD_ARG;
  const int grid = (m + 15) / 16;
#define GO(T)                                                                                   \
  if (out_dtype == UB200_F32)                                                                   \
    gemv_dense_kernel<T, float, 2><<<grid, 256, 0, stream>>>((const T*)x, (const T*)W, ldw,     \
                                                             (float*)out, m, k);                \
  else                                                                                          \
    gemv_dense_kernel<T, T, 2><<<grid, 256, 0, stream>>>((const T*)x, (const T*)W, ldw, (T*)out, \
                                                         m, k)
  if (dtype == UB200_BF16) { GO(__nv_bfloat16); } else { GO(__half); }
#undef GO
  UB_RETURN_LAST();
}
