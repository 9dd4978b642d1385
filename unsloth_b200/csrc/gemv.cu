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
// HBM-bound.  Algorithmic bytes per call: m*k*(0.5 + 4/blocksize) + k*2 + m*2 (NF4);
// m*k*2 + k*2 + m*out_bytes (dense).  One warp owns ROWS consecutive rows and walks k in
// 1024-column steps (16 packed bytes per lane per row), reusing its slice of x across the rows.
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

// THREADS = 128: small CTAs (8 rows) keep the last partial wave short; the 32 KB table limits an
// SM to 7 of them, registers to 4-5.
constexpr int GEMV_THREADS = 128;
constexpr int GEMV_DEPTH = 4;      // 1024-column steps whose packed bytes are fetched together

template <typename T, int ROWS>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_nf4_kernel(
    const T* __restrict__ x, const uint8_t* __restrict__ packed,
    const float* __restrict__ absmax_f32, const uint8_t* __restrict__ absmax_q,
    const float* __restrict__ code2, const float* __restrict__ absmax2,
    const float* __restrict__ offset, const float* __restrict__ code16, T* __restrict__ out, int m,
    int k, int bs_shift, int bs2_shift, const T* __restrict__ lora_B, int ldb,
    const float* __restrict__ lora_t, int r, float s) {
  using T2 = typename P2<T>::T2;
  // byte -> (code[hi nibble], code[lo nibble]) in fp32, replicated 16x so that lane l always reads
  // copy l % 16: the 16 lanes of each 64-bit shared-memory wavefront hit 16 different bank pairs
  // whatever bytes they look up (an un-replicated table serialises ~7x on random nibbles).
  extern __shared__ __align__(16) unsigned char gemv_smem[];
  float2* lut2 = reinterpret_cast<float2*>(gemv_smem);                 // 32 KB table
  int4* xs = reinterpret_cast<int4*>(gemv_smem + 256 * 16 * sizeof(float2));   // x, k * 2 bytes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = (blockIdx.x * (GEMV_THREADS / 32) + warp) * ROWS;
  const bool live = row0 < m;
  const float off = offset ? *offset : 0.f;
  const int64_t row_bytes = (int64_t)k / 2;
  // rows past the end alias the last row (their results are not stored)
  const uint8_t* wrow[ROWS];
  int64_t ebase[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int row = row0 + i < m ? row0 + i : m - 1;
    wrow[i] = packed + row * row_bytes;
    ebase[i] = (int64_t)row * k;
  }
  int4 w[GEMV_DEPTH][ROWS];
  float am[GEMV_DEPTH][ROWS];
  // all packed bytes and block scales of a 4096-column span are requested before anything is
  // expanded: the first span's loads are in flight while the table is being built
  auto fetch = [&](int cbase) {
#pragma unroll
    for (int d = 0; d < GEMV_DEPTH; ++d) {
      const int c0 = cbase + d * 1024 + lane * 32;
      if (c0 < k) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          w[d][i] = __ldcs(reinterpret_cast<const int4*>(wrow[i] + c0 / 2));
          const int64_t blk = (ebase[i] + c0) >> bs_shift;        // block sizes are powers of two:
          am[d][i] = absmax_f32 ? absmax_f32[blk]                 // no 64-bit divisions in the loop
                                : fmaf(code2[absmax_q[blk]], absmax2[blk >> bs2_shift], off);
        }
      }
    }
  };
  if (live) fetch(0);
  for (int e = threadIdx.x; e < 256 * 16; e += GEMV_THREADS) {   // consecutive lanes, consecutive banks
    const int b = e >> 4;
    lut2[e] = make_float2(code16 ? code16[b >> 4] : kNF4g[b >> 4],
                          code16 ? code16[b & 15] : kNF4g[b & 15]);
  }
  // x is staged once per CTA (every warp walks all of it).  16-byte chunk g = columns 8g..8g+7
  // belongs to lane (g % 128) / 4, quarter g % 4 of 1024-column step g / 128; it is stored at
  // step * 128 + quarter * 32 + lane so that the four 16-byte reads of a lane are conflict-free.
  for (int g = threadIdx.x; g < k / 8; g += GEMV_THREADS) {
    const int within = g & 127;
    xs[(g & ~127) + (within & 3) * 32 + (within >> 2)] = __ldg(reinterpret_cast<const int4*>(x) + g);
  }
  __syncthreads();
  if (!live) return;
  // table address of byte b for this lane = b * 128 + (lane % 16) * 8: the byte is moved to bits
  // 7..14 with one shift and merged with the lane bits by one 3-input logic op
  const uint32_t lane_bits = (uint32_t)(lane & 15) << 3;
  const char* lut_bytes = reinterpret_cast<const char*>(lut2);
  float acc[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int cbase = 0; cbase < k; cbase += GEMV_DEPTH * 1024) {
    if (cbase) fetch(cbase);
#pragma unroll
    for (int d = 0; d < GEMV_DEPTH; ++d) {
      const int c0 = cbase + d * 1024 + lane * 32;
      if (c0 < k) {
        float xf[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          union { int4 v; T2 h[4]; } xv;
          xv.v = xs[(c0 >> 10) * 128 + q * 32 + lane];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = P2<T>::up(xv.h[j]);
            xf[q * 8 + 2 * j] = f.x;
            xf[q * 8 + 2 * j + 1] = f.y;
          }
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          const uint32_t u[4] = {(uint32_t)w[d][i].x, (uint32_t)w[d][i].y, (uint32_t)w[d][i].z,
                                 (uint32_t)w[d][i].w};
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int sh = 8 * (j & 3) - 7;
            const uint32_t word = u[j >> 2];
            const uint32_t o = ((sh < 0 ? word << 7 : word >> sh) & 0x7F80u) | lane_bits;
            const float2 c = *reinterpret_cast<const float2*>(lut_bytes + o);
            p0 = fmaf(c.x, xf[2 * j], p0);
            p1 = fmaf(c.y, xf[2 * j + 1], p1);
          }
          acc[i] = fmaf(am[d][i], p0 + p1, acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int row = row0 + i;
      if (row < m) {
        float v = acc[i];
        if (lora_B) {
          float d = 0.f;
          for (int j = 0; j < r; ++j) d = fmaf(DT<T>::to_f(lora_B[(int64_t)row * ldb + j]), lora_t[j], d);
          v = fmaf(s, d, v);
        }
        out[row] = P2<T>::down(v);
      }
    }
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
  constexpr int ROWS = 2, RPC = ROWS * GEMV_THREADS / 32;          // rows per CTA
  const size_t smem = 256 * 16 * sizeof(float2) + (size_t)((k + 1023) / 1024) * 1024 * sizeof(T);
  if (smem > 220 * 1024) return UB200_ERR_UNSUPPORTED;
  static size_t smem_set = 0;        // opt-in above 48 KB, raised monotonically (idempotent)
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gemv_nf4_kernel<T, ROWS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  gemv_nf4_kernel<T, ROWS><<<(m + RPC - 1) / RPC, GEMV_THREADS, smem, st>>>(
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

// bitsandbytes symbols (unsloth/kernels/utils.py:283-284, call at :955-973): A = x[k],
// B = packed weight [m, k/2], absmax already fp32, datatype = the 16-entry code, n == 1.
// `void` return like the original: argument errors are dropped, CUDA errors surface at next sync.
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
