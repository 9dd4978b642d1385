// NF4 double-quantised weight -> bf16/fp16/fp32, and the blockwise quantiser.
//
// Replaces the two bitsandbytes launches behind the reference's `fast_dequantize`
// (unsloth/kernels/utils.py:650-675): `cdequantize_blockwise_fp32` (absmax u8 -> fp32 via the
// 256-entry code and the second-level absmax, then `+= offset`) followed by
// `cdequantize_blockwise_{fp16,bf16}_nf4`.  bitsandbytes is a third-party wheel that is not
// vendored in the reference; the algorithm is restated in oracle/restate.py (parity unpinned).
//
//  * ub200_dequantize_nf4: ONE launch that folds both stages (no fp32 absmax round trip
//    through HBM).  The backward needs no transposed copy: the GEMM consumes the [out,in]
//    buffer as an MN-major operand.
//  * cdequantize_blockwise_fp32 / cdequantize_blockwise_{bf16,fp16}_nf4: the exact
//    bitsandbytes C symbols the reference binds through ctypes (utils.py:273-284), so the
//    reference's own fast_dequantize can run on this library unchanged.
//  * ub200_quantize_nf4: per-block absmax + nearest-code search (used to build synthetic
//    QLoRA models on device; not on the training hot path).
//
// HBM-bound.  Algorithmic bytes per weight: 0.5 + 1/64 (+4/16384) read, out_bytes written.
#include "common.cuh"

namespace ub {

__constant__ float kNF4[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

// Each thread expands 16 packed bytes (32 weights, half a 64-block).  absmax is either a
// ready fp32 array (absmax_f32 != null: bitsandbytes-compatible stage 2) or rebuilt on the fly
// from the double-quantised statistics.
// POW2: both block sizes are powers of two (always true for bitsandbytes' 64 / 256), so the block
// indices are shifts; the generic instantiation keeps the 64-bit divisions (round-1 SASS: two
// division sequences of ~100 instructions per 32 weights made the kernel issue-bound at 0.63 of HBM).
template <typename T, bool POW2>
__global__ void __launch_bounds__(256) dequant_nf4_kernel(
    const uint8_t* __restrict__ packed, const float* __restrict__ absmax_f32,
    const uint8_t* __restrict__ absmax_q, const float* __restrict__ code2,
    const float* __restrict__ absmax2, const float* __restrict__ offset, T* __restrict__ out,
    int64_t n, int blocksize, int blocksize2, int bs_shift, int bs2_shift) {
  __shared__ float lut[16];
  __shared__ int4 stage_all[8 * 32 * 4];   // 2 KB per warp, 8 warps
  if (threadIdx.x < 16) lut[threadIdx.x] = kNF4[threadIdx.x];
  __syncthreads();
  const float off = offset ? *offset : 0.f;
  const int64_t n_chunks = (n + 31) / 32;  // 32 weights per thread-chunk
  const int lane = threadIdx.x & 31;
  const bool bs_ok = blocksize >= 32 && (blocksize % 32) == 0;
  constexpr int NQ = (int)(32 * sizeof(T) / 16);        // 16-byte pieces per thread
  // warp-uniform loop: a warp owns 32 consecutive chunks = 1024 consecutive weights
  for (int64_t ch0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); ch0 < n_chunks;
       ch0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = ch0 + lane;
    const int64_t e0 = ch * 32;
    const bool warp_fast = bs_ok && (ch0 + 32) * 32 <= n;          // all 32 lanes have a full chunk
    if (warp_fast) {
      const int64_t blk = POW2 ? (e0 >> bs_shift) : e0 / blocksize;
      const int64_t blk2 = POW2 ? (blk >> bs2_shift) : blk / blocksize2;
      const float am = absmax_f32 ? absmax_f32[blk]
                                  : __fadd_rn(__fmul_rn(code2[absmax_q[blk]], absmax2[blk2]), off);
      const int4 raw = __ldcs(reinterpret_cast<const int4*>(packed + e0 / 2));
      const uint8_t* b = reinterpret_cast<const uint8_t*>(&raw);
      alignas(16) T o[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        o[2 * i] = cvt_out<T>(lut[b[i] >> 4] * am);
        o[2 * i + 1] = cvt_out<T>(lut[b[i] & 0xF] * am);
      }
      if (NQ == 4) {
        // Stage the warp's 2 KB through shared memory so that every store instruction writes 512
        // contiguous bytes (lane l -> byte 16*l of the run) instead of 32 scattered 16-byte pieces
        // 64 bytes apart.  Slot swizzle (piece ^ ((lane>>1)&3)) keeps both sides at the 4-wavefront
        // minimum for 32 x 16-byte shared-memory accesses.
        int4* stage = stage_all + (threadIdx.x >> 5) * (32 * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stage[lane * 4 + (j ^ ((lane >> 1) & 3))] = reinterpret_cast<int4*>(o)[j];
        __syncwarp();
        int4* dst = reinterpret_cast<int4*>(out + ch0 * 32);         // the warp's first weight
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = j * 32 + lane, src = idx >> 2, piece = idx & 3;
          __stcs(dst + idx, stage[src * 4 + (piece ^ ((src >> 1) & 3))]);
        }
        __syncwarp();
      } else {
        int4* dst = reinterpret_cast<int4*>(out + e0);
#pragma unroll
        for (int i = 0; i < NQ; ++i) dst[i] = reinterpret_cast<int4*>(o)[i];
      }
    } else if (ch < n_chunks) {
      for (int64_t e = e0; e < n && e < e0 + 32; ++e) {
        const int64_t blk = e / blocksize;
        const float am = absmax_f32 ? absmax_f32[blk]
                                    : __fadd_rn(__fmul_rn(code2[absmax_q[blk]], absmax2[blk / blocksize2]), off);
        const uint8_t byte = packed[e >> 1];
        const int q = (e & 1) ? (byte & 0xF) : (byte >> 4);
        out[e] = cvt_out<T>(lut[q] * am);
      }
    }
  }
}

// generic 8-bit blockwise dequant: out[i] = code[A[i]] * absmax[i / blocksize]
__global__ void __launch_bounds__(256) dequant_blockwise_fp32_kernel(
    const float* __restrict__ code, const uint8_t* __restrict__ A, const float* __restrict__ absmax,
    float* __restrict__ out, int blocksize, int64_t n) {
  __shared__ float lut[256];
  lut[threadIdx.x] = code[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = lut[A[i]] * absmax[i / blocksize];
}

// one warp per quantisation block (blocksize 64: two values per lane)
template <typename T>
__global__ void __launch_bounds__(256) quant_nf4_kernel(const T* __restrict__ W,
                                                        uint8_t* __restrict__ packed,
                                                        float* __restrict__ absmax,
                                                        int64_t n_blocks) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t blk = warp; blk < n_blocks; blk += nwarps) {
    const T* w = W + blk * 64 + lane * 2;
    const float a = DT<T>::to_f(w[0]), b = DT<T>::to_f(w[1]);
    const float am = warp_max(fmaxf(fabsf(a), fabsf(b)));
    if (lane == 0) absmax[blk] = am;
    auto nearest = [&](float v) {
      int best = 0;
      float bd = fabsf(v - kNF4[0]);
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const float d = fabsf(v - kNF4[k]);
        if (d < bd) { bd = d; best = k; }
      }
      return best;
    };
    // divide (not multiply by the reciprocal) to match the oracle's x / absmax
    const int qa = nearest(am > 0.f ? a / am : 0.f), qb = nearest(am > 0.f ? b / am : 0.f);
    packed[blk * 32 + lane] = (uint8_t)((qa << 4) | qb);
  }
}

static inline int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t b = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)UB_SM_COUNT * per_sm;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

template <typename T>
static void launch_dequant(const uint8_t* packed, const float* absmax_f32, const uint8_t* absmax_q,
                           const float* code2, const float* absmax2, const float* offset, void* out,
                           int64_t n, int blocksize, int blocksize2, cudaStream_t st) {
  const int64_t chunks = (n + 31) / 32;
  static int occ = 0;       // resident CTAs per SM: launch exactly one persistent wave
  if (!occ) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dequant_nf4_kernel<T, true>, 256, 0);
    if (occ < 1) occ = 1;
  }
  auto lg2 = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  const int s1 = lg2(blocksize), s2 = lg2(blocksize2);
  if (s1 >= 0 && s2 >= 0)
    dequant_nf4_kernel<T, true><<<grid_for(chunks, 256, occ), 256, 0, st>>>(
        packed, absmax_f32, absmax_q, code2, absmax2, offset, (T*)out, n, blocksize, blocksize2, s1, s2);
  else
    dequant_nf4_kernel<T, false><<<grid_for(chunks, 256, occ), 256, 0, st>>>(
        packed, absmax_f32, absmax_q, code2, absmax2, offset, (T*)out, n, blocksize, blocksize2, 0, 0);
}

}  // namespace ub

extern "C" int ub200_dequantize_nf4(const uint8_t* packed, const uint8_t* absmax_q,
                                    const float* code2, const float* absmax2, const float* offset,
                                    void* out, int64_t n, int blocksize, int blocksize2,
                                    int out_dtype, cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  if (blocksize <= 0 || blocksize2 <= 0) return UB200_ERR_BAD_ARG;
  if (out_dtype == UB200_BF16) launch_dequant<__nv_bfloat16>(packed, nullptr, absmax_q, code2, absmax2, offset, out, n, blocksize, blocksize2, stream);
  else if (out_dtype == UB200_F16) launch_dequant<__half>(packed, nullptr, absmax_q, code2, absmax2, offset, out, n, blocksize, blocksize2, stream);
  else if (out_dtype == UB200_F32) launch_dequant<float>(packed, nullptr, absmax_q, code2, absmax2, offset, out, n, blocksize, blocksize2, stream);
  else return UB200_ERR_BAD_ARG;
  UB_RETURN_LAST();
}

// ---- bitsandbytes-compatible symbols (void return, errors surface at the next sync) --------
extern "C" void cdequantize_blockwise_fp32(float* code, unsigned char* A, float* absmax,
                                           float* out, int blocksize, const int n,
                                           cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return;
  dequant_blockwise_fp32_kernel<<<grid_for(n, 256, 16), 256, 0, stream>>>(code, A, absmax, out,
                                                                         blocksize, n);
}
extern "C" void cdequantize_blockwise_bf16_nf4(float* code, unsigned char* A, float* absmax,
                                               __nv_bfloat16* out, int blocksize, const int n,
                                               cudaStream_t stream) {
  (void)code;
  if (n <= 0) return;
  ub::launch_dequant<__nv_bfloat16>(A, absmax, nullptr, nullptr, nullptr, nullptr, out, n,
                                    blocksize, 1, stream);
}
extern "C" void cdequantize_blockwise_fp16_nf4(float* code, unsigned char* A, float* absmax,
                                               __half* out, int blocksize, const int n,
                                               cudaStream_t stream) {
  (void)code;
  if (n <= 0) return;
  ub::launch_dequant<__half>(A, absmax, nullptr, nullptr, nullptr, nullptr, out, n, blocksize, 1,
                             stream);
}

extern "C" int ub200_quantize_nf4(const void* W, int dtype, uint8_t* packed, float* absmax,
                                  int64_t n, int blocksize, cudaStream_t stream) {
  using namespace ub;
  if (n <= 0) return UB200_OK;
  if (blocksize != 64 || n % 64) return UB200_ERR_UNSUPPORTED;
  const int64_t nb = n / 64;
  const int grid = grid_for(nb * 32, 256, 16);
  if (dtype == UB200_BF16) quant_nf4_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)W, packed, absmax, nb);
  else if (dtype == UB200_F16) quant_nf4_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)W, packed, absmax, nb);
  else if (dtype == UB200_F32) quant_nf4_kernel<float><<<grid, 256, 0, stream>>>((const float*)W, packed, absmax, nb);
  else return UB200_ERR_BAD_ARG;
  UB_RETURN_LAST();
}
