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
// 16-bit pair table feeding HFMA2 / mma fragments -- round-2 work.
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

// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL (round-2 candidate; written after the round-1 GPU budget was spent: compiles for
// sm_100a, NOT yet run on hardware, not on any product path -- reachable only through
// ub200_gemv_nf4_mma).  The round-1 finding is that a table lookup + FFMA per weight costs ~6
// issued instructions per weight, an issue bound 3x above the DRAM time.  Here a byte (two
// weights) is looked up ONCE as a packed 16-bit pair and fed straight into the A fragment of
// mma.sync.m16n8k16 (fp32 accumulate; the B fragment carries x, identical in all 8 columns), so
// the multiply-adds cost one instruction per 256 weights and the total is ~2 instructions per
// weight.  k order inside a 64-column block is permuted per lane (lane%4 = c owns columns
// 16c..16c+15), which a dot product does not care about as long as x is permuted alike.
// CTA = 32 rows x all of k: 8 warps interleave the 64-column steps (split-k), partial sums meet in
// shared memory.  Per-block scale: the 4 mma of a step accumulate into a scratch fragment that is
// scaled by the row's absmax once per 64 columns.
// ---------------------------------------------------------------------------------------------
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void run(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                             uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void run(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                             uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) gemv_nf4_mma_kernel(
    const T* __restrict__ x, const uint8_t* __restrict__ packed,
    const float* __restrict__ absmax_f32, const uint8_t* __restrict__ absmax_q,
    const float* __restrict__ code2, const float* __restrict__ absmax2,
    const float* __restrict__ offset, const float* __restrict__ code16, T* __restrict__ out, int m,
    int k, int bs_shift, int bs2_shift, const T* __restrict__ lora_B, int ldb,
    const float* __restrict__ lora_t, int r, float s) {
  // byte -> (code[hi nibble], code[lo nibble]) as a 16-bit pair, replicated per lane: 32 KB,
  // every lookup of a warp is one conflict-free wavefront
  __shared__ uint32_t lut[256 * 32];
  __shared__ float red[8][32];
  for (int e = threadIdx.x; e < 256 * 32; e += 256) {
    const int b = e >> 5;
    lut[e] = MmaOp<T>::pack(code16 ? code16[b >> 4] : kNF4g[b >> 4],
                            code16 ? code16[b & 15] : kNF4g[b & 15]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int row_base = blockIdx.x * 32;
  const float off = offset ? *offset : 0.f;
  const int64_t row_bytes = (int64_t)k / 2;
  const uint32_t lane_bits = (uint32_t)lane << 2;
  const char* lut_bytes = reinterpret_cast<const char*>(lut);
  // rows of this lane: [rg][h] = row_base + rg*16 + g + 8h (clamped; results of clamped rows are dropped)
  const uint8_t* wrow[2][2];
#pragma unroll
  for (int rg = 0; rg < 2; ++rg)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int row = row_base + rg * 16 + g + 8 * h;
      row = row < m ? row : m - 1;
      wrow[rg][h] = packed + row * row_bytes + c * 8;
    }
  // block scales: the 4 lanes of a group share their 4 rows; lane c rebuilds the scale of row
  // #c (= rg*2 + h) and the group exchanges them by shuffle -- one dependent load chain per lane
  // and step instead of four
  int my_row = row_base + (c >> 1) * 16 + g + 8 * (c & 1);
  my_row = my_row < m ? my_row : m - 1;
  const int64_t my_blk0 = ((int64_t)my_row * k) >> bs_shift;       // k % blocksize == 0 (checked by the host)
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int n_steps = k >> 6;
  uint2 wq[2][2], wn[2][2];
  float am_mine, an_mine;
  auto fetch = [&](int step, uint2 (&w)[2][2], float& a_mine) {
#pragma unroll
    for (int rg = 0; rg < 2; ++rg)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        w[rg][h] = __ldcs(reinterpret_cast<const uint2*>(wrow[rg][h] + step * 32));
    const int64_t blk = my_blk0 + ((step << 6) >> bs_shift);
    a_mine = absmax_f32 ? absmax_f32[blk] : fmaf(code2[absmax_q[blk]], absmax2[blk >> bs2_shift], off);
  };
  int step = warp;                                  // the 8 warps interleave the 64-column steps
  if (step < n_steps) fetch(step, wq, am_mine);
#pragma unroll 1
  for (; step < n_steps; step += 8) {
    const bool more = step + 8 < n_steps;
    if (more) fetch(step + 8, wn, an_mine);
    float am[2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i >> 1][i & 1] = __shfl_sync(0xffffffffu, am_mine, (lane & ~3) | i);
    // B fragment source: this lane's 16 columns of x (8 packed pairs), the same for every g
    union { int4 v[2]; uint32_t p[8]; } xv;
    const int4* xp = reinterpret_cast<const int4*>(x + (step << 6) + c * 16);
    xv.v[0] = __ldg(xp);
    xv.v[1] = __ldg(xp + 1);
#pragma unroll
    for (int rg = 0; rg < 2; ++rg) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t w0[2] = {wq[rg][0].x, wq[rg][0].y}, w1[2] = {wq[rg][1].x, wq[rg][1].y};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        // bytes 2t and 2t+1 of the 8-byte chunk: word t/2, byte positions 2(t&1), 2(t&1)+1
        const int sh0 = 16 * (t & 1) - 7, sh1 = sh0 + 8;
        const uint32_t u0 = w0[t >> 1], u1 = w1[t >> 1];
        const uint32_t o00 = ((sh0 < 0 ? u0 << 7 : u0 >> sh0) & 0x7F80u) | lane_bits;   // row g,   byte 2t
        const uint32_t o10 = ((sh0 < 0 ? u1 << 7 : u1 >> sh0) & 0x7F80u) | lane_bits;   // row g+8, byte 2t
        const uint32_t o01 = ((u0 >> sh1) & 0x7F80u) | lane_bits;                        // row g,   byte 2t+1
        const uint32_t o11 = ((u1 >> sh1) & 0x7F80u) | lane_bits;                        // row g+8, byte 2t+1
        MmaOp<T>::run(d, *reinterpret_cast<const uint32_t*>(lut_bytes + o00),
                      *reinterpret_cast<const uint32_t*>(lut_bytes + o10),
                      *reinterpret_cast<const uint32_t*>(lut_bytes + o01),
                      *reinterpret_cast<const uint32_t*>(lut_bytes + o11), xv.p[2 * t], xv.p[2 * t + 1]);
      }
      acc[rg][0] = fmaf(am[rg][0], d[0], acc[rg][0]);      // d[0], d[1]: row g   (all 8 columns equal)
      acc[rg][1] = fmaf(am[rg][1], d[2], acc[rg][1]);      // d[2], d[3]: row g+8
    }
    if (more) {
#pragma unroll
      for (int rg = 0; rg < 2; ++rg)
#pragma unroll
        for (int h = 0; h < 2; ++h) wq[rg][h] = wn[rg][h];
      am_mine = an_mine;
    }
  }
  if (c == 0) {
#pragma unroll
    for (int rg = 0; rg < 2; ++rg)
#pragma unroll
      for (int h = 0; h < 2; ++h) red[warp][rg * 16 + g + 8 * h] = acc[rg][h];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int row = row_base + threadIdx.x;
    if (row < m) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      if (lora_B) {
        float dsum = 0.f;
        for (int j = 0; j < r; ++j) dsum = fmaf(DT<T>::to_f(lora_B[(int64_t)row * ldb + j]), lora_t[j], dsum);
        v = fmaf(s, dsum, v);
      }
      out[row] = P2<T>::down(v);
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

// EXPERIMENTAL entry (see gemv_nf4_mma_kernel): same contract as ub200_gemv_nf4, k % 64 == 0 and
// blocksize >= 64.  Not used by unsloth_b200.kernels; benchmarks/kernel_bench.py gemv_mma times it.
extern "C" int ub200_gemv_nf4_mma(const void* x, const uint8_t* packed, const float* absmax_f32,
                                  const uint8_t* absmax_q, const float* code2, const float* absmax2,
                                  const float* offset, const float* code16, void* out, int m, int k,
                                  int blocksize, int blocksize2, const void* lora_B, int ldb,
                                  const float* lora_t, int r, float s, int dtype,
                                  cudaStream_t stream) {
  using namespace ub;
  if (m <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_UNSUPPORTED;
  if (k <= 0 || k % 64 || blocksize < 64 || (blocksize & (blocksize - 1)) || k % blocksize)
    return UB200_ERR_UNSUPPORTED;
  if (!absmax_f32 && (!absmax_q || !code2 || !absmax2 || blocksize2 <= 0)) return UB200_ERR_BAD_ARG;
  if (!absmax_f32 && (blocksize2 & (blocksize2 - 1))) return UB200_ERR_UNSUPPORTED;
  if (lora_B && (!lora_t || r <= 0 || ldb < r)) return UB200_ERR_BAD_ARG;
  if (!x || !packed || !out || !al16(x) || !al16(packed)) return UB200_ERR_BAD_ARG;
  int bs_shift = 0, bs2_shift = 0;
  while ((1 << bs_shift) < blocksize) ++bs_shift;
  while ((1 << bs2_shift) < blocksize2) ++bs2_shift;
  const int grid = (m + 31) / 32;
  if (dtype == UB200_BF16)
    gemv_nf4_mma_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16,
        (__nv_bfloat16*)out, m, k, bs_shift, bs2_shift, (const __nv_bfloat16*)lora_B, ldb, lora_t, r, s);
  else
    gemv_nf4_mma_kernel<__half><<<grid, 256, 0, stream>>>(
        (const __half*)x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16, (__half*)out,
        m, k, bs_shift, bs2_shift, (const __half*)lora_B, ldb, lora_t, r, s);
  UB_RETURN_LAST();
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
