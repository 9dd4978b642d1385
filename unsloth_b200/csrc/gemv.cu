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
// bitsandbytes is not vendored in the reference: the arithmetic here restates the published NF4 layout.
// Pair-table kernel (the default): code table rounded to the 16-bit dtype (as bitsandbytes' own GEMV holds
// its quant_map), exact products (HMMA, fp32 accumulate), fp32 absmax applied to fp32 partial sums -- the
// oracle's gemv_nf4(code_dtype=).  Lite kernel (shapes outside the pair kernel's domain): fp32 code * fp32
// absmax, fp32 accumulation.  Both at least as accurate as the 16-bit products of the original; parity
// against bitsandbytes itself is unpinned.
//
// HBM-bound by bytes.  Algorithmic bytes per call: m*k*(0.5 + 1/blocksize) + k*2 + m*2 (NF4);
// m*k*2 + k*2 + m*out_bytes (dense).  Measured (profiles/r2_gemv_bench.log, r2_gemv_ncu.txt): NF4 13.8 us for
// 14336 x 4096 = 2.2 TB/s = 0.33 of the measured copy bandwidth (round 1: 27 us, 0.17); dense 0.94.
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

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

// Lite kernel (round 1; now the fallback for k % 128 != 0, k < 1024 or blocksize < 64): one row per warp; the
// 16-entry fp32 code table is replicated per LANE in shared memory (2 KB: entry n of lane l at n*32 + l), so
// every lookup of a warp is conflict-free and costs one wavefront whatever the nibbles are.  ~6 issued
// instructions and ONE shared-memory lookup per weight: 24-28 us for a 14336 x 4096 weight (1.1-1.3 TB/s), an
// issue / LSU bound, not a DRAM bound (profiles/r1_gemv_nf4_ncu.txt).
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
// Round-2 kernel: byte-pair table + tensor-core products.
//
// The lite kernel above issues ~6 instructions and ONE shared-memory lookup per weight; a warp-wide
// LDS is one wavefront per clock per SM, so 58.7 M weights cost >= 12.4 k clocks of lookups alone.
// Here one lookup serves TWO weights and its result needs no unpacking:
//   * table: 256 entries, entry b = (code[b >> 4], code[b & 15]) as a packed 16-bit pair of the
//     activation dtype, replicated per lane (32 KB; entry b of lane l at word b*32 + l => every
//     warp-wide lookup is one conflict-free wavefront).  The code values are rounded to the 16-bit
//     dtype -- what bitsandbytes' own GEMV does with its quant_map; absmax stays fp32 and is applied
//     to fp32 partial sums, so this is still tighter than the original (oracle: gemv_nf4(code_dtype=)).
//   * a packed byte IS one register of an mma.m16n8k16 A fragment (two k-adjacent elements of a
//     row), so the products run as HMMA with fp32 accumulation: 16 rows x 128 columns per 8 MMAs.
//     The sum over k is order-free, so the k index of the fragment is simply "whatever this lane
//     loaded": lane (g, t) loads 16 B (32 weights) of rows g and g+8 at column t*32 of the chunk.
//   * absmax (one per 64-column block) cannot be applied inside the MMA.  The B operand has 8
//     columns and the GEMV needs one: column n carries x masked to the lanes t == n, so
//     D[row, n] is the partial sum of the 32-column segment n of the chunk, scaled afterwards in
//     fp32 by that segment's absmax (any blocksize >= 32).
// ~1.8 issued instructions and 0.5 LDS wavefronts per weight.  One CTA = 16*RG rows, its 8 warps
// split k in 128-column chunks and meet in shared memory.
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint16_t bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
};
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint16_t bits(float v) { return __half_as_ushort(__float2half_rn(v)); }
};

// Work decomposition.  A CTA owns the 16-row groups blockIdx.x, +gridDim.x, ... and walks them as ONE flat
// sequence of stages: stage = 16 rows x 1024 columns (16 x 512 packed bytes = 8 KB), brought into a 4-deep
// shared-memory ring by a producer warp with sixteen 512-byte bulk copies (cp.async.bulk, completion on an
// mbarrier), across group boundaries.  The kernel lives for a few microseconds on 29 MB: with the packed rows
// prefetched in REGISTERS two items ahead (first versions, profiles/r2_gemv_ncu.txt) 24 warps held 48 KB in
// flight per SM against the ~90 KB that 7 TB/s x DRAM latency needs, and long-scoreboard stalls were 5 per
// issued instruction.  The ring holds 32 KB per CTA, 3 CTAs per SM.  The 8 consumer warps take one 128-column
// chunk of the stage each (warp w: columns w*128 ..): 2 LDS.128, 32 lookups, 8 MMAs; the 32 absmax values of a
// warp's item (16 rows x 2 blocks) are fetched one per LANE, four stages ahead, and handed to the accumulating
// lanes by shuffle.  Requires k % 128 == 0, k >= 1024, blocksize >= 64 (else the lite kernel).
constexpr int GEMV_STAGES = 4;
constexpr int GEMV_ROW_STRIDE = 512 + 64;        // rows 64 B apart (mod 128): the two rows of an LDS.128 phase cover all 32 banks
constexpr int GEMV_STAGE_BYTES = 16 * GEMV_ROW_STRIDE;
constexpr int GEMV_SMEM = 32768 + GEMV_STAGES * GEMV_STAGE_BYTES + 1024 + 1024 + 64 + 64;

template <typename T>
__global__ void __launch_bounds__(288, 3) gemv_nf4_pair_kernel(
    const T* __restrict__ x, const uint8_t* __restrict__ packed,
    const float* __restrict__ absmax_f32, const uint8_t* __restrict__ absmax_q,
    const float* __restrict__ code2, const float* __restrict__ absmax2,
    const float* __restrict__ offset, const float* __restrict__ code16, T* __restrict__ out, int m,
    int k, int bs_shift, int bs2_shift, const T* __restrict__ lora_B, int ldb,
    const float* __restrict__ lora_t, int r, float s) {
  using namespace gemm;
  constexpr int S = GEMV_STAGES;
  extern __shared__ __align__(128) uint8_t gsm[];
  uint32_t* lut = reinterpret_cast<uint32_t*>(gsm);                       // [256 byte values][32 lanes]
  uint8_t* ring = gsm + 32768;                                            // [S][16 rows][512 B + 64 B pad]
  float* code2_s = reinterpret_cast<float*>(ring + S * GEMV_STAGE_BYTES); // [256]
  float* red = code2_s + 256;                                             // [2][8][16]
  uint16_t* c16 = reinterpret_cast<uint16_t*>(red + 256);                 // [16]
  const uint32_t bars = smem_u32(c16) + 64;                               // full[S], empty[S]
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (S + i); };

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half_k = k >> 1;
  const int n_chunks = k >> 7;
  const int n_sc = (k + 1023) >> 10;                   // stages per group
  const int n_groups = (m + 15) >> 4;
  const int my_groups = ((int)blockIdx.x < n_groups) ? (n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_groups * n_sc;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 8); }
    fence_barrier_init();
  }
  // global loads of the prologue are issued here and consumed as late as possible (the kernel is a few
  // microseconds long: a DRAM round trip in front of the table build was 40 % of the warps' lifetime)
  const float code2_mine = (code2 && threadIdx.x < 256) ? code2[threadIdx.x] : 0.f;
  if (threadIdx.x < 16) c16[threadIdx.x] = MmaOp<T>::bits(code16 ? code16[threadIdx.x] : kNF4g[threadIdx.x]);
  __syncthreads();

  if (warp == 8) {
    // ================================ producer =====================================================
    int grp = blockIdx.x, sc = 0;
    for (int j = 0; j < total; ++j) {
      const int slot = j % S;
      mbar_wait(empty(slot), (uint32_t)(((j / S) & 1) ^ 1));
      const int off = sc << 9;
      const uint32_t seg = (uint32_t)min(512, half_k - off);
      if (lane == 0) mbar_expect_tx(full(slot), 16u * seg);
      __syncwarp();
      if (lane < 16) {
        const int row = min((grp << 4) + lane, m - 1);                    // clamped rows: computed, dropped
        const uint8_t* src = packed + (int64_t)row * half_k + off;
        const uint32_t dst = smem_u32(ring) + slot * GEMV_STAGE_BYTES + lane * GEMV_ROW_STRIDE;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(seg), "r"(full(slot)) : "memory");
      }
      if (++sc == n_sc) { sc = 0; grp += gridDim.x; }
    }
    return;
  }

  // ================================ consumers ======================================================
  {
    // warp w writes entries w*32 .. w*32+31 of every lane's copy: lane l -> bank l, conflict-free
    const uint32_t hi0 = c16[2 * warp], hi1 = c16[2 * warp + 1];
#pragma unroll
    for (int i = 0; i < 32; ++i) lut[(warp * 32 + i) * 32 + lane] = (i < 16 ? hi0 : hi1) | ((uint32_t)c16[i & 15] << 16);
  }
  const int g = lane >> 2, t = lane & 3;
  const int bpr = k >> bs_shift;                       // absmax blocks per row
  const uint32_t lane_base = smem_u32(lut) + ((uint32_t)lane << 2);
  const float off = offset ? *offset : 0.f;
  const bool x_lane = (g == t);                        // B column n = g carries the segment of lanes t == n
  const int src0 = 2 * g + (t & 1), src1 = 2 * (g + 8) + (t & 1);   // lanes holding this lane's two absmax (t < 2)

  // absmax prefetch: lane L owns (row L >> 1, block L & 1) of the warp's chunk; raw bits, S stages ahead
  int p_grp = blockIdx.x, p_sc = 0;
  auto blk_of = [&](int grp, int sc) -> int {
    const int row = min((grp << 4) + (lane >> 1), m - 1);
    const int cb = min(((((sc << 3) + warp) << 7) + ((lane & 1) << 6)) >> bs_shift, bpr - 1);
    return row * bpr + cb;
  };
  auto scale_fetch = [&]() -> uint32_t {              // the RAW loaded value: nothing depends on it until it is consumed
    uint32_t v = 0u;
    if (p_grp < n_groups) {
      const int blk = blk_of(p_grp, p_sc);
      v = absmax_f32 ? __float_as_uint(absmax_f32[blk]) : (uint32_t)absmax_q[blk];
      if (++p_sc == n_sc) { p_sc = 0; p_grp += gridDim.x; }
    }
    return v;
  };
  uint32_t sq[S];
#pragma unroll
  for (int i = 0; i < S; ++i) sq[i] = scale_fetch();
  if (threadIdx.x < 256) code2_s[threadIdx.x] = code2_mine;
  asm volatile("bar.sync 1, 256;" ::: "memory");       // table + code2_s visible to the 8 consumer warps

  int c_grp = blockIdx.x, c_sc = 0, parity = 0;
  float acc0 = 0.f, acc1 = 0.f;
  uint32_t xb[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) xb[q] = 0u;             // stays zero in the lanes that feed no B column
  const uint32_t ring_lane = smem_u32(ring) + g * GEMV_ROW_STRIDE + warp * 64 + t * 16;

  auto step = [&](int j, uint32_t& sqv) {
    const int slot = j % S;
    const int c = (c_sc << 3) + warp;
    if (x_lane) {
      if (c < n_chunks) {
        const T* xp = x + (c << 7) + t * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int4 v = __ldg(reinterpret_cast<const int4*>(xp) + q);
          xb[4 * q] = (uint32_t)v.x; xb[4 * q + 1] = (uint32_t)v.y; xb[4 * q + 2] = (uint32_t)v.z; xb[4 * q + 3] = (uint32_t)v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) xb[q] = 0u;       // columns past k (stale ring bytes): x = 0
      }
    }
    // absmax2 is a second (cached) global load: issued here, resolved only after the MMAs
    const float a2 = absmax_f32 ? 0.f : absmax2[blk_of(c_grp, c_sc) >> bs2_shift];
    const uint32_t sq_now = sqv;
    sqv = scale_fetch();
    mbar_wait(full(slot), (uint32_t)((j / S) & 1));
    uint32_t u0[4], u1[4];
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u0[0]), "=r"(u0[1]), "=r"(u0[2]), "=r"(u0[3])
                 : "r"(ring_lane + slot * GEMV_STAGE_BYTES));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u1[0]), "=r"(u1[1]), "=r"(u1[2]), "=r"(u1[3])
                 : "r"(ring_lane + slot * GEMV_STAGE_BYTES + 8 * GEMV_ROW_STRIDE));
    float d[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};   // two accumulation chains
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // MMA i takes bytes 2i, 2i+1 of both rows: columns 4i .. 4i+3 of the lane's segment
      const int wd = i >> 1, b0 = (i & 1) * 2;
      auto look = [&](uint32_t u, int byte) -> uint32_t {
        uint32_t bv, addr, v;
        asm("prmt.b32 %0, %1, 0, %2;" : "=r"(bv) : "r"(u), "r"(0x4440u + (uint32_t)byte));   // zero-extended byte
        asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(bv), "r"(lane_base));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
      };
      const uint32_t a[4] = {look(u0[wd], b0), look(u1[wd], b0), look(u0[wd], b0 + 1), look(u1[wd], b0 + 1)};
      if (i & 1) MmaOp<T>::mma(e, a, xb[2 * i], xb[2 * i + 1]);
      else MmaOp<T>::mma(d, a, xb[2 * i], xb[2 * i + 1]);
    }
    // the slot is released only here: every (volatile, ordered) table lookup above consumed the registers the two
    // LDS.128 filled, so the ring bytes have certainly been read before the producer may overwrite them
    __syncwarp();
    if (lane == 0) mbar_arrive(empty(slot));
    const float mine = absmax_f32 ? __uint_as_float(sq_now) : fmaf(code2_s[sq_now], a2, off);
    const float sc0 = __shfl_sync(0xffffffffu, mine, src0), sc1 = __shfl_sync(0xffffffffu, mine, src1);
    // D columns 2t, 2t+1 = segments 2t, 2t+1 = ONE absmax block (>= 64 columns); columns >= 4 are zero
    acc0 = fmaf(sc0, (d[0] + d[1]) + (e[0] + e[1]), acc0);
    acc1 = fmaf(sc1, (d[2] + d[3]) + (e[2] + e[3]), acc1);
    if (++c_sc == n_sc) {
      // lanes t = 0, 1 of a row hold its partial sums; 8 warps hold 8 k-slices
      float v0 = acc0, v1 = acc1;
      v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      if (t == 0) { red[parity * 128 + warp * 16 + g] = v0; red[parity * 128 + warp * 16 + 8 + g] = v1; }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // red[parity] is rewritten two groups later: one barrier per group
      if (threadIdx.x < 16) {
        const int row = (c_grp << 4) + threadIdx.x;
        if (row < m) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) v += red[parity * 128 + w * 16 + threadIdx.x];
          if (lora_B) {
            float dd = 0.f;
            for (int jj = 0; jj < r; ++jj) dd = fmaf(DT<T>::to_f(lora_B[(int64_t)row * ldb + jj]), lora_t[jj], dd);
            v = fmaf(s, dd, v);
          }
          out[row] = P2<T>::down(v);
        }
      }
      c_sc = 0; c_grp += gridDim.x; parity ^= 1;
      acc0 = acc1 = 0.f;
    }
  };
#pragma unroll 1
  for (int j = 0; j < total; j += S) {
#pragma unroll
    for (int i = 0; i < S; ++i)
      if (j + i < total) step(j + i, sq[i]);
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
  static const int variant = [] { const char* e = getenv("UB200_GEMV"); return e ? atoi(e) : 1; }();   // 0: round-1 kernel (A/B)
  if (variant == 0 || k % 128 || k < 1024 || blocksize < 64 || (int64_t)m * k / blocksize >= (1ll << 31)) {
    gemv_nf4_lite_kernel<T><<<(m + 7) / 8, 256, 0, st>>>(
        (const T*)x, packed, absmax_f32, absmax_q, code2, absmax2, offset, code16, (T*)out, m, k,
        bs_shift, bs2_shift, (const T*)lora_B, ldb, lora_t, r, s);
    UB_RETURN_LAST();
  }
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int n_groups = (m + 15) / 16;
  const int grid = n_groups < n_sm * 3 ? n_groups : n_sm * 3;      // 3 CTAs (66 KB, <= 72 registers) per SM
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemv_nf4_pair_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMV_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  gemv_nf4_pair_kernel<T><<<grid, 288, GEMV_SMEM, st>>>(
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
