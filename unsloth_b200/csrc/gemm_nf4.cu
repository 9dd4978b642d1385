// NF4 dequantisation FUSED into the operand staging of the tcgen05 GEMM (north_star: "fused NF4
// dequant -> W@x + B@(A@x) kernel with TMA-staged shared-memory tiles feeding tcgen05"):
//
//     Y[M,N] = X[M,K] . dequant(W_nf4)[N,K]^T  (+ XA[M,R] . B_pad[N,R]^T, the LoRA rank block)
//
// The reference expands the whole weight to 16 bits in HBM first (bitsandbytes'
// cdequantize_blockwise_* launches inside fast_dequantize, unsloth/kernels/utils.py:567-679) and
// then calls cuBLAS (matmul_lora, :1128-1170).  Here the packed 4-bit weight is the only form of W
// that is ever read from HBM: four producer warps per CTA turn each [128 rows x 64 k] slice --
// exactly one 64-weight quantisation block per row, so one thread owns one row -- into the
// 128B-swizzled bf16 K-major smem tile the UMMA descriptor expects, rebuilding the double-quantised
// absmax on the fly; the activation tile still arrives by TMA.  Same CTA-pair pipeline as
// gemm2_kernel (cta_group::2, 256 x 256 tiles, 6-stage ring, double-buffered TMEM accumulators);
// the per-stage "full" barrier now collects the TMA transaction bytes of A plus one arrival per
// producer warp of both CTAs.  Rounding is identical to dequantise-then-GEMM (each weight is
// NF4[q] * absmax rounded once to 16 bits; the k-block order of the accumulation is the same), so
// the result is BIT-IDENTICAL to the two-kernel path -- the parity test checks equality.
//
// Trade-off (measured): the expansion is redone for every 256-row M tile
// (32x at T = 8192) and its ~230 instructions per thread per k-block compete for issue slots with
// the single MMA-issuing thread; in exchange no 16-bit copy of W exists (-14 GB at keep_dequant,
// -2 x 0.5 GB of HBM traffic per projection) and the dequant launch disappears.
#include "tcgen05.cuh"

namespace ub {
namespace gemm {
namespace nf4 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int HALF_N = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;      // TMA, MMA, 4 epilogue warps, 4 dequant producer warps
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr uint32_t B_BYTES = HALF_N * BLOCK_K * 2;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int PREFETCH = 4;           // k-blocks of packed weights in flight per producer thread

__constant__ float kNF4c[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

struct Params {
  CUtensorMap tmap_x;            // X [M, K], box [128 rows, 64 cols]
  CUtensorMap tmap_xa, tmap_b;   // LoRA rank block: XA [M, R] box [128, 64]; B_pad [N, R] box [128, 64]
  const uint8_t* packed;         // [N*K/2]
  const uint8_t* absmax_q;       // [N*K/64]
  const float* code2;            // [256]
  const float* absmax2;          // [N*K/64/256]
  const float* offset;           // device scalar or null
  int M, N, K;
  int kb_dense, kb_lora;         // k-blocks of the dense segment / of the rank block (0 or R/64)
  int m_pairs, n_tiles;
  void* C;
  int64_t ldc;
  int c_dtype;
  int ab_fp16;
  int raster_group;
};

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_nf4_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float lut[16];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8u * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_kb = p.kb_dense + p.kb_lora;
  const int num_tiles = p.m_pairs * p.n_tiles;

  if (threadIdx.x < 16) lut[threadIdx.x] = kNF4c[threadIdx.x];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_x);
    if (p.kb_lora) { prefetch_tmap(&p.tmap_xa); prefetch_tmap(&p.tmap_b); }
  }
  if (warp == 1) {
    if (lane == 0) {
      // full: leader's expect_tx arrive + the peer TMA warp's remote arrive + 4 producer warps x 2 CTAs
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2 + 8); mbar_init(empty_bar(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // m fastest inside groups of `raster_group` M-tiles: concurrently running pairs share weight rows
  auto tile_coords = [&](int tile, int& m_pair, int& n_blk) {
    const int group = p.raster_group;
    const int per = group * p.n_tiles;
    const int g = tile / per;
    const int rem = tile - g * per;
    const int left = p.m_pairs - g * group;
    const int ge = left < group ? left : group;
    n_blk = rem / ge;
    m_pair = g * group + (rem - n_blk * ge);
  };

  if (warp == 0) {
    // ================================ TMA producer: A (and the rank-block operands) ==============
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair; w < num_tiles; w += num_pairs) {
      int m_pair, n_blk;
      tile_coords(w, m_pair, n_blk);
      const int m0 = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M;
      const int n0 = n_blk * BLOCK_N + (int)rank * HALF_N;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const bool dense = kb < p.kb_dense;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(full_bar(stage), dense ? 2u * A_BYTES : 2u * STAGE_BYTES);
          else mbar_arrive_remote(full_bar(stage), 0u);
          if (dense) {
            tma_load_2d_2sm(sa, &p.tmap_x, full_bar(stage), kb * BLOCK_K, m0);
          } else {
            const int k0 = (kb - p.kb_dense) * BLOCK_K;
            tma_load_2d_2sm(sa, &p.tmap_xa, full_bar(stage), k0, m0);
            tma_load_2d_2sm(sb, &p.tmap_b, full_bar(stage), k0, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ===================
    if (rank == 0) {
      const uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N, 0, 0, p.ab_fp16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair; w < num_tiles; w += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa, 16u, 1024u);
          const uint64_t db = make_smem_desc(sb, 16u, 1024u);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16_2sm(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(empty_bar(stage));
            if (kb == total_kb - 1) umma_commit_2sm(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ================================ epilogue (both CTAs, own 128 lanes) ============
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair; w < num_tiles; w += num_pairs) {
      int m_pair, n_blk;
      tile_coords(w, m_pair, n_blk);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        const int col0 = n_blk * BLOCK_N + c;
        if (col0 >= p.N) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), r);
        tmem_ld_wait(r);
        if (row_ok) {
          uint16_t* cptr = reinterpret_cast<uint16_t*>(p.C) + (int64_t)row * p.ldc + col0;
          if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 o;
              uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                ow[j] = p.c_dtype == UB200_BF16
                            ? pack2<__nv_bfloat16>(__uint_as_float(r[i + 2 * j]), __uint_as_float(r[i + 2 * j + 1]))
                            : pack2<__half>(__uint_as_float(r[i + 2 * j]), __uint_as_float(r[i + 2 * j + 1]));
              *reinterpret_cast<uint4*>(cptr + i) = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N)
                cptr[i] = p.c_dtype == UB200_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(r[i])))
                                                  : __half_as_ushort(__float2half_rn(__uint_as_float(r[i])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0u);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================================ NF4 -> bf16 operand producers ==================
    // thread t owns weight row (n0 + t) of this CTA's 128-row half of the B tile: per k-block one
    // 64-weight quantisation block = 32 packed bytes + one double-quantised absmax
    const int t = threadIdx.x - 6 * 32;                 // 0..127
    const float off = p.offset ? *p.offset : 0.f;
    int stage = 0;
    uint32_t phase = 0;
    const int64_t blocks_per_row = p.K >> 6;
    for (int w = pair; w < num_tiles; w += num_pairs) {
      int m_pair, n_blk;
      tile_coords(w, m_pair, n_blk);
      const int n = n_blk * BLOCK_N + (int)rank * HALF_N + t;
      const bool n_ok = n < p.N;
      const int64_t blk0 = (int64_t)n * blocks_per_row;          // first quantisation block of the row
      uint4 raw[PREFETCH][2];
      float am[PREFETCH];
      auto fetch = [&](int kb, int slot) {
        if (n_ok && kb < p.kb_dense) {
          const int64_t blk = blk0 + kb;
          const uint4* src = reinterpret_cast<const uint4*>(p.packed + blk * 32);
          raw[slot][0] = __ldcs(src);
          raw[slot][1] = __ldcs(src + 1);
          am[slot] = __fadd_rn(__fmul_rn(p.code2[p.absmax_q[blk]], p.absmax2[blk >> 8]), off);
        } else {
          raw[slot][0] = make_uint4(0x77777777u, 0x77777777u, 0x77777777u, 0x77777777u);   // code 7 = 0.0
          raw[slot][1] = raw[slot][0];
          am[slot] = 0.f;
        }
      };
#pragma unroll
      for (int s = 0; s < PREFETCH; ++s) fetch(s, s);
#pragma unroll 1
      for (int kb0 = 0; kb0 < total_kb; kb0 += PREFETCH) {
#pragma unroll
        for (int s = 0; s < PREFETCH; ++s) {
          const int kb = kb0 + s;
          if (kb >= total_kb) break;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (kb < p.kb_dense) {
            uint8_t* row = smem_gen + stage * STAGE_BYTES + A_BYTES + t * 128;
            const float a = am[s];
            const uint32_t* words = reinterpret_cast<const uint32_t*>(&raw[s][0]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {                         // 8 weights = 4 packed bytes = one 16-byte chunk
              const uint32_t wv = words[c];
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t byte = (wv >> (8 * j)) & 0xFFu;     // first element in the HIGH nibble
                const float hi = lut[byte >> 4] * a, lo = lut[byte & 15u] * a;
                o[j] = p.ab_fp16 ? pack2<__half>(hi, lo) : pack2<__nv_bfloat16>(hi, lo);
              }
              *reinterpret_cast<uint4*>(row + ((uint32_t)(c ^ (t & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(full_bar(stage));
            else mbar_arrive_remote(full_bar(stage), 0u);
          }
          fetch(kb + PREFETCH, s);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace nf4
}  // namespace gemm
}  // namespace ub

extern "C" int ub200_gemm_nf4(int M, int N, int K, const void* X, int64_t ldx, const uint8_t* packed,
                              const uint8_t* absmax_q, const float* code2, const float* absmax2,
                              const float* offset, int blocksize, int blocksize2, const void* lora_xa,
                              int64_t ld_xa, const void* lora_b, int64_t ld_b, int lora_k, void* C,
                              int64_t ldc, int dtype, cudaStream_t stream) {
  using namespace ub;
  using namespace ub::gemm;
  using namespace ub::gemm::nf4;
  if (M <= 0 || N <= 0) return UB200_OK;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  if (blocksize != 64 || blocksize2 != 256 || K % 64) return UB200_ERR_UNSUPPORTED;   // one quant block per (row, k-block)
  if (lora_k % 64) return UB200_ERR_BAD_ARG;
  Params p;
  memset(&p, 0, sizeof(p));
  const int fp16 = dtype == UB200_F16;
  int rc;
  if ((rc = make_tmap(&p.tmap_x, X, M, K, ldx, BLOCK_M, fp16))) return rc;
  if (lora_k > 0) {
    if ((rc = make_tmap(&p.tmap_xa, lora_xa, M, lora_k, ld_xa, BLOCK_M, fp16))) return rc;
    if ((rc = make_tmap(&p.tmap_b, lora_b, N, lora_k, ld_b, HALF_N, fp16))) return rc;
  }
  p.packed = packed; p.absmax_q = absmax_q; p.code2 = code2; p.absmax2 = absmax2; p.offset = offset;
  p.M = M; p.N = N; p.K = K;
  p.kb_dense = K / 64; p.kb_lora = lora_k / 64;
  p.m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  p.n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  p.C = C; p.ldc = ldc; p.c_dtype = dtype; p.ab_fp16 = fp16;
  {
    const int64_t a_tile = (int64_t)2 * BLOCK_M * K * 2;
    int64_t gm = (28ll << 20) / a_tile;
    if (gm < 1) gm = 1;
    if (gm > p.m_pairs) gm = p.m_pairs;
    p.raster_group = (int)gm;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nf4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = p.m_pairs * p.n_tiles;
  const int pairs = tiles < UB_SM_COUNT / 2 ? tiles : UB_SM_COUNT / 2;
  gemm_nf4_kernel<<<2 * pairs, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  UB_RETURN_LAST();
}
