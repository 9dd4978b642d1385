// Multi-segment bf16/fp16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//     C[M,N] (+)= alpha * sum_s  A_s[M,K_s] . B_s[N,K_s]^T
//
// This is the primitive under the LoRA projections of the reference:
//   unsloth/kernels/utils.py:1128-1170  matmul_lora:  X @ W.T  + (X @ A.T) @ (s B.T)
//   unsloth/kernels/fast_lora.py:172-204, 476-517, 639-647   dX and dA/dB GEMMs
// where the reference issues one cuBLAS call per term and rounds the running result to bf16
// between terms.  Here every term is a K-SEGMENT of one launch: the dense base GEMM, the
// rank-r LoRA update (as one extra K block) and -- for dX -- the sum over projections all
// accumulate in the same fp32 TMEM accumulator and are rounded once.
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) -> smem ring, mbarrier tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees stages
//   warps 2..5  epilogue: tcgen05.ld accumulator -> registers -> alpha/beta -> global
//   accumulators are double buffered in TMEM so the epilogue of tile i overlaps the
//   main loop of tile i+1.
// EPI = 1 instantiations (ub200_gemm_glu): the same main loop with SIXTEEN epilogue warps (576 threads) that apply
// the gated activation of LoRA_MLP to the accumulator tile -- see "Gated-activation epilogue" below.
// Operand layouts: each operand may be K-major (row-major [MN, K]) or MN-major (row-major
// [K, MN]); MN-major is what makes dX (= dY @ W with W stored [out,in]) and the dA/dB
// reductions over tokens (X^T @ G) run without any transpose pass.
//
// Roofline: tensor-bound; flops = 2*M*N*sum(K_s).
#include "tcgen05.cuh"
#include "glu_math.cuh"

namespace ub {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
// kernels with the gated-activation epilogue run SIXTEEN epilogue warps (four per TMEM lane quarter, interleaved
// 16-column units): the epilogue does ~37 instructions per element on 3-5x the bytes of a plain store, and with
// four warps (one per scheduler) it was latency-bound at 1.6x the tile's MMA time (profiles/r2_glu_epilogue_ncu.txt)
constexpr int GLU_WARPS_PER_QUARTER = 4;
constexpr int NUM_THREADS_GLU = 64 + 128 * GLU_WARPS_PER_QUARTER;
__host__ __device__ constexpr int threads_of(int epi) { return epi ? NUM_THREADS_GLU : NUM_THREADS; }
constexpr int MAX_SEGS = UB200_GEMM_MAX_SEGMENTS;
constexpr uint32_t SMEM_BUDGET = 200 * 1024;

struct Params {
  CUtensorMap tmap_a[MAX_SEGS];
  CUtensorMap tmap_b[MAX_SEGS];
  int seg_kblocks[MAX_SEGS];
  int n_segs;
  int M, N;
  int a_mn, b_mn;        // operand majors (uniform over segments)
  int m_tiles, n_tiles;
  int split_k;           // >= 1
  void* C;               // output (or fp32 workspace when split_k > 1)
  int64_t ldc;
  int64_t split_stride;  // elements between split slices of the workspace
  int c_dtype;           // UB200_F32 / F16 / BF16
  int accumulate;        // C = alpha*acc + C
  float alpha;
  int ab_fp16;           // operands are fp16 instead of bf16
  int raster_mode;       // 0: groups of `raster_group` M-tiles stay L2-resident, sweep N; 1: N-groups, sweep M
  int raster_group;
  // gated-activation epilogue (kernels instantiated with EPI = 1, ub200_gemm_glu); 16-bit output only
  int glu_mode;          // UB200_GLU_EPI_FWD / UB200_GLU_EPI_BWD
  int glu_act;           // ACT_SWIGLU / ACT_GEGLU_APPROX / ACT_GEGLU_EXACT
  void* glu_e;           // [M, N] fwd: read (gate output);      bwd: read e, overwritten with df
  void* glu_g;           // [M, N] fwd: written (up output);     bwd: read g, overwritten with de
  int64_t ld_eg;
};

// L2-aware rasterisation of the tile grid.  The two 63 MB L2 halves each keep their own copy of
// data shared across dies, so the operand that is re-used across a sweep must fit in ~30 MB:
//   mode 0: `group` M-tiles stay resident while all N-tiles stream past (m fastest inside a group,
//           so the CTAs running together share the same weight tile);
//   mode 1: `group` N-tiles stay resident while all M-tiles stream past.
__device__ __forceinline__ void raster_coords(int tile, int m_total, int n_total, int mode, int group,
                                              int& m_idx, int& n_idx) {
  if (mode == 0) {
    const int per = group * n_total;
    const int g = tile / per;
    const int rem = tile - g * per;
    const int left = m_total - g * group;
    const int ge = left < group ? left : group;
    n_idx = rem / ge;
    m_idx = g * group + (rem - n_idx * ge);
  } else {
    const int per = group * m_total;
    const int g = tile / per;
    const int rem = tile - g * per;
    const int left = n_total - g * group;
    const int ge = left < group ? left : group;
    m_idx = rem / ge;
    n_idx = g * group + (rem - m_idx * ge);
  }
}

// Epilogue store of 32 consecutive accumulator columns of one row: alpha, optional beta=1
// read-modify-write, conversion to the output dtype, 16-byte global stores.
__device__ __forceinline__ void store_chunk(const Params& p, const uint32_t (&r)[32], bool has_k,
                                            int row, int col0, int sp) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) * p.alpha : 0.f;
  const bool full = (col0 + 32 <= p.N);
  if (p.c_dtype == UB200_F32) {
    float* cptr = reinterpret_cast<float*>(p.C) + (int64_t)sp * p.split_stride +
                  (int64_t)row * p.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(cptr + i);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(cptr + i) = o;
      }
    } else {
      for (int i = 0; i < 32 && col0 + i < p.N; ++i)
        cptr[i] = p.accumulate ? cptr[i] + v[i] : v[i];
    }
  } else {
    // 16-bit output (bf16 / fp16)
    uint16_t* cptr = reinterpret_cast<uint16_t*>(p.C) + (int64_t)row * p.ldc + col0;
    const bool bf = p.c_dtype == UB200_BF16;
    if (full && ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o;
        if (p.accumulate) {
          const uint4 old = *reinterpret_cast<const uint4*>(cptr + i);
          const uint16_t* oh = reinterpret_cast<const uint16_t*>(&old);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[i + j] += bf ? __bfloat162float(__ushort_as_bfloat16(oh[j]))
                           : __half2float(__ushort_as_half(oh[j]));
        }
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (bf) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
            ow[j] = *reinterpret_cast<uint32_t*>(&t);
          } else {
            __half2 t = __floats2half2_rn(v[i + 2 * j], v[i + 2 * j + 1]);
            ow[j] = *reinterpret_cast<uint32_t*>(&t);
          }
        }
        *reinterpret_cast<uint4*>(cptr + i) = o;
      }
    } else {
      for (int i = 0; i < 32 && col0 + i < p.N; ++i) {
        float val = v[i];
        if (p.accumulate)
          val += bf ? __bfloat162float(__ushort_as_bfloat16(cptr[i]))
                    : __half2float(__ushort_as_half(cptr[i]));
        cptr[i] = bf ? __bfloat16_as_ushort(__float2bfloat16_rn(val))
                     : __half_as_ushort(__float2half_rn(val));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Gated-activation epilogue (EPI = 1).  The accumulator tile never makes the round trip through HBM
// that the two-kernel form (GEMM, then glu.cu) pays:
//   backward (fast_lora.py:155-157 + swiglu.py:86-109 / geglu.py:214-244): the tile is DW = dY @ W_down
//     (+ LoRA); it is rounded to the tensor dtype exactly where the reference's matmul output is, then
//     h = f(e) * g -> C,  df = DW * f -> e's buffer,  de = (DW * g) * f'(e) -> g's buffer (in place);
//   forward (fast_lora.py:84-87 + swiglu.py:37-47): the tile is g = X @ W_up (+ LoRA); e (the gate
//     projection, written by the previous launch) is read, g is stored and h = f(e) * g -> C.
// Same per-element code and rounding points as glu.cu (glu_math.cuh), so both forms give the same bits.
// One thread owns one row; the unit of work is 16 consecutive columns = ONE 32-byte sector of every tensor
// (LDG.E.256 / STG.E.256).  The operands of a unit are loaded one unit ahead of their use (two register
// buffers; across tile boundaries too, see glu_epilogue_tile); loads are evict-first, the operand tiles of
// the main loop own the L2.  History of the design with measurements: DESIGN.md 4.1c.
// ---------------------------------------------------------------------------------------
struct GluRegs {
  uint32_t e[8];      // 16 consecutive 16-bit values of this thread's row (one 32-byte sector)
  uint32_t g[8];
};

// 32-byte global accesses (sm_100: LDG.E.256 / STG.E.256): one thread moves a whole 32-byte sector per
// instruction.  Loads are evict-first (the operand tiles of the main loop own the L2).
__device__ __forceinline__ void ldg256_cs(const void* ptr, uint32_t* r) {
  asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
}
__device__ __forceinline__ void stg256(void* ptr, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// operands of one 16-column unit (row, col0 .. col0 + 15)
__device__ __forceinline__ void glu_prefetch(const Params& p, GluRegs& q, int row, int col0) {
  ldg256_cs(reinterpret_cast<const uint16_t*>(p.glu_e) + (int64_t)row * p.ld_eg + col0, q.e);
  if (p.glu_mode == UB200_GLU_EPI_BWD)
    ldg256_cs(reinterpret_cast<const uint16_t*>(p.glu_g) + (int64_t)row * p.ld_eg + col0, q.g);
}

// element `hi` (0 / 1) of a packed pair of 16-bit values, as fp32 (register-only: no address is taken)
template <typename T> __device__ __forceinline__ float unpack16(uint32_t w, int hi);
template <> __device__ __forceinline__ float unpack16<__nv_bfloat16>(uint32_t w, int hi) {
  return __uint_as_float(hi ? (w & 0xffff0000u) : (w << 16));
}
template <> __device__ __forceinline__ float unpack16<__half>(uint32_t w, int hi) {
  return __half2float(__ushort_as_half((unsigned short)(hi ? (w >> 16) : (w & 0xffffu))));
}
// two floats -> one packed pair, round to nearest even (the rounding of DT<T>::from_f)
template <typename T> __device__ __forceinline__ uint32_t pack16(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack16<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack16<__half>(float lo, float hi) {
  const __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// one 16-column unit: accumulators r (alpha == 1, checked by the host) + operands q -> 2 or 3 sector stores
template <typename T, int ACT>
__device__ __forceinline__ void glu_finish_t(const Params& p, const uint32_t (&r)[16], const GluRegs& q,
                                             int row, int col0) {
  const bool bwd = p.glu_mode == UB200_GLU_EPI_BWD;
  uint32_t w0[8], w1[8], w2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float o0[2], o1[2], o2[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float acc = DT<T>::rnd(__uint_as_float(r[2 * k + j]));
      const float ek = unpack16<T>(q.e[k], j);
      if (bwd) glu_bwd_elem<T, ACT>(acc, ek, unpack16<T>(q.g[k], j), o0[j], o1[j], o2[j]);
      else { o0[j] = glu_fwd_elem<T, ACT>(ek, acc); o1[j] = 0.f; o2[j] = acc; }
    }
    w0[k] = pack16<T>(o0[0], o0[1]);
    w1[k] = pack16<T>(o1[0], o1[1]);
    w2[k] = pack16<T>(o2[0], o2[1]);
  }
  stg256(reinterpret_cast<T*>(p.C) + (int64_t)row * p.ldc + col0, w0);                      // h
  if (bwd) stg256(reinterpret_cast<T*>(p.glu_e) + (int64_t)row * p.ld_eg + col0, w1);       // df (over e)
  stg256(reinterpret_cast<T*>(p.glu_g) + (int64_t)row * p.ld_eg + col0, w2);                // de (over g) / g
}

__device__ __forceinline__ void glu_finish(const Params& p, const uint32_t (&r)[16], const GluRegs& q, int row,
                                           int col0) {
  if (p.c_dtype == UB200_BF16) {
    if (p.glu_act == ACT_SWIGLU) glu_finish_t<__nv_bfloat16, ACT_SWIGLU>(p, r, q, row, col0);
    else if (p.glu_act == ACT_GEGLU_APPROX) glu_finish_t<__nv_bfloat16, ACT_GEGLU_APPROX>(p, r, q, row, col0);
    else glu_finish_t<__nv_bfloat16, ACT_GEGLU_EXACT>(p, r, q, row, col0);
  } else {
    if (p.glu_act == ACT_SWIGLU) glu_finish_t<__half, ACT_SWIGLU>(p, r, q, row, col0);
    else if (p.glu_act == ACT_GEGLU_APPROX) glu_finish_t<__half, ACT_GEGLU_APPROX>(p, r, q, row, col0);
    else glu_finish_t<__half, ACT_GEGLU_EXACT>(p, r, q, row, col0);
  }
}

// One accumulator tile through the gated-activation epilogue, for ONE of the GLU_WARPS_PER_QUARTER warps that
// share a TMEM lane quarter: warp `sub` takes the 16-column units sub, sub + 4, sub + 8, ...  (BLOCK_N / 64 per
// tile; the host guarantees N % BLOCK_N == 0, so every tile is whole).  Two register buffers: q0 arrives loaded
// with the tile's first unit; while a unit is computed the operands of the next one are in flight, and during
// the LAST unit those of the NEXT TILE's first unit (nrow, ncol; nrow < 0: none) -- no load latency is exposed
// at a tile boundary.
template <int BLOCK_N>
__device__ __forceinline__ void glu_epilogue_tile(const Params& p, uint32_t tmem_tile, GluRegs& q0, int row,
                                                  bool row_ok, int col_tile, int sub, int nrow, int ncol) {
  constexpr int UNITS = BLOCK_N / (16 * GLU_WARPS_PER_QUARTER);
  uint32_t r[16];
  if constexpr (UNITS == 1) {
    tmem_ld16(tmem_tile + (uint32_t)(16 * sub), r);
    tmem_ld_wait16(r);
    if (row_ok) glu_finish(p, r, q0, row, col_tile + 16 * sub);
    if (nrow >= 0) glu_prefetch(p, q0, nrow, ncol);
  } else {
    GluRegs q1;
#pragma unroll 1
    for (int j = 0; j < UNITS; j += 2) {
      const int cA = 16 * sub + 64 * j, cB = cA + 64;
      if (row_ok) glu_prefetch(p, q1, row, col_tile + cB);
      tmem_ld16(tmem_tile + (uint32_t)cA, r);
      tmem_ld_wait16(r);
      if (row_ok) glu_finish(p, r, q0, row, col_tile + cA);
      if (j + 2 < UNITS) {
        if (row_ok) glu_prefetch(p, q0, row, col_tile + cB + 64);
      } else if (nrow >= 0) {
        glu_prefetch(p, q0, nrow, ncol);
      }
      tmem_ld16(tmem_tile + (uint32_t)cB, r);
      tmem_ld_wait16(r);
      if (row_ok) glu_finish(p, r, q1, row, col_tile + cB);
    }
  }
}

template <int BLOCK_N>
struct Cfg {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr uint32_t TMEM_COLS = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);  // power of 2
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int EPI = 0>
__global__ void __launch_bounds__(threads_of(EPI), 1)
gemm_kernel(const __grid_constant__ Params p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 8u * (2 * C::STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int total_kb = 0;
  for (int s = 0; s < p.n_segs; ++s) total_kb += p.seg_kblocks[s];
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int num_work = num_tiles * p.split_k;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_segs; ++s) { prefetch_tmap(&p.tmap_a[s]); prefetch_tmap(&p.tmap_b[s]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI ? 4 * GLU_WARPS_PER_QUARTER : 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // split-K range helper: k-blocks [kb0, kb1) of the concatenated segment list
  auto work_range = [&](int w, int& tile, int& kb0, int& kb1) {
    tile = w % num_tiles;
    const int sp = w / num_tiles;
    const int per = (total_kb + p.split_k - 1) / p.split_k;
    kb0 = sp * per;
    kb1 = min(total_kb, kb0 + per);
  };
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    raster_coords(tile, p.m_tiles, p.n_tiles, p.raster_mode, p.raster_group, m_blk, n_blk);
  };

  if (warp == 0) {
    // ================================ TMA producer ===================================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int tile, kb0, kb1, m_blk, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_blk, n_blk);
        int seg = 0, seg_start = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          while (kb >= seg_start + p.seg_kblocks[seg]) { seg_start += p.seg_kblocks[seg]; ++seg; }
          const int k0 = (kb - seg_start) * BLOCK_K;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d(sa, &p.tmap_a[seg], full_bar(stage), k0, m_blk * BLOCK_M);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d(sa + j * 8192u, &p.tmap_a[seg], full_bar(stage), m_blk * BLOCK_M + j * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(sb, &p.tmap_b[seg], full_bar(stage), k0, n_blk * BLOCK_N);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_2d(sb + j * 8192u, &p.tmap_b[seg], full_bar(stage), n_blk * BLOCK_N + j * 64, k0);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =====================================
    {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, p.a_mn, p.b_mn, p.ab_fp16);
      // per-UMMA_K advance of the descriptor start address (in 16-byte units)
      const uint32_t a_adv = p.a_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
      const uint32_t b_adv = p.b_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int tile, kb0, kb1;
        work_range(w, tile, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = make_smem_desc(sa, a_lbo, 1024u);
          const uint64_t db = make_smem_desc(sb, b_lbo, 1024u);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_f16(d_tmem, da + (uint64_t)(a_adv * k), db + (uint64_t)(b_adv * k), idesc,
                       (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));        // frees the smem stage when the MMAs retire
            if (kb == kb1 - 1) umma_commit(tfull_bar(acc));   // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) {                         // empty split: still hand the buffer over
          if (elect_one()) umma_commit(tfull_bar(acc));
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue =======================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI) {
      // gated-activation epilogue: GLU_WARPS_PER_QUARTER warps per lane quarter, see glu_epilogue_tile
      const int sub = (warp - 2) >> 2;
      auto coords = [&](int w, int& row, int& col_tile) {
        int tile, kb0, kb1, m_blk, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_blk, n_blk);
        row = m_blk * BLOCK_M + q * 32 + lane;
        col_tile = n_blk * BLOCK_N;
      };
      GluRegs gq0;
      int row, col_tile;
      if ((int)blockIdx.x < num_work) {
        coords(blockIdx.x, row, col_tile);
        if (row < p.M) glu_prefetch(p, gq0, row, col_tile + 16 * sub);
      }
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        coords(w, row, col_tile);
        int nrow = -1, ncol = 0;
        if (w + (int)gridDim.x < num_work) {
          int nr, nc;
          coords(w + gridDim.x, nr, nc);
          if (nr < p.M) { nrow = nr; ncol = nc + 16 * sub; }
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        glu_epilogue_tile<BLOCK_N>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N), gq0,
                                   row, row < p.M, col_tile, sub, nrow, ncol);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else {
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int tile, kb0, kb1, m_blk, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_blk, n_blk);
        const int sp = w / num_tiles;
        const int row = m_blk * BLOCK_M + q * 32 + lane;
        const bool row_ok = row < p.M;
        const bool has_k = kb1 > kb0;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          const int col0 = n_blk * BLOCK_N + c;
          if (col0 >= p.N) break;                   // warp-uniform
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c);
          tmem_ld32(taddr, r);
          tmem_ld_wait(r);
          if (row_ok) store_chunk(p, r, has_k, row, col0, sp);
        }
        // release the accumulator buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// CTA-pair kernel: cluster (2,1,1), tcgen05 cta_group::2, pair tile 256 (M) x BLOCK_N.
// Each CTA stages its own 128 rows of A and its own half (BLOCK_N/2 rows) of B, so the pair
// moves (256 + BLOCK_N) x 64 operand elements per k-block for a 256 x BLOCK_N x 64 MMA: 1.5x
// fewer L2->SMEM bytes per flop than the single-CTA 128 x 256 tile, which the ncu capture of
// round 1 showed to be the limiter (profiles/gemm_r1_summary.md).  The leader CTA's elected
// thread issues the MMAs for both SMs; tcgen05.commit multicasts the stage-free and
// accumulator-ready arrivals to both CTAs; each CTA's epilogue drains its own 128 TMEM lanes.
// ---------------------------------------------------------------------------------------
template <int BLOCK_N>
struct Cfg2 {
  static constexpr int HALF_N = BLOCK_N / 2;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB (this CTA's 128 rows)
  static constexpr uint32_t B_BYTES = HALF_N * BLOCK_K * 2;    // this CTA's half of B
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;           // double-buffered accumulator
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BLOCK_N, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads_of(EPI), 1)
gemm2_kernel(const __grid_constant__ Params p) {
  using C = Cfg2<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 8u * (2 * C::STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  int total_kb = 0;
  for (int s = 0; s < p.n_segs; ++s) total_kb += p.seg_kblocks[s];
  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_tiles = m_pairs * p.n_tiles;
  const int num_work = num_tiles * p.split_k;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_segs; ++s) { prefetch_tmap(&p.tmap_a[s]); prefetch_tmap(&p.tmap_b[s]); }
  }
  if (warp == 1) {
    if (lane == 0) {
      // full: leader's own arrive.expect_tx + the peer producer's remote arrive
      for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
      // tmem_empty (leader's copy is the one used): 4 epilogue warps x 2 CTAs
      for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI ? 8 * GLU_WARPS_PER_QUARTER : 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  __syncthreads();        // (redundant with the cluster barrier; keeps racecheck's CTA model happy)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto work_range = [&](int w, int& tile, int& kb0, int& kb1) {
    tile = w % num_tiles;
    const int sp = w / num_tiles;
    const int per = (total_kb + p.split_k - 1) / p.split_k;
    kb0 = sp * per;
    kb1 = min(total_kb, kb0 + per);
  };
  auto tile_coords = [&](int tile, int& m_pair, int& n_blk) {
    raster_coords(tile, m_pairs, p.n_tiles, p.raster_mode, p.raster_group, m_pair, n_blk);
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) =======================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        int tile, kb0, kb1, m_pair, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_pair, n_blk);
        const int m0 = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M;     // this CTA's A rows
        const int n0 = n_blk * BLOCK_N + (int)rank * C::HALF_N;         // this CTA's B rows
        int seg = 0, seg_start = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          while (kb >= seg_start + p.seg_kblocks[seg]) { seg_start += p.seg_kblocks[seg]; ++seg; }
          const int k0 = (kb - seg_start) * BLOCK_K;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * C::STAGE_BYTES);
            else mbar_arrive_remote(full_bar(stage), 0u);
            if (!p.a_mn) {
              tma_load_2d_2sm(sa, &p.tmap_a[seg], full_bar(stage), k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d_2sm(sa + j * 8192u, &p.tmap_a[seg], full_bar(stage), m0 + j * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_2sm(sb, &p.tmap_b[seg], full_bar(stage), k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < C::HALF_N / 64; ++j)
                tma_load_2d_2sm(sb + j * 8192u, &p.tmap_b[seg], full_bar(stage), n0 + j * 64, k0);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ===================
    if (rank == 0) {
      const uint32_t idesc = make_idesc(2 * BLOCK_M, BLOCK_N, p.a_mn, p.b_mn, p.ab_fp16);
      const uint32_t a_adv = p.a_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
      const uint32_t b_adv = p.b_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        int tile, kb0, kb1;
        work_range(w, tile, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = make_smem_desc(sa, a_lbo, 1024u);
          const uint64_t db = make_smem_desc(sb, b_lbo, 1024u);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_f16_2sm(d_tmem, da + (uint64_t)(a_adv * k), db + (uint64_t)(b_adv * k), idesc,
                           (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(stage));      // frees the stage in BOTH CTAs
            if (kb == kb1 - 1) umma_commit_2sm(tfull_bar(acc));   // accumulator ready in BOTH CTAs
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) {
          if (elect_one()) umma_commit_2sm(tfull_bar(acc));
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue (both CTAs, own 128 lanes) ============
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI) {
      const int sub = (warp - 2) >> 2;
      auto coords = [&](int w, int& row, int& col_tile) {
        int tile, kb0, kb1, m_pair, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_pair, n_blk);
        row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
        col_tile = n_blk * BLOCK_N;
      };
      GluRegs gq0;
      int row, col_tile;
      if (pair < num_work) {
        coords(pair, row, col_tile);
        if (row < p.M) glu_prefetch(p, gq0, row, col_tile + 16 * sub);
      }
      for (int w = pair; w < num_work; w += num_pairs) {
        coords(w, row, col_tile);
        int nrow = -1, ncol = 0;
        if (w + num_pairs < num_work) {
          int nr, nc;
          coords(w + num_pairs, nr, nc);
          if (nr < p.M) { nrow = nr; ncol = nc + 16 * sub; }
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        glu_epilogue_tile<BLOCK_N>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N), gq0,
                                   row, row < p.M, col_tile, sub, nrow, ncol);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0u);   // leader's barrier
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else {
      for (int w = pair; w < num_work; w += num_pairs) {
        int tile, kb0, kb1, m_pair, n_blk;
        work_range(w, tile, kb0, kb1);
        tile_coords(tile, m_pair, n_blk);
        const int sp = w / num_tiles;
        const int row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + q * 32 + lane;
        const bool row_ok = row < p.M;
        const bool has_k = kb1 > kb0;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          const int col0 = n_blk * BLOCK_N + c;
          if (col0 >= p.N) break;
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c);
          tmem_ld32(taddr, r);
          tmem_ld_wait(r);
          if (row_ok) store_chunk(p, r, has_k, row, col0, sp);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0u);   // leader's barrier
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

// sum the fp32 split-K slices; optional 16-bit output
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws,
                                                            int64_t split_stride, int splits,
                                                            void* __restrict__ out, int64_t ldo,
                                                            int out_dtype, int accumulate, int M,
                                                            int N, int64_t ld_ws) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N, c = i - r * N;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(int64_t)s * split_stride + r * ld_ws + c];
    const int64_t o = r * ldo + c;
    if (out_dtype == UB200_F32) {
      float* p = reinterpret_cast<float*>(out);
      p[o] = accumulate ? p[o] + acc : acc;
    } else if (out_dtype == UB200_BF16) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out);
      p[o] = __float2bfloat16_rn(accumulate ? __bfloat162float(p[o]) + acc : acc);
    } else {
      __half* p = reinterpret_cast<__half*>(out);
      p[o] = __float2half_rn(accumulate ? __half2float(p[o]) + acc : acc);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI = 0>
static int launch(const Params& p, int grid, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BLOCK_N, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  gemm_kernel<BLOCK_N, EPI><<<grid, threads_of(EPI), C::SMEM_BYTES, st>>>(p);
  return UB200_OK;
}

template <int BLOCK_N, int EPI = 0>
static int launch2(const Params& p, int grid, cudaStream_t st) {
  using C = Cfg2<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<BLOCK_N, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  gemm2_kernel<BLOCK_N, EPI><<<grid, threads_of(EPI), C::SMEM_BYTES, st>>>(p);   // __cluster_dims__(2,1,1)
  return UB200_OK;
}

}  // namespace gemm
}  // namespace ub

extern "C" int ub200_gemm_workspace_bytes(int M, int N, int split_k, int64_t* bytes) {
  if (!bytes) return UB200_ERR_BAD_ARG;
  *bytes = split_k > 1 ? (int64_t)split_k * M * N * 4 : 0;
  return UB200_OK;
}

namespace ub {
namespace gemm {
struct GluArgs { int mode, act; void* e; void* g; int64_t ld_eg; };
}  // namespace gemm
}  // namespace ub

static int gemm_impl(int M, int N, const ub200_gemm_segment* segs, int n_segs, int a_mn_major,
                     int b_mn_major, int ab_dtype, void* C, int64_t ldc, int c_dtype,
                     float alpha, int accumulate, int split_k, void* workspace,
                     int block_n, int cta_group, const ub::gemm::GluArgs* glu, cudaStream_t stream) {
  using namespace ub;
  using namespace ub::gemm;
  if (M <= 0 || N <= 0) return UB200_OK;
  if (n_segs < 1 || n_segs > MAX_SEGS || !segs) return UB200_ERR_BAD_ARG;
  if (ab_dtype != UB200_BF16 && ab_dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !workspace) return UB200_ERR_BAD_ARG;
  int bn = block_n;
  if (bn == 0) bn = (N >= 256 || N > 128) ? 256 : (N > 64 ? 128 : 64);
  if (bn != 256 && bn != 128 && bn != 64) return UB200_ERR_BAD_ARG;
  if (cta_group != 0 && cta_group != 1 && cta_group != 2) return UB200_ERR_BAD_ARG;
  // CTA pairs pay off when the tile grid is large; the skinny rank-block GEMMs stay single-CTA
  bool pair = cta_group == 2 || (cta_group == 0 && bn >= 128 && M > BLOCK_M);
  if (pair && bn < 128) return UB200_ERR_BAD_ARG;

  Params p;
  memset(&p, 0, sizeof(p));
  if (glu) {
    p.glu_mode = glu->mode; p.glu_act = glu->act; p.glu_e = glu->e; p.glu_g = glu->g; p.ld_eg = glu->ld_eg;
  }
  p.n_segs = n_segs;
  p.M = M;
  p.N = N;
  p.a_mn = a_mn_major ? 1 : 0;
  p.b_mn = b_mn_major ? 1 : 0;
  p.ab_fp16 = ab_dtype == UB200_F16;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = (N + bn - 1) / bn;
  int total_kb = 0;
  for (int s = 0; s < n_segs; ++s) {
    const ub200_gemm_segment& g = segs[s];
    if (g.k <= 0) return UB200_ERR_BAD_ARG;
    p.seg_kblocks[s] = (int)((g.k + BLOCK_K - 1) / BLOCK_K);
    total_kb += p.seg_kblocks[s];
    int rc;
    if (!p.a_mn) rc = make_tmap(&p.tmap_a[s], g.a, M, g.k, g.lda, BLOCK_M, p.ab_fp16);
    else         rc = make_tmap(&p.tmap_a[s], g.a, g.k, M, g.lda, 64, p.ab_fp16);
    if (rc) return rc;
    if (!p.b_mn) rc = make_tmap(&p.tmap_b[s], g.b, N, g.k, g.ldb, pair ? bn / 2 : bn, p.ab_fp16);
    else         rc = make_tmap(&p.tmap_b[s], g.b, g.k, N, g.ldb, 64, p.ab_fp16);
    if (rc) return rc;
  }
  if (split_k > total_kb) split_k = total_kb;
  p.split_k = split_k;
  p.alpha = alpha;
  if (split_k > 1) {
    p.C = workspace;
    p.ldc = N;
    p.split_stride = (int64_t)M * N;
    p.c_dtype = UB200_F32;
    p.accumulate = 0;
  } else {
    p.C = C;
    p.ldc = ldc;
    p.split_stride = 0;
    p.c_dtype = c_dtype;
    p.accumulate = accumulate;
  }
  {
    // choose which operand stays L2-resident (see raster_coords)
    const int64_t k_total = (int64_t)total_kb * BLOCK_K;
    const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
    const int m_t = (M + tile_m - 1) / tile_m;
    const int64_t a_tile = (int64_t)tile_m * k_total * 2, b_tile = (int64_t)bn * k_total * 2;
    const int64_t a_total = (int64_t)M * k_total * 2, b_total = (int64_t)N * k_total * 2;
    const int64_t cap = 28ll << 20;
    int64_t gm = cap / a_tile; if (gm < 1) gm = 1; if (gm > m_t) gm = m_t;
    int64_t gn = cap / b_tile; if (gn < 1) gn = 1; if (gn > p.n_tiles) gn = p.n_tiles;
    const int64_t cost_a = a_total + b_total * ((m_t + gm - 1) / gm);
    const int64_t cost_b = b_total + a_total * ((p.n_tiles + gn - 1) / gn);
    if (cost_a <= cost_b) { p.raster_mode = 0; p.raster_group = (int)gm; }
    else { p.raster_mode = 1; p.raster_group = (int)gn; }
  }
  int rc;
  if (pair) {
    const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int num_work = m_pairs * p.n_tiles * split_k;
    int pairs = num_work < UB_SM_COUNT / 2 ? num_work : UB_SM_COUNT / 2;
    if (glu) rc = bn == 256 ? launch2<256, 1>(p, 2 * pairs, stream) : launch2<128, 1>(p, 2 * pairs, stream);
    else rc = bn == 256 ? launch2<256>(p, 2 * pairs, stream) : launch2<128>(p, 2 * pairs, stream);
  } else {
    const int num_work = p.m_tiles * p.n_tiles * split_k;
    const int grid = num_work < UB_SM_COUNT ? num_work : UB_SM_COUNT;
    if (glu) {
      if (bn == 256) rc = launch<256, 1>(p, grid, stream);
      else if (bn == 128) rc = launch<128, 1>(p, grid, stream);
      else rc = launch<64, 1>(p, grid, stream);
    } else if (bn == 256) rc = launch<256>(p, grid, stream);
    else if (bn == 128) rc = launch<128>(p, grid, stream);
    else rc = launch<64>(p, grid, stream);
  }
  if (rc) return rc;
  if (split_k > 1) {
    const int64_t total = (int64_t)M * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > UB_SM_COUNT * 8) blocks = UB_SM_COUNT * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>((const float*)workspace, (int64_t)M * N,
                                                     split_k, C, ldc, c_dtype, accumulate, M, N, N);
  }
  UB_RETURN_LAST();
}

extern "C" int ub200_gemm(int M, int N, const ub200_gemm_segment* segs, int n_segs, int a_mn_major,
                          int b_mn_major, int ab_dtype, void* C, int64_t ldc, int c_dtype,
                          float alpha, int accumulate, int split_k, void* workspace,
                          int block_n, int cta_group, cudaStream_t stream) {
  return gemm_impl(M, N, segs, n_segs, a_mn_major, b_mn_major, ab_dtype, C, ldc, c_dtype, alpha, accumulate,
                   split_k, workspace, block_n, cta_group, nullptr, stream);
}

extern "C" int ub200_gemm_glu(int mode, int act, int M, int N, const ub200_gemm_segment* segs, int n_segs,
                              int a_mn_major, int b_mn_major, int dtype, void* C, int64_t ldc, void* e,
                              void* g, int64_t ld_eg, float alpha, int block_n, int cta_group,
                              cudaStream_t stream) {
  if (M <= 0 || N <= 0) return UB200_OK;
  if (mode != UB200_GLU_EPI_FWD && mode != UB200_GLU_EPI_BWD) return UB200_ERR_BAD_ARG;
  if (act != ub::ACT_SWIGLU && act != ub::ACT_GEGLU_APPROX && act != ub::ACT_GEGLU_EXACT)
    return UB200_ERR_BAD_ARG;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  if (!C || !e || !g) return UB200_ERR_BAD_ARG;
  // the epilogue moves 32-byte vectors of 32-column chunks: every row of C / e / g must start 32-byte aligned
  if ((reinterpret_cast<uintptr_t>(C) & 31) || (reinterpret_cast<uintptr_t>(e) & 31) ||
      (reinterpret_cast<uintptr_t>(g) & 31) || (ldc % 16) || (ld_eg % 16) || ldc < N || ld_eg < N)
    return UB200_ERR_BAD_ARG;
  // whole tiles only and no scaling (the caller falls back to two launches otherwise)
  {
    int bn = block_n;
    if (bn == 0) bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
    if (bn != 256 && bn != 128 && bn != 64) return UB200_ERR_BAD_ARG;
    if (N % bn || alpha != 1.0f) return UB200_ERR_UNSUPPORTED;
  }
  ub::gemm::GluArgs ga{mode, act, e, g, ld_eg};
  return gemm_impl(M, N, segs, n_segs, a_mn_major, b_mn_major, dtype, C, ldc, dtype, alpha, 0, 1, nullptr,
                   block_n, cta_group, &ga, stream);
}
