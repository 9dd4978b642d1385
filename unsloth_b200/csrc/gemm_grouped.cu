// Grouped multi-segment GEMM: every GEMM of one phase of a LoRA projection group -- the dense
// tcgen05 GEMMs AND the rank-block products around them -- in ONE persistent launch.
//
// The reference issues, per projection group, one cuBLAS call per term (unsloth/kernels/
// fast_lora.py:116-229, 432-540, 617-650; kernels/utils.py:1128-1170): X@A, (XA)@B added onto X@W,
// and in the backward dY@B, X^T@(...), (XA)^T@dY, dY@W, ... -- 26 GEMM launches per decoder layer of
// which 19 are rank-r ("skinny") products that each re-read a full [T, in] or [T, out] activation
// and cannot fill 148 SMs.  Round 1 fused the terms of each OUTPUT into one launch but still ran the
// 19 skinny products as separate launches (8.4 % of the step at 0.19-0.37 of HBM bandwidth, each
// paying ~13 us of fixed launch / pipeline-fill / drain cost).
//
// Here a launch takes a LIST of problems  C_p[M_p,N_p] (+)= alpha_p * sum_s A_ps . B_ps^T  and one
// persistent grid of CTA pairs walks the concatenated tile list:
//   * the rank-block producers (XA = X @ A_cat^T in the forward, G = sum_i dY_i @ sB_i in the
//     backward) are ordinary tiles at the FRONT of the list; a dense tile that needs them as its
//     last K-segment (Y = X W^T + XA B^T,  dX = sum_i dY_i W_i + G A_cat) waits, in its TMA
//     producer warp only, on a per-row-block completion counter in global memory -- by the time a
//     dense tile has streamed its 64-450 weight k-blocks the producer tile has long finished;
//   * the LoRA gradients dB_i = s dY_i^T XA and dA = X^T G (reductions over the T tokens) are
//     split-K tiles of the same launch -- "dA/dB in-kernel" -- whose fp32 partials are reduced by the
//     LAST split to arrive, summing in fixed split order (bit-reproducible run to run);
//   * dA waits for ALL tiles of G (a whole-problem counter).
// Deadlock freedom: the grid is persistent with every CTA pair resident, each pair walks its work
// items in increasing index, and a wait only ever targets a problem placed EARLIER in the list, so
// the chain of waits ends at a tile that waits for nothing.  The flag/counter scratch is
// self-cleaning (the last CTA to finish zeroes it), so launches -- and CUDA-graph replays -- need
// no memset.
//
// Tile = CTA pair, 256 rows x BN columns (BN 64/128/256 per problem), tcgen05 cta_group::2,
// TMA 128B-swizzled operands K-major or MN-major, 6-stage mbarrier ring, double-buffered TMEM
// accumulators -- the round-1 gemm2_kernel pipeline with per-work-item problem decoding.
//
// Roofline: tensor-bound for the dense problems (flops = 2 M N sum K); the rank-block problems ride
// along inside the same launch and are bounded by the L2->SMEM operand stream of their K loop.
#include "tcgen05.cuh"

namespace ub {
namespace gemm {
namespace grouped {

constexpr int BLOCK_M = 128;          // per CTA; a pair covers 256 rows
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MAX_PROBS = UB200_GROUPED_MAX_PROBLEMS;
constexpr int MAX_SEGS = UB200_GROUPED_MAX_SEGMENTS;
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB: this CTA's 128 rows of A
constexpr uint32_t B_BYTES_MAX = 128 * BLOCK_K * 2;        // 16 KB: this CTA's half of a 256-wide B tile
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int STAGES = 6;
constexpr uint32_t TMEM_COLS = 512;                        // 2 x 256 accumulator columns
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Prob {
  CUtensorMap tmap_a[MAX_SEGS];
  CUtensorMap tmap_b[MAX_SEGS];
  int seg_kblocks[MAX_SEGS];
  int n_segs, total_kb;
  int M, N;
  int a_mn, b_mn;
  int block_n;             // 64 / 128 / 256
  int m_pairs, n_tiles;    // tile grid (256 x block_n tiles)
  int split_k;
  int work_begin;          // first global work index of this problem; work = tiles * split_k
  int raster_mode, raster_group;
  void* C;
  int64_t ldc;
  int c_dtype, accumulate;
  float alpha;
  float* ws;               // fp32 [split_k, M, N] when split_k > 1
  int cnt_base;            // split-K arrival counters: scratch[cnt_base + tile*8 + slice]
  int flag_base;           // completion counters: scratch[flag_base + m_pair], total at [flag_base + m_pairs]
  int signals;             // bump the completion counters when an output tile is final
  int wait_prob, wait_seg, wait_all;   // dependency (wait_prob < 0: none)
};

struct Params {
  Prob probs[MAX_PROBS];
  int n_probs;
  int total_work;
  int ab_fp16;
  int* scratch;            // zero on entry; zeroed again by the last CTA to leave
  int n_scratch;           // ints to clean
};

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

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

struct Out {          // where an output chunk goes
  void* C;
  int64_t ldc;
  int c_dtype, accumulate, N;
};

// 32 consecutive columns of one row (values already scaled): optional beta = 1, dtype conversion
__device__ __forceinline__ void store_row32(const Out& o, float (&v)[32], int row, int col0) {
  const bool full = (col0 + 32 <= o.N);
  if (o.c_dtype == UB200_F32) {
    float* cptr = reinterpret_cast<float*>(o.C) + (int64_t)row * o.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (o.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(cptr + i);
          t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
        }
        *reinterpret_cast<float4*>(cptr + i) = t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)          // fully unrolled: v[] stays in registers
        if (col0 + i < o.N) cptr[i] = o.accumulate ? cptr[i] + v[i] : v[i];
    }
  } else {
    uint16_t* cptr = reinterpret_cast<uint16_t*>(o.C) + (int64_t)row * o.ldc + col0;
    const bool bf = o.c_dtype == UB200_BF16;
    if (full && ((reinterpret_cast<uintptr_t>(cptr) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 q;
        if (o.accumulate) {
          const uint4 old = *reinterpret_cast<const uint4*>(cptr + i);
          const uint16_t* oh = reinterpret_cast<const uint16_t*>(&old);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[i + j] += bf ? __bfloat162float(__ushort_as_bfloat16(oh[j])) : __half2float(__ushort_as_half(oh[j]));
        }
        uint32_t* ow = reinterpret_cast<uint32_t*>(&q);
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
        *reinterpret_cast<uint4*>(cptr + i) = q;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (col0 + i < o.N) {
          float val = v[i];
          if (o.accumulate)
            val += bf ? __bfloat162float(__ushort_as_bfloat16(cptr[i])) : __half2float(__ushort_as_half(cptr[i]));
          cptr[i] = bf ? __bfloat16_as_ushort(__float2bfloat16_rn(val)) : __half_as_ushort(__float2half_rn(val));
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_grouped_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
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

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_probs; ++i)
      for (int s = 0; s < p.probs[i].n_segs; ++s) {
        prefetch_tmap(&p.probs[i].tmap_a[s]);
        prefetch_tmap(&p.probs[i].tmap_b[s]);
      }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
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

  // work item -> (problem, tile, split range)
  auto decode = [&](int w, int& pi, int& tile, int& sp, int& kb0, int& kb1) {
    pi = 0;
#pragma unroll 1
    for (int i = 1; i < p.n_probs; ++i)
      if (w >= p.probs[i].work_begin) pi = i;
    const Prob& q = p.probs[pi];
    const int local = w - q.work_begin;
    const int tiles = q.m_pairs * q.n_tiles;
    sp = local / tiles;
    tile = local - sp * tiles;
    const int per = (q.total_kb + q.split_k - 1) / q.split_k;
    kb0 = sp * per;
    kb1 = min(q.total_kb, kb0 + per);
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) =======================
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair; w < p.total_work; w += num_pairs) {
      int pi, tile, sp, kb0, kb1, m_pair, n_blk;
      decode(w, pi, tile, sp, kb0, kb1);
      const Prob& q = p.probs[pi];
      raster_coords(tile, q.m_pairs, q.n_tiles, q.raster_mode, q.raster_group, m_pair, n_blk);
      const int half_n = q.block_n >> 1;
      const int m0 = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M;
      const int n0 = n_blk * q.block_n + (int)rank * half_n;
      const uint32_t stage_tx = A_BYTES + (uint32_t)half_n * BLOCK_K * 2;
      bool waited = q.wait_prob < 0;
      const int a_mn = q.a_mn, b_mn = q.b_mn, wait_seg = q.wait_seg;
      // segment boundaries in registers: the k loop must not chase the parameter bank
      int seg_end[MAX_SEGS];
      {
        int acc_kb = 0;
#pragma unroll
        for (int s = 0; s < MAX_SEGS; ++s) { acc_kb += (s < q.n_segs) ? q.seg_kblocks[s] : 0; seg_end[s] = acc_kb; }
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        int seg = 0, seg_start = 0;
#pragma unroll
        for (int s = 0; s < MAX_SEGS - 1; ++s)
          if (kb >= seg_end[s]) { seg = s + 1; seg_start = seg_end[s]; }
        if (!waited && seg >= wait_seg) {
          // the operand of this segment is produced by an earlier problem of THIS launch
          const Prob& d = p.probs[q.wait_prob];
          const int* flag = p.scratch + d.flag_base + (q.wait_all ? d.m_pairs : m_pair);
          const int target = 8 * (q.wait_all ? d.m_pairs * d.n_tiles : d.n_tiles);
          while (ld_acquire(flag) < target) __nanosleep(64);
          fence_proxy_async_all();          // generic-proxy writes of the producer -> our TMA reads
          __syncwarp();
          waited = true;
        }
        const int k0 = (kb - seg_start) * BLOCK_K;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * stage_tx);
          else mbar_arrive_remote(full_bar(stage), 0u);
          if (!a_mn) {
            tma_load_2d_2sm(sa, &q.tmap_a[seg], full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d_2sm(sa + j * 8192u, &q.tmap_a[seg], full_bar(stage), m0 + j * 64, k0);
          }
          if (!b_mn) {
            tma_load_2d_2sm(sb, &q.tmap_b[seg], full_bar(stage), k0, n0);
          } else {
            for (int j = 0; j < half_n / 64; ++j)
              tma_load_2d_2sm(sb + j * 8192u, &q.tmap_b[seg], full_bar(stage), n0 + j * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ===================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair; w < p.total_work; w += num_pairs) {
        int pi, tile, sp, kb0, kb1;
        decode(w, pi, tile, sp, kb0, kb1);
        const Prob& q = p.probs[pi];
        const uint32_t idesc = make_idesc(2 * BLOCK_M, q.block_n, q.a_mn, q.b_mn, p.ab_fp16);
        const uint32_t a_adv = q.a_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
        const uint32_t b_adv = q.b_mn ? (UMMA_K * 128u) >> 4 : (UMMA_K * 2u) >> 4;
        const uint32_t a_lbo = q.a_mn ? 8192u : 16u, b_lbo = q.b_mn ? 8192u : 16u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa, a_lbo, 1024u);
          const uint64_t db = make_smem_desc(sb, b_lbo, 1024u);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_f16_2sm(d_tmem, da + (uint64_t)(a_adv * k), db + (uint64_t)(b_adv * k), idesc,
                           (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar(stage));
            if (kb == kb1 - 1) umma_commit_2sm(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
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
    const int qd = warp & 3;                      // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair; w < p.total_work; w += num_pairs) {
      int pi, tile, sp, kb0, kb1, m_pair, n_blk;
      decode(w, pi, tile, sp, kb0, kb1);
      const Prob& q = p.probs[pi];
      raster_coords(tile, q.m_pairs, q.n_tiles, q.raster_mode, q.raster_group, m_pair, n_blk);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_pair * 2 * BLOCK_M + (int)rank * BLOCK_M + qd * 32 + lane;
      const bool row_ok = row < q.M;
      const bool has_k = kb1 > kb0;
      const bool split = q.split_k > 1;
      Out o;
      if (split) { o.C = q.ws + (int64_t)sp * q.M * q.N; o.ldc = q.N; o.c_dtype = UB200_F32; o.accumulate = 0; }
      else { o.C = q.C; o.ldc = q.ldc; o.c_dtype = q.c_dtype; o.accumulate = q.accumulate; }
      o.N = q.N;
#pragma unroll 1
      for (int c = 0; c < q.block_n; c += 32) {
        const int col0 = n_blk * q.block_n + c;
        if (col0 >= q.N) break;                   // warp-uniform
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * 256 + c);
        tmem_ld32(taddr, r);
        tmem_ld_wait(r);
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) * q.alpha : 0.f;
          store_row32(o, v, row, col0);
        }
      }
      // release the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0u);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }

      bool final_here = !split;
      if (split) {
        // Last split to arrive for this 32-row slice reduces all partials in FIXED split order.
        const int slice = (int)rank * 4 + qd;
        int* cnt = p.scratch + q.cnt_base + tile * 8 + slice;
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(cnt, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == q.split_k - 1) {
          __threadfence();
          if (lane == 0) *cnt = 0;                // ready for the next launch
          Out f;
          f.C = q.C; f.ldc = q.ldc; f.c_dtype = q.c_dtype; f.accumulate = q.accumulate; f.N = q.N;
          if (row_ok) {
#pragma unroll 1
            for (int c = 0; c < q.block_n; c += 32) {
              const int col0 = n_blk * q.block_n + c;
              if (col0 >= q.N) break;
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
              const int ncol = min(32, q.N - col0);
              for (int s = 0; s < q.split_k; ++s) {
                const float* src = q.ws + ((int64_t)s * q.M + row) * q.N + col0;
                if (ncol == 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
                  for (int i = 0; i < 32; i += 4) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(src + i));
                    v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (i < ncol) v[i] += __ldcg(src + i);
                }
              }
              store_row32(f, v, row, col0);
            }
          }
          final_here = true;
        }
      }
      if (final_here && q.signals) {
        // publish this warp's rows: consumers read them through TMA (async proxy)
        __threadfence();
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) {
          red_release_add(p.scratch + q.flag_base + m_pair, 1);
          red_release_add(p.scratch + q.flag_base + q.m_pairs, 1);
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
  // self-cleaning scratch: the last CTA to leave zeroes the completion counters (and the exit
  // counter itself, kept at scratch[n_scratch])
  if (threadIdx.x == 0 && p.n_scratch > 0) {
    __threadfence();
    const int old = atomicAdd(p.scratch + p.n_scratch, 1);
    if (old == (int)gridDim.x - 1) {
      for (int i = 0; i < p.n_scratch; ++i) p.scratch[i] = 0;
      p.scratch[p.n_scratch] = 0;
      __threadfence();
    }
  }
}

}  // namespace grouped
}  // namespace gemm
}  // namespace ub

extern "C" int ub200_gemm_grouped_scratch_ints(const ub200_gemm_problem* probs, int n_probs, int* ints) {
  using namespace ub::gemm::grouped;
  if (!probs || !ints || n_probs < 1 || n_probs > MAX_PROBS) return UB200_ERR_BAD_ARG;
  int n = 0;
  for (int i = 0; i < n_probs; ++i) {
    const ub200_gemm_problem& g = probs[i];
    const int bn = g.block_n;
    if (bn != 64 && bn != 128 && bn != 256) return UB200_ERR_BAD_ARG;
    const int m_pairs = (g.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int n_tiles = (g.N + bn - 1) / bn;
    n += m_pairs + 1;
    if (g.split_k > 1) n += m_pairs * n_tiles * 8;
  }
  *ints = n + 1;      // + the exit counter (layout: flags | exit counter | split-K arrival counters)
  return UB200_OK;
}

extern "C" int ub200_gemm_grouped(const ub200_gemm_problem* probs, int n_probs, int ab_dtype,
                                  int* scratch, cudaStream_t stream) {
  using namespace ub;
  using namespace ub::gemm;
  using namespace ub::gemm::grouped;
  if (!probs || n_probs < 1 || n_probs > MAX_PROBS || !scratch) return UB200_ERR_BAD_ARG;
  if (ab_dtype != UB200_BF16 && ab_dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  static thread_local Params p;      // ~10 KB: keep it off the stack of the caller's thread
  memset(&p, 0, sizeof(p));
  p.n_probs = n_probs;
  p.ab_fp16 = ab_dtype == UB200_F16;
  p.scratch = scratch;
  int work = 0, nsc = 0;
  for (int i = 0; i < n_probs; ++i) {
    const ub200_gemm_problem& g = probs[i];
    Prob& q = p.probs[i];
    if (g.M <= 0 || g.N <= 0 || g.n_segs < 1 || g.n_segs > MAX_SEGS || !g.segs || !g.C) return UB200_ERR_BAD_ARG;
    const int bn = g.block_n;
    if (bn != 64 && bn != 128 && bn != 256) return UB200_ERR_BAD_ARG;
    if (g.b_mn_major && bn < 128) return UB200_ERR_BAD_ARG;   // MN-major B: 64-wide atoms per CTA half
    q.n_segs = g.n_segs;
    q.M = g.M; q.N = g.N;
    q.a_mn = g.a_mn_major ? 1 : 0;
    q.b_mn = g.b_mn_major ? 1 : 0;
    q.block_n = bn;
    q.m_pairs = (g.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    q.n_tiles = (g.N + bn - 1) / bn;
    int total_kb = 0;
    for (int s = 0; s < g.n_segs; ++s) {
      const ub200_gemm_segment& sg = g.segs[s];
      if (sg.k <= 0) return UB200_ERR_BAD_ARG;
      q.seg_kblocks[s] = (int)((sg.k + BLOCK_K - 1) / BLOCK_K);
      total_kb += q.seg_kblocks[s];
      int rc;
      if (!q.a_mn) rc = make_tmap(&q.tmap_a[s], sg.a, g.M, sg.k, sg.lda, BLOCK_M, p.ab_fp16);
      else         rc = make_tmap(&q.tmap_a[s], sg.a, sg.k, g.M, sg.lda, 64, p.ab_fp16);
      if (rc) return rc;
      if (!q.b_mn) rc = make_tmap(&q.tmap_b[s], sg.b, g.N, sg.k, sg.ldb, bn / 2, p.ab_fp16);
      else         rc = make_tmap(&q.tmap_b[s], sg.b, sg.k, g.N, sg.ldb, 64, p.ab_fp16);
      if (rc) return rc;
    }
    q.total_kb = total_kb;
    int split = g.split_k < 1 ? 1 : g.split_k;
    if (split > total_kb) split = total_kb;
    if (split > 1 && !g.workspace) return UB200_ERR_BAD_ARG;
    q.split_k = split;
    q.ws = reinterpret_cast<float*>(g.workspace);
    q.C = g.C; q.ldc = g.ldc; q.c_dtype = g.c_dtype; q.accumulate = g.accumulate; q.alpha = g.alpha;
    q.signals = g.signals ? 1 : 0;
    q.wait_prob = g.wait_problem;
    q.wait_seg = g.wait_segment;
    q.wait_all = g.wait_all ? 1 : 0;
    if (q.wait_prob >= i) return UB200_ERR_BAD_ARG;            // waits only on EARLIER problems
    if (q.wait_prob >= 0) {
      const ub200_gemm_problem& d = probs[q.wait_prob];
      if (!d.signals) return UB200_ERR_BAD_ARG;
      if (!q.wait_all && !q.a_mn && ((d.M + 255) / 256) != q.m_pairs) return UB200_ERR_BAD_ARG;
      if (q.wait_seg < 0 || q.wait_seg >= g.n_segs) return UB200_ERR_BAD_ARG;
      if (q.wait_all == 0 && q.a_mn) return UB200_ERR_BAD_ARG;  // row-block waits need A rows = tokens
    }
    q.flag_base = nsc;
    nsc += q.m_pairs + 1;
    q.work_begin = work;
    work += q.m_pairs * q.n_tiles * split;
    {
      const int64_t k_total = (int64_t)total_kb * BLOCK_K;
      const int64_t a_tile = (int64_t)2 * BLOCK_M * k_total * 2, b_tile = (int64_t)bn * k_total * 2;
      const int64_t a_total = (int64_t)g.M * k_total * 2, b_total = (int64_t)g.N * k_total * 2;
      const int64_t cap = 28ll << 20;
      int64_t gm = cap / a_tile; if (gm < 1) gm = 1; if (gm > q.m_pairs) gm = q.m_pairs;
      int64_t gn = cap / b_tile; if (gn < 1) gn = 1; if (gn > q.n_tiles) gn = q.n_tiles;
      const int64_t cost_a = a_total + b_total * ((q.m_pairs + gm - 1) / gm);
      const int64_t cost_b = b_total + a_total * ((q.n_tiles + gn - 1) / gn);
      if (cost_a <= cost_b) { q.raster_mode = 0; q.raster_group = (int)gm; }
      else { q.raster_mode = 1; q.raster_group = (int)gn; }
    }
  }
  p.total_work = work;
  p.n_scratch = nsc;            // completion flags [0, nsc) + exit counter at [nsc]; cleaned by the last CTA
  int cnt = nsc + 1;            // split-K arrival counters follow (each reducer resets its own)
  for (int i = 0; i < n_probs; ++i) {
    Prob& q = p.probs[i];
    q.cnt_base = cnt;
    if (q.split_k > 1) cnt += q.m_pairs * q.n_tiles * 8;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int pairs = work < UB_SM_COUNT / 2 ? work : UB_SM_COUNT / 2;
  gemm_grouped_kernel<<<2 * pairs, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  UB_RETURN_LAST();
}
