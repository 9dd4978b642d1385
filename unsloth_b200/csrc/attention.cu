// Causal attention for the QLoRA training step on the 5th-generation tensor cores:
// tcgen05.mma with S = Q K^T and O = P V accumulating in TMEM, TMA-staged 128B-swizzled Q/K/V
// tiles, exp2 online softmax with LAZY rescaling of the TMEM-resident O (the running max is only
// advanced when it grows by more than 2^8, so O is almost never touched between tiles),
// sliding window (Mistral, Gemma-2 even layers), tanh soft-capping (Gemma-2), GQA and packed
// (varlen) rows.
//
// Replaces the external library call between RoPE and apply_o of the reference
// (unsloth/utils/attention_dispatch.py:298-617: flash-attn | xformers | SDPA;
//  models/mistral.py:112-157 window_size=(sw, sw); models/gemma2.py:139-199 softcap + window):
// flash-attn 2 is an mma.sync-era kernel with no TMA / tcgen05 / TMEM path.
//
// Layout: Q [B,S,Hq,D], K/V [B,S,Hk,D] are the projection buffers themselves (row stride H*D,
// head h at column h*D) -- no transposes, no copies; O [B,S,Hq,D] contiguous; LSE [B,Hq,S] fp32
// (natural log of the row sum of exp(scaled, soft-capped scores)), saved for the backward.
//
// Forward kernel anatomy (one CTA per (128-query tile, head, batch | document), 320 threads):
//   warp 0      TMA producer: Q once, then K / V tiles (2-stage ring)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S[j+1] = Q K[j+1]^T is issued BEFORE O += P[j] V[j], so the tensor core works on
//                 the next score tile while the softmax warps process the current one
//   warps 2..9  softmax: each thread owns one query row (= its TMEM lane) and half of the tile's
//                 columns: ONE tcgen05.ld of its scores, mask, running max (halves meet through smem),
//                 P = exp2(.) -> bf16 -> 128B-swizzled smem (the A operand of the PV MMA), final 1/l
//                 scaling of O and the LSE
// The same 128B-swizzled smem tile serves as a K-major operand (reduction over D) and as an
// MN-major operand (reduction over rows) just by the UMMA descriptor -- V is the MN-major B of PV.
//
// Roofline: tensor-bound in flops (4*S^2*D*Hq*B/2 causal) but in practice bounded by the MUFU
// exp2 rate (16 / clk / SM: 1024 clk per 128x128 tile against 1024 clk of MMA for D = 128).
#include <cstdlib>
#include <type_traits>

#include "tcgen05.cuh"

namespace ub {
namespace attn {

using namespace ub::gemm;

constexpr int BM = 128;            // query rows per CTA
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps
constexpr float LAZY_TAU = 8.0f;   // rescale O only when the running max (log2 domain) grows by > 8

struct FwdParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;   // 2-D [tokens, H*D]; boxes [BM | BN rows, 64 cols]
  void* O;                              // [tokens, Hq*D]
  float* lse;                           // [B, Hq, S]  (varlen: [Hq, tokens])
  const int32_t* cu_seqlens;            // nullable; [n_docs + 1]
  int64_t ldo;
  int B, S, Hq, Hk;
  int total_tokens;
  float scale_log2;                     // softmax_scale * log2(e)    (softcap == 0)
  float softcap, scale_over_cap, cap_log2;   // softcap > 0: t = cap_log2 * tanh(s * scale_over_cap)
  int window;                           // keys j >= i - window visible; < 0: unlimited
  int is_fp16;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): the softmax warps are ISSUE-bound next to the MMAs (about four
// instructions per score), one instruction per TWO scores for the scale-subtract and the row sum
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of element (row, col) inside a [rows x 64] 16-bit tile stored with the 128B swizzle
// (rows of 128 B; the 16-byte chunk index is XORed with row % 8) -- what TMA writes and UMMA reads
__device__ __forceinline__ uint32_t swz_off(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) + (((uint32_t)col & 7u) << 1);
}

template <int D, int BN, int PBUF>
struct FwdCfg {
  static constexpr int DB = D / 64;                                  // 64-wide column blocks of a row
  static constexpr uint32_t Q_BYTES = (uint32_t)DB * BM * 128;
  static constexpr uint32_t KV_BYTES = (uint32_t)DB * BN * 128;      // one K (or V) tile
  static constexpr uint32_t P_BYTES = (uint32_t)(BN / 64) * BM * 128;
  static constexpr int STAGES = 2;
  // D <= 128 (S = 2048: 16 query tiles of 8.5 key tiles on average): 64-key tiles + ONE P buffer = 112.6 KB and 256
  // TMEM columns, so TWO CTAs share an SM and the prologue / epilogue / hand-off latencies of one hide behind
  // the MMAs and exps of the other.  D = 256: one CTA per SM, P double-buffered.
  static constexpr uint32_t XCH_BYTES = 512;                         // row-max exchange: 2 halves x 128 rows x bf16
  static constexpr uint32_t SMEM_BYTES = Q_BYTES + STAGES * 2 * KV_BYTES + PBUF * P_BYTES + 136 + XCH_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static constexpr uint32_t TMEM_S0 = 0;                             // S buffers: 2 x BN columns
  static constexpr uint32_t TMEM_O = 2 * BN;                         // O: D columns
  static constexpr uint32_t TMEM_COLS = (2 * BN + D) <= 256 ? 256 : 512;
};

template <int D, int BN, int PBUF, int MINB>
__global__ void __launch_bounds__(NUM_THREADS, MINB) attn_fwd_kernel(const __grid_constant__ FwdParams p) {
  using C = FwdCfg<D, BN, PBUF>;
  // the swizzled tiles need 1024-byte alignment; the budget has no room for an align-up pad
  // (224 KB of tiles), so the dynamic segment is declared aligned and checked
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  uint8_t* smem_gen = smem_raw;
  const uint32_t q_smem = smem_base;
  const uint32_t kv_smem = q_smem + C::Q_BYTES;                      // stage s: K at +s*2*KV, V at +KV
  const uint32_t p_smem = kv_smem + C::STAGES * 2 * C::KV_BYTES;
  uint8_t* p_gen = smem_gen + (p_smem - smem_base);
  const uint32_t bar_base = p_smem + PBUF * C::P_BYTES;
  // barriers: q_full, k_full[2], v_full[2], v_empty[2], s_full[2], p_full[2], p_empty[2], o_full, k_empty[2]
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (3 + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (5 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (14 + s); };   // K is released right after QK (not after PV)
  auto s_full = [&](int s) { return bar_base + 8u * (7 + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (9 + s); };
  auto p_empty = [&](int s) { return bar_base + 8u * (11 + s); };
  const uint32_t o_full = bar_base + 8u * 13;
  const uint32_t tmem_ptr_smem = bar_base + 8u * 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t xch_smem = bar_base + 136u;                        // 2 halves x 128 rows x bf16 (row-max exchange); 16 barriers + tmem ptr = 132 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which rows ---------------------------------------------------------------------------
  const int head = blockIdx.y;
  const int kv_head = head / (p.Hq / p.Hk);
  int seq_start, seq_len;
  if (p.cu_seqlens) {
    seq_start = p.cu_seqlens[blockIdx.z];
    seq_len = p.cu_seqlens[blockIdx.z + 1] - seq_start;
  } else {
    seq_start = blockIdx.z * p.S;
    seq_len = p.S;
  }
  const int q_tiles = (seq_len + BM - 1) / BM;
  const int tile = (int)gridDim.x - 1 - (int)blockIdx.x;             // heaviest (last) tiles first
  if (tile >= q_tiles) return;
  const int m0 = tile * BM;
  // visible keys of this query tile: [lo, hi)
  const int hi = min(seq_len, m0 + BM);
  const int lo = p.window >= 0 ? max(0, m0 - p.window) : 0;
  const int j_begin = lo / BN, j_end = (hi + BN - 1) / BN;
  const int n_tiles = j_end - j_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_q); prefetch_tmap(&p.tmap_k); prefetch_tmap(&p.tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); mbar_init(k_empty(s), 1); mbar_init(s_full(s), 1);
      }
      for (int s = 0; s < 2; ++s) { mbar_init(p_full(s), 8); mbar_init(p_empty(s), 1); }   // one arrival per softmax warp
      mbar_init(o_full, 1);
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

  if (warp == 0) {
    // ================================ TMA producer ===================================
    if (elect_one()) {
      mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int b = 0; b < C::DB; ++b)
        tma_load_2d(q_smem + b * (BM * 128), &p.tmap_q, q_full, head * D + b * 64, seq_start + m0);
    }
    __syncwarp();
    // K is released right after its score MMA and prefetched ONE tile further ahead than V: the score MMA
    // of tile t+1 must never wait for a load that could only start after PV of tile t-1 (that chain
    // serialised MMA and softmax in the first version: 26 % tensor-pipe activity)
    auto load_k = [&](int t) {
      const int st = t & 1;
      mbar_wait(k_empty(st), (uint32_t)(((t >> 1) & 1) ^ 1));
      const uint32_t ks = kv_smem + st * 2 * C::KV_BYTES;
      if (elect_one()) {
        mbar_expect_tx(k_full(st), C::KV_BYTES);
#pragma unroll
        for (int b = 0; b < C::DB; ++b)
          tma_load_2d(ks + b * (BN * 128), &p.tmap_k, k_full(st), kv_head * D + b * 64, seq_start + (j_begin + t) * BN);
      }
      __syncwarp();
    };
    if (n_tiles > 0) load_k(0);
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t & 1;
      if (t + 1 < n_tiles) load_k(t + 1);
      mbar_wait(v_empty(st), (uint32_t)(((t >> 1) & 1) ^ 1));
      const uint32_t vs = kv_smem + st * 2 * C::KV_BYTES + C::KV_BYTES;
      if (elect_one()) {
        mbar_expect_tx(v_full(st), C::KV_BYTES);
#pragma unroll
        for (int b = 0; b < C::DB; ++b)
          tma_load_2d(vs + b * (BN * 128), &p.tmap_v, v_full(st), kv_head * D + b * 64, seq_start + (j_begin + t) * BN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =====================================
    const uint32_t idesc_qk = make_idesc(BM, BN, 0, 0, p.is_fp16);   // S = Q K^T: both K-major over D
    const uint32_t idesc_pv = make_idesc(BM, D, 0, 1, p.is_fp16);    // O += P V: V MN-major (rows = keys)
    auto issue_qk = [&](int t) {
      const int st = t & 1;
      mbar_wait(k_full(st), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint32_t ks = kv_smem + st * 2 * C::KV_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int b = 0; b < C::DB; ++b) {
          const uint64_t da = make_smem_desc(q_smem + b * (BM * 128), 16u, 1024u);
          const uint64_t db = make_smem_desc(ks + b * (BN * 128), 16u, 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + C::TMEM_S0 + (uint32_t)(st * BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                     idesc_qk, (b > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(s_full(st));
        umma_commit(k_empty(st));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    if (n_tiles > 0) issue_qk(0);
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t & 1;
      if (t + 1 < n_tiles) issue_qk(t + 1);
      const int pb = t % PBUF;
      mbar_wait(p_full(pb), (uint32_t)((t / PBUF) & 1));
      mbar_wait(v_full(st), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint32_t vs = kv_smem + st * 2 * C::KV_BYTES + C::KV_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < BN / 64; ++kb) {
          const uint64_t da = make_smem_desc(p_smem + pb * C::P_BYTES + kb * (BM * 128), 16u, 1024u);
          const uint64_t db = make_smem_desc(vs + kb * 8192u, (uint32_t)(BN * 128), 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + C::TMEM_O, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_pv,
                     (t > 0 || kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(v_empty(st));
        umma_commit(p_empty(pb));
        if (t == n_tiles - 1) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax / epilogue =============================
    // 8 warps: warps w and w+4 share TMEM lane quarter (w & 3) = 32 query rows; each thread owns one
    // row and HALF of the tile's columns (two warps per scheduler hide the MUFU / TMEM latencies).
    // The two partial row maxima meet through shared memory + a 64-thread named barrier.
    constexpr int HC = BN / 2;                    // score columns per thread
    constexpr int HD = D / 2;                     // O columns per thread
    const int sw = warp - 2;
    const int qd = warp & 3;
    const int ch = sw >> 2;                       // which half of the columns
    const int r = qd * 32 + lane;                 // query row of this thread inside the tile
    const int i_row = m0 + r;                     // position inside the sequence
    const uint32_t lane_addr = ((uint32_t)(qd * 32) << 16);
    uint16_t* xch = reinterpret_cast<uint16_t*>(smem_gen + (xch_smem - smem_base));   // [2 halves][128 rows] bf16
    float m_used = -INFINITY;                     // max the stored P / O are relative to (log2 domain)
    uint64_t l2 = pack2(0.f, 0.f);                // row sum, as two partial sums (even / odd columns)
    const bool capped = p.softcap > 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t & 1, pb = t % PBUF;
      const int n0 = (j_begin + t) * BN;
      mbar_wait(s_full(st), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const bool need_mask = (n0 + BN - 1 > m0) || (n0 + BN > seq_len) || (p.window >= 0 && n0 < m0 + BM - 1 - p.window);
      const int cb = ch * HC;                                   // first column of this thread
      const int c_hi = min(i_row, seq_len - 1) - n0 - cb;       // last visible column (relative to cb)
      const int c_lo = p.window >= 0 ? (i_row - p.window - n0 - cb) : -(1 << 30);
      const uint32_t s_addr = tmem_base + lane_addr + C::TMEM_S0 + (uint32_t)(st * BN + cb);
      // ---- one TMEM read of this thread's score columns -----------------------------------------
      float sv[HC];
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(s_addr + c0, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c0 + i] = __uint_as_float(v[i]);
      }
      // scores in the log2 domain: t = s * scale * log2(e)   (soft-capped: cap*log2(e) * tanh(s*scale/cap))
      if (capped) {
        const uint64_t soc2 = pack2(p.scale_over_cap, p.scale_over_cap), cl2 = pack2(p.cap_log2, p.cap_log2);
#pragma unroll
        for (int i = 0; i < HC; i += 2) {
          float a0, a1;
          unpack2(fmul2(pack2(sv[i], sv[i + 1]), soc2), a0, a1);
          unpack2(fmul2(pack2(tanh_approx(a0), tanh_approx(a1)), cl2), sv[i], sv[i + 1]);
        }
      }
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (i > c_hi || i < c_lo) sv[i] = -INFINITY;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < HC; ++i) mx = fmaxf(mx, sv[i]);
      if (!capped) mx *= p.scale_log2;                          // scale > 0: max commutes with it
      // ---- row maximum over both halves -----------------------------------------------------------
      // both partners use the bf16-rounded (upwards) values, so they agree on the reference maximum bit for bit
      {
        const __nv_bfloat16 own = __float2bfloat16_ru(mx);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");          // partner has read the previous tile's value
        xch[ch * 128 + r] = __bfloat16_as_ushort(own);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        mx = fmaxf(__bfloat162float(own), __bfloat162float(__ushort_as_bfloat16(xch[(ch ^ 1) * 128 + r])));
      }
      const float m_new = fmaxf(m_used, mx);
      // lazy rescale: advance the reference maximum only when it grew by more than 2^TAU (P stays
      // <= 2^TAU, harmless in fp32 / bf16), so O is rarely touched.  tcgen05.ld/st are warp-wide
      // (.sync.aligned): the decision is taken per WARP, rows that do not need it scale by 1.
      const bool want = (m_new > m_used + LAZY_TAU) || (m_used == -INFINITY && m_new > -INFINITY);
      const float m_next = want ? m_new : m_used;
      const float alpha = (m_used == -INFINITY) ? 1.0f : ex2f(m_used - m_next);
      const bool touch_o = t > 0 && want && m_used > -INFINITY;
      l2 = fmul2(l2, pack2(alpha, alpha));
      m_used = m_next;
      const float m_sub = m_used == -INFINITY ? 0.f : m_used;
      // ---- P = exp2(t - m) -> 16-bit, kept in registers until the previous PV MMA has retired ----
      const float mul = capped ? 1.0f : p.scale_log2;
      uint32_t pw[HC / 2];
      const uint64_t mul2 = pack2(mul, mul), nm2 = pack2(-m_sub, -m_sub);
      auto exps = [&](auto fp16_tag) {
#pragma unroll
        for (int i = 0; i < HC; i += 2) {
          float a0, a1;
          unpack2(ffma2(pack2(sv[i], sv[i + 1]), mul2, nm2), a0, a1);
          const float e0 = ex2f(a0), e1 = ex2f(a1);              // exp2(-inf) = 0 for masked entries
          l2 = fadd2(l2, pack2(e0, e1));
          if constexpr (decltype(fp16_tag)::value) { __half2 h = __floats2half2_rn(e0, e1); pw[i >> 1] = *reinterpret_cast<uint32_t*>(&h); }
          else { __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1); pw[i >> 1] = *reinterpret_cast<uint32_t*>(&h); }
        }
      };
      if (p.is_fp16) exps(std::true_type{}); else exps(std::false_type{});   // uniform branch: one F2FP per pair, not two predicated
      // this P buffer's previous user (the PV MMA of tile t - PBUF) must have retired
      mbar_wait(p_empty(pb), (uint32_t)(((t / PBUF) & 1) ^ 1));
      if (__any_sync(0xffffffffu, touch_o)) {
        // O is rescaled between PV(t-1) and PV(t): wait for PV(t-1)
        mbar_wait(p_empty((t - 1) % PBUF), (uint32_t)(((t - 1) / PBUF) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < HD; c0 += 32) {
          uint32_t v[32];
          const uint32_t o_addr = tmem_base + lane_addr + C::TMEM_O + (uint32_t)(ch * HD + c0);
          tmem_ld32(o_addr, v);
          tmem_ld_wait(v);
          const uint64_t al2 = pack2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            unpack2(fmul2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al2), a0, a1);
            v[i] = __float_as_uint(a0); v[i + 1] = __float_as_uint(a1);
          }
          tmem_st32(o_addr, v);
        }
        tmem_st_wait();
      }
#pragma unroll
      for (int g = 0; g < HC / 8; ++g) {
        const int col = cb + g * 8;                              // column inside the tile
        *reinterpret_cast<uint4*>(p_gen + pb * C::P_BYTES + (col >> 6) * (BM * 128) + swz_off(r, col & 63)) =
            make_uint4(pw[4 * g], pw[4 * g + 1], pw[4 * g + 2], pw[4 * g + 3]);
      }
      fence_proxy_async_smem();                   // generic-proxy smem writes -> tensor-core reads
      tc_fence_before();
      __syncwarp();                               // 256 arrivals on one mbarrier serialise: one per warp
      if (lane == 0) mbar_arrive(p_full(pb));
    }
    // ---- epilogue: O / l -> 16-bit rows, LSE ---------------------------------------------------
    if (n_tiles > 0) { mbar_wait(o_full, 0); tc_fence_after(); }
    float l;
    { float la, lb; unpack2(l2, la, lb); l = la + lb; }
    {
      // the two halves of a row add their partial sums through the (now idle) P region
      float* xb = reinterpret_cast<float*>(p_gen);
      xb[ch * 128 + r] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
      l += xb[(ch ^ 1) * 128 + r];
    }
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    const bool row_ok = i_row < seq_len;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + (int64_t)(seq_start + i_row) * p.ldo + (int64_t)head * D + ch * HD;
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v[32];
      if (n_tiles > 0) {
        tmem_ld32(tmem_base + lane_addr + C::TMEM_O + (uint32_t)(ch * HD + c0), v);
        tmem_ld_wait(v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = __uint_as_float(v[g * 8 + 2 * j]) * inv_l, b = __uint_as_float(v[g * 8 + 2 * j + 1]) * inv_l;
            if (p.is_fp16) { __half2 h = __floats2half2_rn(a, b); w[j] = *reinterpret_cast<uint32_t*>(&h); }
            else { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); w[j] = *reinterpret_cast<uint32_t*>(&h); }
          }
          *reinterpret_cast<uint4*>(orow + c0 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    if (row_ok && p.lse && ch == 0) {
      const float lse2 = (l > 0.f) ? m_used + lg2f(l) : -INFINITY;       // log2 domain
      const int64_t idx = p.cu_seqlens ? ((int64_t)head * p.total_tokens + seq_start + i_row)
                                       : (((int64_t)blockIdx.z * p.Hq + head) * p.S + i_row);
      p.lse[idx] = lse2 * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int make_tmap_rows(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                          int fp16) {
  return make_tmap(map, ptr, rows, cols, ld, box_rows, fp16);
}

template <int D, int BN, int PBUF, int MINB>
static int launch_fwd(const FwdParams& p, dim3 grid, cudaStream_t st) {
  using C = FwdCfg<D, BN, PBUF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<D, BN, PBUF, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  attn_fwd_kernel<D, BN, PBUF, MINB><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(p);
  return UB200_OK;
}

// =================================================================================================
// Backward.  P is recomputed from the saved LSE; no atomics anywhere (bit-reproducible):
//   delta_i = sum_d dO_i,d * O_i,d                                   (attn_delta_kernel)
//   dK_j, dV_j: one CTA per 128-key tile and KV head, loops over the query heads of the group and
//               over 64-query units:  S^T = K Q^T, dP^T = V dO^T  (keys in the TMEM lanes),
//               P^T = exp2(S^T - lse), dS^T = P^T (dP^T - delta) scale,  dV += P^T dO,  dK += dS^T Q
//   dQ_i:       one CTA per 128-query tile and head, loops over 64-key units:
//               S = Q K^T, dP = dO V^T, dS = P (dP - delta) scale,  dQ += dS K
// Both are the SAME kernel template: an OWNER tile of 128 rows (TMEM lanes) with two resident
// operands, a stream of 64-row units with two operands each, two "score" MMAs per unit into
// double-buffered TMEM, 8 compute warps turning them into 16-bit operand tiles in smem, and the
// accumulating MMAs that consume those tiles.  The score MMAs of unit u+1 are issued before the
// accumulating MMAs of unit u, so the tensor core is busy while the compute warps work.
// Every streamed operand tile is used both K-major (reduction over D) and MN-major (reduction
// over its rows) through two UMMA descriptors over the same 128B-swizzled bytes.
// =================================================================================================
constexpr int BU = 64;             // rows per streamed unit

struct BwdParams {
  CUtensorMap tmap_own1, tmap_own2;     // owner operands (dKV: K, V;  dQ: Q, dO), box [128 rows, 64 cols]
  CUtensorMap tmap_str1, tmap_str2;     // streamed operands (dKV: Q, dO;  dQ: K, V), box [64 rows, 64 cols]
  void* out1;                           // dKV: dK;  dQ: dQ
  void* out2;                           // dKV: dV;  dQ: unused
  int64_t ld_out1, ld_out2;
  const float* lse;
  const float* delta;
  const int32_t* cu_seqlens;
  int B, S, Hq, Hk, total_tokens;
  float scale, scale_log2, softcap, scale_over_cap, cap_log2;
  int window, is_fp16;
};

template <typename T>
__global__ void __launch_bounds__(256) attn_delta_kernel(const T* __restrict__ O, int64_t ldo,
                                                         const T* __restrict__ dO, int64_t lddo,
                                                         float* __restrict__ delta, const int32_t* cu, int B,
                                                         int S, int Hq, int D, int total_tokens) {
  // one warp per (token, head)
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)total_tokens * Hq;
  if (warp_id >= n) return;
  const int64_t tok = warp_id / Hq;
  const int h = (int)(warp_id - tok * Hq);
  const T* o = O + tok * ldo + (int64_t)h * D;
  const T* g = dO + tok * lddo + (int64_t)h * D;
  float acc = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + d);
    const uint4 b = *reinterpret_cast<const uint4*>(g + d);
    const T* av = reinterpret_cast<const T*>(&a);
    const T* bv = reinterpret_cast<const T*>(&b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += DT<T>::to_f(av[i]) * DT<T>::to_f(bv[i]);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    int64_t idx;
    if (cu) idx = (int64_t)h * total_tokens + tok;
    else { const int64_t b = tok / S; idx = ((b * Hq + h) * S) + (tok - b * S); }
    delta[idx] = acc;
  }
}

// MODE 0: dK and dV together (D <= 128)      owner K, V    stream Q, dO    tiles P^T, dS^T
// MODE 1: dQ                                   owner Q, dO   stream K, V     tile  dS
// MODE 2: dV only  (D = 256: dK + dV + the score tiles would need 768 TMEM columns)   tile P^T
// MODE 3: dK only  (D = 256)                                                          tile dS^T
enum { BWD_DKV = 0, BWD_DQ = 1, BWD_DV = 2, BWD_DK = 3 };

template <int D, int MODE>
struct BwdCfg {
  static constexpr int DB = D / 64;
  static constexpr bool KEYS_OWN = MODE != BWD_DQ;                   // owner rows are keys (lanes = keys)
  static constexpr int NOWN = MODE == BWD_DV ? 1 : 2;                // resident owner operands
  static constexpr int NKIND = MODE == BWD_DV ? 1 : 2;               // score MMAs per unit (S | S and dP)
  static constexpr int NT = MODE == BWD_DKV ? 2 : 1;                 // produced 16-bit tiles per unit
  static constexpr int NACC = MODE == BWD_DKV ? 2 : 1;               // accumulators
  static constexpr bool WANT_P = MODE == BWD_DKV || MODE == BWD_DV;
  static constexpr bool WANT_DS = MODE != BWD_DV;
  // D = 256: two resident [128 x 256] operands already take 128 KB -> single-stage stream / tile.
  // D <= 128: THREE stream stages -- a stage is only free once the accumulating MMAs of its unit have retired,
  // so with two stages the load of unit u+1 could not start before compute(u-1) + acc(u-1) had finished and
  // the score MMAs of u+1 (issued ahead of acc(u)) stalled the issuing thread on that load.
  static constexpr int STAGES = D == 256 ? (MODE == BWD_DV ? 2 : 1) : 3;
  static constexpr int TBUF = D == 256 ? 1 : 2;
  static constexpr uint32_t OWN_BYTES = (uint32_t)DB * 128 * 128;    // one owner operand [128 x D]
  static constexpr uint32_t STR_BYTES = (uint32_t)DB * BU * 128;     // one streamed operand [64 x D]
  static constexpr uint32_t T_BYTES = 128 * 128;                     // one 16-bit [128 x 64] operand tile
  static constexpr uint32_t SMEM_BYTES = NOWN * OWN_BYTES + STAGES * 2 * STR_BYTES + TBUF * NT * T_BYTES + 256 + 2048;
  static constexpr uint32_t TMEM_SC = 0;                             // score buffers: [2 units][2 kinds] x 64 columns
  static constexpr uint32_t TMEM_ACC = 256;                          // accumulators: NACC x D columns
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(256 + NACC * D <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int D, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) attn_bwd_kernel(const __grid_constant__ BwdParams p) {
  using C = BwdCfg<D, MODE>;
  constexpr bool DKV = C::KEYS_OWN;               // owner rows are keys, per-column scalars come from the queries
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // no room for an align-up pad: declared aligned, checked
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  uint8_t* smem_gen = smem_raw;
  const uint32_t own_smem = smem_base;                               // own1 | own2
  const uint32_t str_smem = own_smem + C::NOWN * C::OWN_BYTES;       // stage s: str1 | str2
  const uint32_t t_smem = str_smem + C::STAGES * 2 * C::STR_BYTES;   // buffer b: tile0 (| tile1)
  uint8_t* t_gen = smem_gen + (t_smem - smem_base);
  const uint32_t bar_base = t_smem + C::TBUF * C::NT * C::T_BYTES;
  // barriers: own_full, str_full[2], str_empty[2], sc_full[2], t_full[2], t_empty[2], out_full
  const uint32_t own_full = bar_base;
  auto str_full = [&](int s) { return bar_base + 8u * (1 + s); };     // up to 3 stages
  auto str_empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto sc_full = [&](int s) { return bar_base + 8u * (7 + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (9 + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (11 + s); };
  const uint32_t out_full = bar_base + 8u * 13;
  const uint32_t tmem_ptr_smem = bar_base + 8u * 14;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  float* vec = reinterpret_cast<float*>(smem_gen + (bar_base + 128u - smem_base));   // [2 buffers][2][64] lse2, delta

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- owner tile and its stream of units ---------------------------------------------------------
  const int own_head = blockIdx.y;                   // dKV: kv head;  dQ: q head
  const int G = p.Hq / p.Hk;
  int seq_start, seq_len;
  if (p.cu_seqlens) {
    seq_start = p.cu_seqlens[blockIdx.z];
    seq_len = p.cu_seqlens[blockIdx.z + 1] - seq_start;
  } else {
    seq_start = blockIdx.z * p.S;
    seq_len = p.S;
  }
  const int own_tiles = (seq_len + 127) / 128;
  const int tile = DKV ? (int)blockIdx.x : (int)gridDim.x - 1 - (int)blockIdx.x;
  if (tile >= own_tiles) return;
  const int o0 = tile * 128;                         // first owner row (key for dKV, query for dQ)
  int u_lo, u_hi;                                    // streamed units [u_lo, u_hi) of 64 rows
  if (DKV) {
    // queries i that see some key of [o0, o0+128): i >= o0, i <= o0 + 127 + window
    u_lo = o0 / BU;
    const int last = p.window >= 0 ? min(seq_len - 1, o0 + 127 + p.window) : seq_len - 1;
    u_hi = last / BU + 1;
  } else {
    // keys j visible to some query of [o0, o0+128): j <= o0 + 127, j >= o0 - window
    const int hi = min(seq_len, o0 + 128);
    u_lo = p.window >= 0 ? max(0, o0 - p.window) / BU : 0;
    u_hi = (hi + BU - 1) / BU;
  }
  const int units_per_head = max(0, u_hi - u_lo);
  const int n_units = DKV ? units_per_head * G : units_per_head;
  const int own_col1 = own_head * D;                 // column of the owner operands
  auto unit_head = [&](int u) { return DKV ? own_head * G + u / units_per_head : own_head; };   // q head of unit u
  auto unit_row0 = [&](int u) { return (u_lo + (DKV ? u % units_per_head : u)) * BU; };
  auto str_col = [&](int u) { return DKV ? unit_head(u) * D : (own_head / G) * D; };             // column of streamed operands

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_own1); prefetch_tmap(&p.tmap_own2); prefetch_tmap(&p.tmap_str1); prefetch_tmap(&p.tmap_str2);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(own_full, 1);
      for (int s = 0; s < 3; ++s) { mbar_init(str_full(s), 1); mbar_init(str_empty(s), 1); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(sc_full(s), 1);
        mbar_init(t_full(s), 8); mbar_init(t_empty(s), 1);        // one arrival per compute warp
      }
      mbar_init(out_full, 1);
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

  if (warp == 0) {
    // ================================ TMA producer ===================================
    if (elect_one()) {
      mbar_expect_tx(own_full, C::NOWN * C::OWN_BYTES);
#pragma unroll
      for (int b = 0; b < C::DB; ++b) {
        tma_load_2d(own_smem + b * (128 * 128), &p.tmap_own1, own_full, own_col1 + b * 64, seq_start + o0);
        if (C::NOWN == 2)
          tma_load_2d(own_smem + C::OWN_BYTES + b * (128 * 128), &p.tmap_own2, own_full, own_col1 + b * 64, seq_start + o0);
      }
    }
    __syncwarp();
    for (int u = 0; u < n_units; ++u) {
      const int st = u % C::STAGES;
      mbar_wait(str_empty(st), (uint32_t)(((u / C::STAGES) & 1) ^ 1));
      const uint32_t s1 = str_smem + st * 2 * C::STR_BYTES, s2 = s1 + C::STR_BYTES;
      const int col = str_col(u), row = seq_start + unit_row0(u);
      if (elect_one()) {
        mbar_expect_tx(str_full(st), 2 * C::STR_BYTES);
#pragma unroll
        for (int b = 0; b < C::DB; ++b) {
          tma_load_2d(s1 + b * (BU * 128), &p.tmap_str1, str_full(st), col + b * 64, row);
          tma_load_2d(s2 + b * (BU * 128), &p.tmap_str2, str_full(st), col + b * 64, row);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =====================================
    const uint32_t idesc_sc = make_idesc(128, BU, 0, 0, p.is_fp16);        // scores: both K-major over D
    const uint32_t idesc_acc = make_idesc(128, D, 0, 1, p.is_fp16);        // accumulate: tile K-major (64), streamed operand MN-major
    auto issue_scores = [&](int u) {
      const int st = u % C::STAGES, sb = u & 1;
      mbar_wait(str_full(st), (uint32_t)((u / C::STAGES) & 1));
      tc_fence_after();
      const uint32_t s1 = str_smem + st * 2 * C::STR_BYTES, s2 = s1 + C::STR_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kind = 0; kind < C::NKIND; ++kind) {
          const uint32_t a_base = own_smem + kind * C::OWN_BYTES, b_base = kind ? s2 : s1;
          const uint32_t d_tmem = tmem_base + C::TMEM_SC + (uint32_t)(sb * 128 + kind * 64);
#pragma unroll
          for (int b = 0; b < C::DB; ++b) {
            const uint64_t da = make_smem_desc(a_base + b * (128 * 128), 16u, 1024u);
            const uint64_t db = make_smem_desc(b_base + b * (BU * 128), 16u, 1024u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_sc, (b > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(sc_full(sb));
      }
      __syncwarp();
    };
    mbar_wait(own_full, 0);
    tc_fence_after();
    if (n_units > 0) issue_scores(0);
    for (int u = 0; u < n_units; ++u) {
      const int st = u % C::STAGES, tbi = u % C::TBUF;
      if (C::STAGES >= 2 && u + 1 < n_units) issue_scores(u + 1);     // next unit's scores ahead of this unit's accumulation
      mbar_wait(t_full(tbi), (uint32_t)((u / C::TBUF) & 1));
      tc_fence_after();
      const uint32_t s1 = str_smem + st * 2 * C::STR_BYTES, s2 = s1 + C::STR_BYTES;
      const uint32_t tb = t_smem + tbi * C::NT * C::T_BYTES;
      if (elect_one()) {
        if (MODE == BWD_DKV) {
          // dV += P^T (tile 0) x dO_u (str2, MN-major);  dK += dS^T (tile 1) x Q_u (str1, MN-major)
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const uint64_t da = make_smem_desc(tb + which * C::T_BYTES, 16u, 1024u);
            const uint64_t db = make_smem_desc(which == 0 ? s2 : s1, (uint32_t)(BU * 128), 1024u);
            const uint32_t d_tmem = tmem_base + C::TMEM_ACC + (uint32_t)((which == 0 ? 1 : 0) * D);   // dK at +0, dV at +D
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_acc, (u > 0 || k > 0) ? 1u : 0u);
          }
        } else {
          // dQ += dS x K_u (str1) | dV += P^T x dO_u (str2) | dK += dS^T x Q_u (str1): streamed operand MN-major
          const uint64_t da = make_smem_desc(tb, 16u, 1024u);
          const uint64_t db = make_smem_desc(MODE == BWD_DV ? s2 : s1, (uint32_t)(BU * 128), 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + C::TMEM_ACC, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_acc,
                     (u > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(str_empty(st));
        umma_commit(t_empty(tbi));
        if (u == n_units - 1) umma_commit(out_full);
      }
      __syncwarp();
      if (C::STAGES == 1 && u + 1 < n_units) issue_scores(u + 1);     // single stage: only after it has been refilled
    }
  } else {
    // ================================ compute warps ==================================
    const int sw = warp - 2;
    const int qd = warp & 3;
    const int ch = sw >> 2;                       // half of the unit's 64 columns
    const int r = qd * 32 + lane;                 // owner row (TMEM lane)
    const int cw = threadIdx.x - 64;              // 0..255 among the compute threads
    const uint32_t lane_addr = ((uint32_t)(qd * 32) << 16);
    const bool capped = p.softcap > 0.f;
    const int own_pos = o0 + r;                   // key position (dKV) / query position (dQ)
    auto vec_index = [&](int head, int pos) -> int64_t {
      return p.cu_seqlens ? ((int64_t)head * p.total_tokens + seq_start + pos)
                          : (((int64_t)blockIdx.z * p.Hq + head) * p.S + pos);
    };
    // dQ: the per-row scalars are fixed for the whole kernel
    float row_lse2 = 0.f, row_delta = 0.f;
    if (!DKV && own_pos < seq_len) {
      row_lse2 = p.lse[vec_index(own_head, own_pos)] * -1.4426950408889634f;       // -lse2
      row_delta = p.delta[vec_index(own_head, own_pos)] * p.scale;                 // delta * scale
    }
    // dKV: per-column scalars of unit u live in vec[u & 1]; the loader threads (cw < 128) fetch the
    // next unit's while the current one is processed
    float pre = 0.f;
    auto prefetch_vec = [&](int u) {
      if (DKV && cw < 128 && u < n_units) {
        const int c = cw & 63, pos = unit_row0(u) + c;
        const float* src = (cw < 64) ? p.lse : p.delta;
        pre = pos < seq_len ? src[vec_index(unit_head(u), pos)] : 0.f;
        pre *= (cw < 64) ? -1.4426950408889634f : p.scale;     // stored as -lse2 | delta * scale (what the packed math adds)
      }
    };
    if (DKV) {
      prefetch_vec(0);
      if (cw < 128) vec[cw] = pre;                // buffer 0: [-lse2 64 | delta*scale 64]
    }
    for (int u = 0; u < n_units; ++u) {
      const int st = u & 1;                       // score buffer / scalar buffer
      const int tbi = u % C::TBUF;
      if (DKV) {
        prefetch_vec(u + 1);
        asm volatile("bar.sync 5, 256;" ::: "memory");          // vec[st] visible; vec[st^1] free
      }
      const int u0 = unit_row0(u);
      mbar_wait(sc_full(st), (uint32_t)((u >> 1) & 1));
      tc_fence_after();
      mbar_wait(t_empty(tbi), (uint32_t)(((u / C::TBUF) & 1) ^ 1));   // this tile buffer's previous use consumed
      const int cb = ch * 32;
      uint32_t sraw[32], draw[32];
      tmem_ld32(tmem_base + lane_addr + C::TMEM_SC + (uint32_t)(st * 128 + cb), sraw);
      if (C::WANT_DS) tmem_ld32(tmem_base + lane_addr + C::TMEM_SC + (uint32_t)(st * 128 + 64 + cb), draw);
      tmem_ld_wait(sraw);
      if (C::WANT_DS) tmem_ld_wait(draw);
      // visibility: query i sees key j iff j <= i, i - j <= window, both inside the sequence
      const bool need_mask = DKV ? ((u0 < o0 + 127) || (p.window >= 0 && u0 + BU - 1 - o0 > p.window) ||
                                    (u0 + BU > seq_len) || (o0 + 128 > seq_len))
                                 : ((u0 + BU - 1 > o0) || (p.window >= 0 && o0 + 127 - u0 > p.window) ||
                                    (u0 + BU > seq_len) || (o0 + 128 > seq_len));
      const float* vb = vec + st * 128;
      uint32_t pw[16], dw[16];
      // P = exp2(t - lse2), dS = P * (dP - delta) * scale (* (1 - tanh^2) when soft-capped), two columns per
      // instruction (FFMA2 / FMUL2); dtype and capping are uniform branches around the whole loop
      const uint64_t sl2 = pack2(p.scale_log2, p.scale_log2), sc2 = pack2(p.scale, p.scale);
      const uint64_t row_nl2 = pack2(row_lse2, row_lse2), row_nd2 = pack2(-row_delta, -row_delta);
      auto elementwise = [&](auto fp16_tag, auto capped_tag) {
        constexpr bool FP16 = decltype(fp16_tag)::value, CAPPED = decltype(capped_tag)::value;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int c = cb + i;
          const float s0 = __uint_as_float(sraw[i]), s1 = __uint_as_float(sraw[i + 1]);
          uint64_t nl2 = row_nl2, nd2 = row_nd2;             // -lse2 | -(delta * scale) of the two columns
          if (DKV) {
            const float2 a = *reinterpret_cast<const float2*>(vb + c), b = *reinterpret_cast<const float2*>(vb + 64 + c);
            nl2 = pack2(a.x, a.y); nd2 = pack2(-b.x, -b.y);
          }
          float th0 = 0.f, th1 = 0.f, t0, t1;
          if (CAPPED) {
            th0 = tanh_approx(s0 * p.scale_over_cap); th1 = tanh_approx(s1 * p.scale_over_cap);
            unpack2(ffma2(pack2(th0, th1), pack2(p.cap_log2, p.cap_log2), nl2), t0, t1);
          } else {
            unpack2(ffma2(pack2(s0, s1), sl2, nl2), t0, t1);
          }
          float pe0 = ex2f(t0), pe1 = ex2f(t1);
          if (need_mask) {
            const int q0 = DKV ? u0 + c : own_pos, k0 = DKV ? own_pos : u0 + c;
            const int q1 = DKV ? q0 + 1 : q0, k1 = DKV ? k0 : k0 + 1;
            const bool v0 = (k0 <= q0) && (p.window < 0 || q0 - k0 <= p.window) && (q0 < seq_len) && (k0 < seq_len);
            const bool v1 = (k1 <= q1) && (p.window < 0 || q1 - k1 <= p.window) && (q1 < seq_len) && (k1 < seq_len);
            if (!v0) pe0 = 0.f;
            if (!v1) pe1 = 0.f;
          }
          float d0 = 0.f, d1 = 0.f;
          if (C::WANT_DS) {
            // (dP * scale - delta * scale) * P
            uint64_t x2 = ffma2(pack2(__uint_as_float(draw[i]), __uint_as_float(draw[i + 1])), sc2, nd2);
            x2 = fmul2(x2, pack2(pe0, pe1));
            if (CAPPED) x2 = fmul2(x2, pack2(1.0f - th0 * th0, 1.0f - th1 * th1));
            unpack2(x2, d0, d1);
          }
          if (FP16) {
            __half2 h = __floats2half2_rn(pe0, pe1); pw[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            __half2 g = __floats2half2_rn(d0, d1); dw[i >> 1] = *reinterpret_cast<uint32_t*>(&g);
          } else {
            __nv_bfloat162 h = __floats2bfloat162_rn(pe0, pe1); pw[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            __nv_bfloat162 g = __floats2bfloat162_rn(d0, d1); dw[i >> 1] = *reinterpret_cast<uint32_t*>(&g);
          }
        }
      };
      if (p.is_fp16) { if (capped) elementwise(std::true_type{}, std::true_type{}); else elementwise(std::true_type{}, std::false_type{}); }
      else { if (capped) elementwise(std::false_type{}, std::true_type{}); else elementwise(std::false_type{}, std::false_type{}); }
      uint8_t* tb = t_gen + tbi * C::NT * C::T_BYTES;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t off = swz_off(r, cb + g * 8);
        if (MODE == BWD_DKV) {
          *reinterpret_cast<uint4*>(tb + off) = make_uint4(pw[4 * g], pw[4 * g + 1], pw[4 * g + 2], pw[4 * g + 3]);
          *reinterpret_cast<uint4*>(tb + C::T_BYTES + off) = make_uint4(dw[4 * g], dw[4 * g + 1], dw[4 * g + 2], dw[4 * g + 3]);
        } else if (MODE == BWD_DV) {
          *reinterpret_cast<uint4*>(tb + off) = make_uint4(pw[4 * g], pw[4 * g + 1], pw[4 * g + 2], pw[4 * g + 3]);
        } else {
          *reinterpret_cast<uint4*>(tb + off) = make_uint4(dw[4 * g], dw[4 * g + 1], dw[4 * g + 2], dw[4 * g + 3]);
        }
      }
      if (DKV && cw < 128) vec[(st ^ 1) * 128 + cw] = pre;      // next unit's scalars
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_full(tbi));
    }
    // ---- epilogue: accumulators -> 16-bit rows -------------------------------------------------
    if (n_units > 0) { mbar_wait(out_full, 0); tc_fence_after(); }
    const bool row_ok = own_pos < seq_len;
    constexpr int HD = D / 2;
#pragma unroll 1
    for (int which = 0; which < C::NACC; ++which) {
      uint16_t* orow = reinterpret_cast<uint16_t*>(which == 0 ? p.out1 : p.out2) +
                       (int64_t)(seq_start + own_pos) * (which == 0 ? p.ld_out1 : p.ld_out2) + (int64_t)own_head * D + ch * HD;
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t v[32];
        if (n_units > 0) {
          tmem_ld32(tmem_base + lane_addr + C::TMEM_ACC + (uint32_t)(which * D + ch * HD + c0), v);
          tmem_ld_wait(v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = __uint_as_float(v[g * 8 + 2 * j]), b = __uint_as_float(v[g * 8 + 2 * j + 1]);
              if (p.is_fp16) { __half2 h = __floats2half2_rn(a, b); w[j] = *reinterpret_cast<uint32_t*>(&h); }
              else { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); w[j] = *reinterpret_cast<uint32_t*>(&h); }
            }
            *reinterpret_cast<uint4*>(orow + c0 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
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

template <int D, int MODE>
static int launch_bwd(const BwdParams& p, dim3 grid, cudaStream_t st) {
  using C = BwdCfg<D, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  attn_bwd_kernel<D, MODE><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(p);
  return UB200_OK;
}

}  // namespace attn
}  // namespace ub

extern "C" int ub200_attention_bwd(const void* dO, const void* Q, const void* K, const void* V, const void* O,
                                   const float* lse, float* delta, void* dQ, void* dK, void* dV,
                                   const int32_t* cu_seqlens, int n_docs, int max_seqlen, int batch, int seqlen,
                                   int n_heads_q, int n_heads_k, int head_dim, int64_t q_row_stride,
                                   int64_t k_row_stride, int64_t v_row_stride, int64_t do_row_stride,
                                   float softmax_scale, int window_left, float softcap, int dtype,
                                   cudaStream_t stream) {
  using namespace ub;
  using namespace ub::attn;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  if (head_dim != 64 && head_dim != 128 && head_dim != 256) return UB200_ERR_UNSUPPORTED;
  if (n_heads_q % n_heads_k) return UB200_ERR_BAD_ARG;
  const int64_t tokens = (int64_t)batch * seqlen;
  if (tokens <= 0) return UB200_OK;
  const int fp16 = dtype == UB200_F16;
  const int64_t ldq = (int64_t)n_heads_q * head_dim, ldk = (int64_t)n_heads_k * head_dim;
  // delta = rowsum(dO * O)   (O and dQ/dK/dV are contiguous [tokens, H*D])
  {
    const int64_t warps = tokens * n_heads_q;
    const int blocks = (int)((warps * 32 + 255) / 256);
    if (fp16) attn_delta_kernel<__half><<<blocks, 256, 0, stream>>>((const __half*)O, ldq, (const __half*)dO, do_row_stride,
                                                                     delta, cu_seqlens, batch, seqlen, n_heads_q, head_dim, (int)tokens);
    else attn_delta_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)O, ldq, (const __nv_bfloat16*)dO,
                                                                       do_row_stride, delta, cu_seqlens, batch, seqlen,
                                                                       n_heads_q, head_dim, (int)tokens);
  }
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.lse = lse; p.delta = delta; p.cu_seqlens = cu_seqlens;
  p.B = batch; p.S = seqlen; p.Hq = n_heads_q; p.Hk = n_heads_k; p.total_tokens = (int)tokens;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.softcap = softcap;
  if (softcap > 0.f) { p.scale_over_cap = softmax_scale / softcap; p.cap_log2 = softcap * 1.4426950408889634f; }
  p.window = window_left;
  p.is_fp16 = fp16;
  const int longest = cu_seqlens ? max_seqlen : seqlen;
  const int nb = cu_seqlens ? n_docs : batch;
  int rc;
  // ---- dK, dV: owner = K, V tiles; stream = Q, dO units ------------------------------------------
  if ((rc = make_tmap(&p.tmap_own1, K, tokens, ldk, k_row_stride, 128, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_own2, V, tokens, ldk, v_row_stride, 128, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_str1, Q, tokens, ldq, q_row_stride, BU, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_str2, dO, tokens, ldq, do_row_stride, BU, fp16))) return rc;
  p.out1 = dK; p.out2 = dV; p.ld_out1 = ldk; p.ld_out2 = ldk;
  {
    dim3 grid((longest + 127) / 128, n_heads_k, nb);
    if (head_dim == 256) {
      // dK + dV + score tiles would need 768 TMEM columns: two passes over the key tile
      p.out1 = dV;
      if ((rc = launch_bwd<256, BWD_DV>(p, grid, stream))) return rc;
      p.out1 = dK;
      rc = launch_bwd<256, BWD_DK>(p, grid, stream);
    } else {
      rc = head_dim == 64 ? launch_bwd<64, BWD_DKV>(p, grid, stream) : launch_bwd<128, BWD_DKV>(p, grid, stream);
    }
    if (rc) return rc;
  }
  // ---- dQ: owner = Q, dO tiles; stream = K, V units ------------------------------------------------
  if ((rc = make_tmap(&p.tmap_own1, Q, tokens, ldq, q_row_stride, 128, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_own2, dO, tokens, ldq, do_row_stride, 128, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_str1, K, tokens, ldk, k_row_stride, BU, fp16))) return rc;
  if ((rc = make_tmap(&p.tmap_str2, V, tokens, ldk, v_row_stride, BU, fp16))) return rc;
  p.out1 = dQ; p.out2 = nullptr; p.ld_out1 = ldq; p.ld_out2 = 0;
  {
    dim3 grid((longest + 127) / 128, n_heads_q, nb);
    rc = head_dim == 64 ? launch_bwd<64, BWD_DQ>(p, grid, stream)
         : head_dim == 128 ? launch_bwd<128, BWD_DQ>(p, grid, stream) : launch_bwd<256, BWD_DQ>(p, grid, stream);
    if (rc) return rc;
  }
  UB_RETURN_LAST();
}

extern "C" int ub200_attention_fwd(const void* Q, const void* K, const void* V, void* O, float* lse,
                                   const int32_t* cu_seqlens, int n_docs, int max_seqlen, int batch,
                                   int seqlen, int n_heads_q, int n_heads_k, int head_dim,
                                   int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride,
                                   int64_t o_row_stride, float softmax_scale, int window_left,
                                   float softcap, int dtype, cudaStream_t stream) {
  using namespace ub;
  using namespace ub::attn;
  if (dtype != UB200_BF16 && dtype != UB200_F16) return UB200_ERR_BAD_ARG;
  if (head_dim != 64 && head_dim != 128 && head_dim != 256) return UB200_ERR_UNSUPPORTED;
  if (n_heads_q % n_heads_k) return UB200_ERR_BAD_ARG;
  const int64_t tokens = (int64_t)batch * seqlen;
  if (tokens <= 0) return UB200_OK;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  const int fp16 = dtype == UB200_F16;
  const int fwd_mode = [] { const char* e = getenv("UB200_ATTN_FWD"); return e ? atoi(e) : 2; }();   // 1: 128-key tiles, 1 CTA/SM; 2: 64-key tiles, 2 CTAs/SM
  const int BN = (head_dim == 256 || fwd_mode == 2) ? 64 : 128;
  int rc;
  if ((rc = make_tmap_rows(&p.tmap_q, Q, tokens, (int64_t)n_heads_q * head_dim, q_row_stride, BM, fp16))) return rc;
  if ((rc = make_tmap_rows(&p.tmap_k, K, tokens, (int64_t)n_heads_k * head_dim, k_row_stride, BN, fp16))) return rc;
  if ((rc = make_tmap_rows(&p.tmap_v, V, tokens, (int64_t)n_heads_k * head_dim, v_row_stride, BN, fp16))) return rc;
  p.O = O; p.lse = lse; p.cu_seqlens = cu_seqlens; p.ldo = o_row_stride;
  p.B = batch; p.S = seqlen; p.Hq = n_heads_q; p.Hk = n_heads_k;
  p.total_tokens = (int)tokens;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.softcap = softcap;
  if (softcap > 0.f) { p.scale_over_cap = softmax_scale / softcap; p.cap_log2 = softcap * 1.4426950408889634f; }
  p.window = window_left;
  p.is_fp16 = fp16;
  const int longest = cu_seqlens ? max_seqlen : seqlen;
  dim3 grid((longest + BM - 1) / BM, n_heads_q, cu_seqlens ? n_docs : batch);
  if (head_dim == 256) rc = launch_fwd<256, 64, 2, 1>(p, grid, stream);
  else if (fwd_mode == 2) rc = head_dim == 64 ? launch_fwd<64, 64, 1, 2>(p, grid, stream) : launch_fwd<128, 64, 1, 2>(p, grid, stream);
  else rc = head_dim == 64 ? launch_fwd<64, 128, 2, 1>(p, grid, stream) : launch_fwd<128, 128, 2, 1>(p, grid, stream);
  if (rc) return rc;
  UB_RETURN_LAST();
}
