/*
 * unsloth_b200 -- C ABI of the B200 (sm_100a) fused-kernel QLoRA fine-tuning hot path.
 *
 * The reference (unslothai/unsloth @ 015bdef) has no C ABI on this path: its boundary is a set
 * of Python callables (unsloth/kernels/__init__.py:15-62) whose bodies launch Triton kernels,
 * cuBLAS GEMMs and three bitsandbytes C symbols bound through ctypes
 * (unsloth/kernels/utils.py:273-284).  This header declares what a ctypes binding for the path
 * binds instead; each entry point cites the reference function whose device work it replaces.
 *
 * Conventions: the caller owns every buffer (outputs pre-allocated by the
 * caller, e.g. torch.empty); all pointers are device pointers unless stated; functions never
 * allocate, never synchronise and only enqueue work on `stream` (pass the CURRENT stream, not one
 * cached at import); no thread-local or global mutable state, so forward and autograd's backward
 * thread may call concurrently.  Return value: 0 (UB200_OK), a negative UB200_ERR_* code, or a
 * positive cudaError_t from the launch.  dtype codes: UB200_F32 / UB200_F16 / UB200_BF16.
 * Row strides are in ELEMENTS.
 */
#ifndef UNSLOTH_B200_H_
#define UNSLOTH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define UB200_OK 0
#define UB200_ERR_BAD_ARG (-1)
#define UB200_ERR_UNSUPPORTED (-2)
#define UB200_ERR_NO_DRIVER (-3) /* cuTensorMapEncodeTiled could not be resolved */
#define UB200_ERR_TMAP (-4)      /* tensor-map encoding rejected the operand */

#define UB200_F32 0
#define UB200_F16 1
#define UB200_BF16 2

#define UB200_ACT_SWIGLU 0
#define UB200_ACT_GEGLU_APPROX 1
#define UB200_ACT_GEGLU_EXACT 2

#define UB200_GEMM_MAX_SEGMENTS 8

/* ABI version of this header (bumped on any signature change). */
int ub200_abi_version(void);

/* ---- RMSNorm --------------------------------------------------------------------------------
 * fast_rms_layernorm / Fast_RMS_Layernorm (unsloth/kernels/rms_layernorm.py:162-255).
 * fwd: Y = (X * rsqrt(mean(X^2) + eps)).to(W.dtype) * W   (gemma: fp32, * (1 + W)); r[T] fp32.
 * bwd: dX only (weights frozen); dX may alias dY (the reference's in-place contract).       */
int ub200_rms_layernorm_fwd(const void* X, int64_t x_row_stride, const void* W, int w_dtype,
                            void* Y, int64_t y_row_stride, float* r, int64_t n_rows, int n_cols,
                            float eps, int gemma, int dtype, cudaStream_t stream);
int ub200_rms_layernorm_bwd(const void* dY, int64_t dy_row_stride, const void* X,
                            int64_t x_row_stride, const void* W, int w_dtype, const float* r,
                            void* dX, int64_t dx_row_stride, int64_t n_rows, int n_cols, int gemma,
                            int dtype, cudaStream_t stream);

/* Residual add fused with the following RMSNorm (models/llama.py:838-844 does `residual +
 * hidden_states` and fast_rms_layernorm as two passes): S = A + B, Y = RMSNorm(S) * W, r[T].
 * Backward companion: dS += rms_bwd(dY, S, W, r), accumulated in place into the residual-stream
 * gradient.  16-bit activations with weights of the same dtype (Llama / Mistral form).           */
int ub200_add_rms_layernorm_fwd(const void* A, int64_t a_row_stride, const void* B,
                                int64_t b_row_stride, const void* W, void* S, int64_t s_row_stride,
                                void* Y, int64_t y_row_stride, float* r, int64_t n_rows, int n_cols,
                                float eps, int dtype, cudaStream_t stream);
int ub200_rms_layernorm_bwd_acc(const void* dY, int64_t dy_row_stride, const void* X,
                                int64_t x_row_stride, const void* W, const float* r, void* dS,
                                int64_t ds_row_stride, int64_t n_rows, int n_cols, int dtype,
                                cudaStream_t stream);

/* ---- RoPE -----------------------------------------------------------------------------------
 * fast_rope_embedding (unsloth/kernels/rope_embedding.py:265-280): Fast_RoPE_Embedding (:169-261)
 * and Fast_RoPE_Embedding_QK (:283-399) as ONE strided in-place kernel over Q and K.
 * Element (b,h,s,d) of Q is at Q + b*q_batch_stride + h*q_head_stride + s*q_seq_stride + d.
 * K may be NULL.  indices: int32 [batch*seqlen] row of cos/sin per token, or NULL (row = s).
 * backward != 0 negates sin.  compute_dtype: dtype in which products/sums are rounded (table
 * dtype for the no-index form, promoted dtype for the QK form).                              */
int ub200_rope_qk(void* Q, int64_t q_batch_stride, int64_t q_head_stride, int64_t q_seq_stride,
                  void* K, int64_t k_batch_stride, int64_t k_head_stride, int64_t k_seq_stride,
                  const void* cos, int64_t cos_row_stride, const void* sin, int64_t sin_row_stride,
                  const int32_t* indices, int batch, int seqlen, int n_heads_q, int n_heads_k,
                  int head_dim, int backward, int dtype, int table_dtype, int compute_dtype,
                  cudaStream_t stream);

/* ---- SwiGLU / GEGLU -------------------------------------------------------------------------
 * swiglu_fg_kernel / swiglu_DWf_DW_dfg_kernel (unsloth/kernels/swiglu.py:50-64, 112-125),
 * geglu_{approx,exact}_{forward,backward}_kernel (unsloth/kernels/geglu.py).
 * fwd: h = act(e).to(dtype) * g.  bwd (in place): DW <- h, e <- df = DW*f, g <- de.           */
int ub200_glu_fwd(int act, const void* e, const void* g, void* h, int64_t n, int dtype,
                  cudaStream_t stream);
int ub200_glu_bwd(int act, void* DW, void* e, void* g, int64_t n, int dtype, cudaStream_t stream);

/* ---- cross entropy on materialised logits ---------------------------------------------------
 * Fast_CrossEntropyLoss (unsloth/kernels/cross_entropy_loss.py:288-418); one path for any
 * vocabulary size (the reference splits at 65,536).  labels int64, -100 = ignore.
 * fwd writes loss[T], logsumexp[T] (fp32).  bwd overwrites logits with
 * dloss[row*dloss_stride] * d loss / d logits.  softcap / scale == 0 disables them.          */
int ub200_cross_entropy_fwd(const void* logits, int64_t row_stride, const int64_t* labels,
                            float* loss, float* lse, int64_t n_rows, int vocab, float softcap,
                            float scale, int dtype, cudaStream_t stream);
int ub200_cross_entropy_bwd(void* logits, int64_t row_stride, const float* lse,
                            const int64_t* labels, const float* dloss, int64_t dloss_stride,
                            int64_t n_rows, int vocab, float softcap, float scale, int dtype,
                            cudaStream_t stream);

/* ---- NF4 -------------------------------------------------------------------------------------
 * fast_dequantize (unsloth/kernels/utils.py:567-679).  One launch for both stages of the
 * double-quantised format: absmax = code2[absmax_q[i]] * absmax2[i / blocksize2] + *offset,
 * out[2j] = NF4[packed[j] >> 4] * absmax[2j / blocksize], out[2j+1] = NF4[packed[j] & 15] * ...
 * `offset` is a DEVICE scalar (may be NULL = 0).  n = number of weights.                       */
int ub200_dequantize_nf4(const uint8_t* packed, const uint8_t* absmax_q, const float* code2,
                         const float* absmax2, const float* offset, void* out, int64_t n,
                         int blocksize, int blocksize2, int out_dtype, cudaStream_t stream);
/* Blockwise NF4 quantiser (blocksize 64): packed[n/2], absmax[n/64] fp32 (first level only).  */
int ub200_quantize_nf4(const void* W, int dtype, uint8_t* packed, float* absmax, int64_t n,
                       int blocksize, cudaStream_t stream);
/* The bitsandbytes symbols the reference binds (unsloth/kernels/utils.py:273-284, call sites
 * :650-675) with their exact signatures: void return, errors surface at the next sync.        */
void cdequantize_blockwise_fp32(float* code, unsigned char* A, float* absmax, float* out,
                                int blocksize, const int n, cudaStream_t stream);
#ifdef __CUDACC__
void cdequantize_blockwise_bf16_nf4(float* code, unsigned char* A, float* absmax,
                                    __nv_bfloat16* out, int blocksize, const int n,
                                    cudaStream_t stream);
void cdequantize_blockwise_fp16_nf4(float* code, unsigned char* A, float* absmax, __half* out,
                                    int blocksize, const int n, cudaStream_t stream);
#else
void cdequantize_blockwise_bf16_nf4(float* code, unsigned char* A, float* absmax, void* out,
                                    int blocksize, const int n, cudaStream_t stream);
void cdequantize_blockwise_fp16_nf4(float* code, unsigned char* A, float* absmax, void* out,
                                    int blocksize, const int n, cudaStream_t stream);
#endif

/* ---- decode-time GEMV ---------------------------------------------------------
 * out[m] = W[m,k] . x[k] with W NF4-packed, expanded in registers: the launch behind `fast_gemv`
 * (unsloth/kernels/utils.py:874-973) and the q_len == 1 branch of `fast_linear_forward`
 * (:1082-1125).  absmax either ready in fp32 (absmax_f32 != NULL, the bitsandbytes contract) or
 * rebuilt in-kernel from the double-quantised statistics (absmax_q, code2, absmax2, offset):
 * this folds the reference's cdequantize_blockwise_fp32 launch and `+= offset` (:938-948).
 * code16: the 16-entry NF4 table (`quant_state.code`), NULL = built-in.  Optional LoRA epilogue
 * out[row] += s * lora_B[row,:r] . lora_t[:r] (lora_t = A x in fp32; :1108-1112).
 * x, out, lora_B in `dtype` (BF16 / F16); k % 32 == 0, blocksize % 32 == 0.                     */
int ub200_gemv_nf4(const void* x, const uint8_t* packed, const float* absmax_f32,
                   const uint8_t* absmax_q, const float* code2, const float* absmax2,
                   const float* offset, const float* code16, void* out, int m, int k,
                   int blocksize, int blocksize2, const void* lora_B, int ldb, const float* lora_t,
                   int r, float s, int dtype, cudaStream_t stream);
/* 16-bit dense rows: out[m] = W[m,k] . x[k] (`torch.mv(lm_head, h)`, models/llama.py:1460; the
 * LoRA temp A x).  out_dtype F32 or `dtype`.                                                    */
int ub200_gemv_dense(const void* x, const void* W, int64_t ldw, void* out, int m, int k, int dtype,
                     int out_dtype, cudaStream_t stream);
/* bitsandbytes symbols bound at unsloth/kernels/utils.py:283-284 (call :955-973): A = x[k],
 * B = packed [m, k/2], absmax fp32, datatype = 16-entry code, n == 1; void return.              */
#ifdef __CUDACC__
void cgemm_4bit_inference_naive_bf16(int m, int n, int k, __nv_bfloat16* A, unsigned char* B,
                                     float* absmax, float* datatype, __nv_bfloat16* out, int lda,
                                     int ldb, int ldc, int blocksize, cudaStream_t stream);
void cgemm_4bit_inference_naive_fp16(int m, int n, int k, __half* A, unsigned char* B,
                                     float* absmax, float* datatype, __half* out, int lda, int ldb,
                                     int ldc, int blocksize, cudaStream_t stream);
#else
void cgemm_4bit_inference_naive_bf16(int m, int n, int k, void* A, unsigned char* B, float* absmax,
                                     float* datatype, void* out, int lda, int ldb, int ldc,
                                     int blocksize, cudaStream_t stream);
void cgemm_4bit_inference_naive_fp16(int m, int n, int k, void* A, unsigned char* B, float* absmax,
                                     float* datatype, void* out, int lda, int ldb, int ldc,
                                     int blocksize, cudaStream_t stream);
#endif

/* ---- tcgen05 GEMM ---------------------------------------------------------------------------
 * The primitive under matmul_lora (unsloth/kernels/utils.py:1128-1170) and the dX / dA / dB
 * GEMMs of LoRA_MLP / LoRA_QKV / LoRA_W.backward (unsloth/kernels/fast_lora.py:116-229,
 * 432-540, 617-650):
 *      C[M,N] = alpha * sum_s A_s . B_s^T   (+ C if accumulate)
 * Every segment contributes K_s to the reduction; all accumulate in fp32 in TMEM.
 *   a_mn_major == 0: A_s is row-major [M, K_s] (lda);  != 0: A_s is row-major [K_s, M].
 *   b_mn_major == 0: B_s is row-major [N, K_s] (ldb);  != 0: B_s is row-major [K_s, N].
 * Operands bf16 (or fp16), 16-byte aligned, ld multiple of 8.  C: bf16/fp16/fp32, ldc elements.
 * split_k > 1 needs `workspace` of ub200_gemm_workspace_bytes(); reduction order is fixed
 * (deterministic).  block_n: 0 = auto, or 64 / 128 / 256.  cta_group: 0 = auto, 1 = one CTA
 * per 128 x block_n tile, 2 = CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x block_n
 * tile (block_n >= 128).                                                                       */
typedef struct {
  const void* a;
  int64_t lda;
  const void* b;
  int64_t ldb;
  int64_t k;
} ub200_gemm_segment;
int ub200_gemm(int M, int N, const ub200_gemm_segment* segs, int n_segs, int a_mn_major,
               int b_mn_major, int ab_dtype, void* C, int64_t ldc, int c_dtype, float alpha,
               int accumulate, int split_k, void* workspace, int block_n, int cta_group,
               cudaStream_t stream);
int ub200_gemm_workspace_bytes(int M, int N, int split_k, int64_t* bytes);

/* ---- tcgen05 GEMM with the gated activation in its epilogue -----------------------------------
 * The two places where LoRA_MLP runs an elementwise kernel straight after a projection
 * (unsloth/kernels/fast_lora.py:84-87 forward, :155-157 backward) as ONE launch: the fp32
 * accumulator tile is rounded to `dtype` exactly where the reference's matmul output is, then the
 * per-element code of ub200_glu_fwd / ub200_glu_bwd runs on it in registers (same bits as the
 * two-launch form).  Operand conventions as ub200_gemm; no split-K, no accumulate; `dtype` (bf16 /
 * fp16) is the operand AND output dtype; C, e, g 32-byte aligned, ldc and ld_eg multiples of 16, N a multiple of the tile width (block_n, auto: 256 / 128 / 64 as for ub200_gemm) and alpha == 1, else UB200_ERR_UNSUPPORTED.
 *   mode UB200_GLU_EPI_FWD: acc = up projection.   g <- acc,  C <- act(e).to(dtype) * g   (e read only:
 *                           the gate projection written by the previous launch)
 *   mode UB200_GLU_EPI_BWD: acc = DW = dY @ W_down (+ LoRA).  C <- h = f(e) * g,  e <- df = DW * f,
 *                           g <- de  (swiglu.py:86-109, geglu.py:74-123, 214-244: DW / e / g reuse)    */
#define UB200_GLU_EPI_FWD 1
#define UB200_GLU_EPI_BWD 2
int ub200_gemm_glu(int mode, int act, int M, int N, const ub200_gemm_segment* segs, int n_segs,
                   int a_mn_major, int b_mn_major, int dtype, void* C, int64_t ldc, void* e, void* g,
                   int64_t ld_eg, float alpha, int block_n, int cta_group, cudaStream_t stream);

/* ---- grouped tcgen05 GEMM: one persistent launch for a whole phase of a LoRA projection group --
 * All GEMMs the reference issues for one phase of LoRA_MLP / LoRA_QKV / LoRA_W (forward: X@A, the
 * dense products with the rank update; backward: dY@B, the dX sum, and the dA / dB reductions over
 * tokens -- unsloth/kernels/fast_lora.py:116-229, 432-540, 617-650) as a LIST of problems
 *      C_p[M_p,N_p] = alpha_p * sum_s A_ps . B_ps^T   (+ C_p if accumulate)
 * walked by one persistent grid (csrc/gemm_grouped.cu).  Operand conventions as ub200_gemm.
 *   block_n      64 / 128 / 256 output columns per tile (MN-major B needs >= 128; N may be smaller
 *                than block_n: TMA zero-fills the missing columns)
 *   split_k > 1  fp32 `workspace` [split_k, M, N]; the last split to arrive reduces in fixed order
 *   signals      consumers in the same launch may wait on this problem's output
 *   wait_problem index (< own index) of a problem of this launch that PRODUCES an operand of this
 *                one, or -1; wait_segment: first segment that reads it.  wait_all == 0: the producer's
 *                rows [256 m, 256 m + 256) are needed by this problem's row block m (the operand is
 *                a K-major A with the same M); wait_all != 0: the whole producer output is needed
 *                (it is a reduction operand, e.g. dA = X^T @ G).
 * `scratch`: ub200_gemm_grouped_scratch_ints() int32s, ZERO before the first use; the kernel
 * leaves it zero again (CUDA-graph replays need no memset).  One scratch buffer per stream.     */
#define UB200_GROUPED_MAX_PROBLEMS 8
#define UB200_GROUPED_MAX_SEGMENTS 4
typedef struct {
  int M, N;
  const ub200_gemm_segment* segs;
  int n_segs;
  int a_mn_major, b_mn_major;
  void* C;
  int64_t ldc;
  int c_dtype;
  float alpha;
  int accumulate;
  int split_k;
  void* workspace;
  int block_n;
  int signals;
  int wait_problem, wait_segment, wait_all;
} ub200_gemm_problem;
int ub200_gemm_grouped(const ub200_gemm_problem* probs, int n_probs, int ab_dtype, int* scratch,
                       cudaStream_t stream);
int ub200_gemm_grouped_scratch_ints(const ub200_gemm_problem* probs, int n_probs, int* ints);

/* ---- NF4 dequantisation fused into the tcgen05 GEMM's operand staging ---------------------------
 * Y[M,N] = X[M,K] . dequant(W)[N,K]^T (+ lora_xa[M,lora_k] . lora_b[N,lora_k]^T): matmul_lora
 * (unsloth/kernels/utils.py:1128-1170) without the 16-bit copy of W that fast_dequantize (:567-679)
 * writes first.  W: bitsandbytes NF4 double-quant fields as in ub200_dequantize_nf4; blocksize 64 /
 * 256 and K % 64 == 0 (one quantisation block per row and k-block).  lora_k: 0 or a multiple of 64
 * (zero-padded rank block, lora_b already scaled).  Bit-identical to dequantise-then-ub200_gemm.  */
int ub200_gemm_nf4(int M, int N, int K, const void* X, int64_t ldx, const uint8_t* packed,
                   const uint8_t* absmax_q, const float* code2, const float* absmax2,
                   const float* offset, int blocksize, int blocksize2, const void* lora_xa,
                   int64_t ld_xa, const void* lora_b, int64_t ld_b, int lora_k, void* C, int64_t ldc,
                   int dtype, cudaStream_t stream);

/* ---- causal attention (tcgen05 / TMEM / TMA) ---------------------------------------------------
 * The attention product between fast_rope_embedding and apply_o, which the reference delegates to
 * flash-attn / xformers / SDPA (unsloth/utils/attention_dispatch.py:298-617; sliding window
 * models/mistral.py:112-157; soft-capping + window models/gemma2.py:139-199).
 * Q [tokens, Hq*D], K / V [tokens, Hk*D]: rows of the projection buffers (row strides in elements,
 * head h at column h*D; tokens = batch*seqlen), O [tokens, Hq*D], lse fp32 [batch, Hq, seqlen]
 * (varlen: [Hq, tokens]) = natural-log row sums, needed by the backward (may be NULL).
 * Causal always.  window_left >= 0: key j visible to query i iff i - window_left <= j <= i (the
 * reference's flash-attn window_size=(w, w) under causal masking); < 0: unlimited.  softcap > 0:
 * scores = softcap * tanh(scale * qk / softcap).  cu_seqlens != NULL (int32 [n_docs+1], device):
 * packed rows, batch must be 1, attention is block-diagonal per document.  D in {64, 128, 256}.  */
int ub200_attention_fwd(const void* Q, const void* K, const void* V, void* O, float* lse,
                        const int32_t* cu_seqlens, int n_docs, int max_seqlen, int batch, int seqlen,
                        int n_heads_q, int n_heads_k, int head_dim, int64_t q_row_stride,
                        int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                        float softmax_scale, int window_left, float softcap, int dtype,
                        cudaStream_t stream);

/* Backward of ub200_attention_fwd: dQ [tokens, Hq*D], dK / dV [tokens, Hk*D] (contiguous) from dO, the
 * forward's O (contiguous [tokens, Hq*D]) and lse.  `delta` is caller-provided fp32 scratch shaped
 * like lse.  P is recomputed; dK/dV are accumulated per key tile over all query heads of the group
 * and dQ per query tile, both in TMEM -- no atomics, run-to-run bit-identical.  D in {64, 128, 256}
 * (D = 256: dV and dK in two passes, since dK + dV + score tiles exceed the 512 TMEM columns).     */
int ub200_attention_bwd(const void* dO, const void* Q, const void* K, const void* V, const void* O,
                        const float* lse, float* delta, void* dQ, void* dK, void* dV,
                        const int32_t* cu_seqlens, int n_docs, int max_seqlen, int batch, int seqlen,
                        int n_heads_q, int n_heads_k, int head_dim, int64_t q_row_stride,
                        int64_t k_row_stride, int64_t v_row_stride, int64_t do_row_stride,
                        float softmax_scale, int window_left, float softcap, int dtype,
                        cudaStream_t stream);

/* ---- small helpers of the LoRA path ---------------------------------------------------------
 * Writes the whole [dst_rows, dst_cols] destination: the block at (dst_row_off, dst_col_off)
 * receives scale * src (src is [rows, cols]; transposed first when transpose != 0, i.e. the
 * block is then [cols, rows]) and everything else is zero.  This is A.to(dtype) /
 * (s*B).to(dtype) of matmul_lora (unsloth/kernels/utils.py:1163-1167), zero-padded to the
 * 64-wide K block of the GEMM and placed at the adapter's slot of a shared rank block.        */
int ub200_cast_pad_2d(const void* src, int src_dtype, int64_t src_ld, int rows, int cols,
                      void* dst, int dst_dtype, int64_t dst_ld, int dst_rows, int dst_cols,
                      int dst_row_off, int dst_col_off, float scale, int transpose,
                      cudaStream_t stream);

/* Batched forms of the two per-adapter housekeeping steps of a LoRA training step (448 tensors each on
 * Llama-3-8B): descriptors are a HOST array, forwarded by value in the kernel parameters (chunks of 40 per
 * launch), so the calls are CUDA-graph capturable and need no device staging.
 *  ub200_cast_pad_multi : every element = ub200_cast_pad_2d of one descriptor (the per-step A.to(dtype) /
 *                         (s*B).to(dtype) refresh of unsloth/kernels/utils.py:1163-1167, fast_lora.py:138-153)
 *  ub200_accumulate_multi: dst[r, c] (contiguous fp32 [rows, cols]) += src[r*src_rs + c*src_cs] -- what
 *                         autograd's AccumulateGrad does for the d_A / d_B views the reference's backward
 *                         returns (fast_lora.py:206-229, 519-540, 639-650), as one launch per 40 adapters.      */
typedef struct {
  const void* src; void* dst;
  int64_t src_ld, dst_ld;
  int src_dtype, dst_dtype;
  int rows, cols, dst_rows, dst_cols, row_off, col_off;
  float scale;
  int transpose;
} ub200_cast_desc;
typedef struct {
  const float* src; float* dst;
  int64_t src_rs, src_cs;
  int rows, cols;
} ub200_acc_desc;
int ub200_cast_pad_multi(const ub200_cast_desc* descs, int n, cudaStream_t stream);
int ub200_accumulate_multi(const ub200_acc_desc* descs, int n, cudaStream_t stream);

/* Flat AdamW over the LoRA parameter bucket (fp32 p, g, m, v of length n), decoupled weight
 * decay; bias corrections are passed in (1 - beta^t).  grad_scale multiplies g first.          */
int ub200_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, float bias_corr1,
                     float bias_corr2, float grad_scale, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNSLOTH_B200_H_ */
