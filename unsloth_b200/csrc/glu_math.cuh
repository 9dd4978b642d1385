// Gated-activation math shared by the standalone GLU kernels (glu.cu) and the GLU epilogue of the
// tcgen05 GEMM (gemm_tcgen05.cu): ONE definition, so the fused and the two-kernel paths round at
// the same points and produce the same bits.
//   unsloth/kernels/swiglu.py:27-47, 67-109; unsloth/kernels/geglu.py:31-53, 74-123, 142-167, 188-244
#pragma once
#include "common.cuh"

namespace ub {

enum { ACT_SWIGLU = 0, ACT_GEGLU_APPROX = 1, ACT_GEGLU_EXACT = 2 };

// returns f(e) (fp32, before rounding) and df/de.  FAST (16-bit tensors): the sigmoid uses the
// MUFU exp2 / reciprocal path (relative error ~1e-6, far below the bf16/fp16 output rounding);
// fp32 tensors keep the accurate expf and IEEE division so the 1e-5 gate holds.
template <int ACT, bool FAST>
__device__ __forceinline__ void act_eval(float e, float& f, float& dfde) {
  if (ACT == ACT_SWIGLU) {
    const float se = FAST ? __frcp_rn(1.0f + __expf(-e)) : 1.0f / (1.0f + expf(-e));
    f = e * se;
    dfde = se * (1.0f + e * (1.0f - se));
  } else if (ACT == ACT_GEGLU_APPROX) {
    const float s = 0.7978845608028654f;
    const float a = s * e;
    const float b = a * 0.044715f * e * e;
    const float T = 1.0f + tanhf(a + b);
    const float T2 = 0.5f * T;
    const float Q2 = -T2 * (T - 2.0f) * (a + 3.0f * b);
    f = T2 * e;
    dfde = T2 + Q2;
  } else {
    const float fp = 0.5f * (erff(0.70710678118654752f * e) + 1.0f);
    f = fp * e;
    dfde = fp + 0.3989422804014327f * e * expf(-0.5f * e * e);
  }
}

// One element of the in-place backward (glu.cu's glu_bwd_kernel body): h = f*g, df = DW*f, de = (DW*g)*f'(e),
// with f and DW*g rounded to the tensor dtype first (swiglu.py:86-109).
template <typename T, int ACT>
__device__ __forceinline__ void glu_bwd_elem(float dwk, float ek, float gk, float& h, float& df, float& de) {
  float f, d;
  act_eval<ACT, sizeof(T) == 2>(ek, f, d);
  const float fr = DT<T>::rnd(f);
  h = fr * gk;
  df = dwk * fr;
  const float dg = DT<T>::rnd(dwk * gk);
  de = dg * d;
}
// One element of the forward: h = round(f(e)) * g (swiglu.py:37-47).
template <typename T, int ACT>
__device__ __forceinline__ float glu_fwd_elem(float ek, float gk) {
  float f, d;
  act_eval<ACT, sizeof(T) == 2>(ek, f, d);
  return DT<T>::rnd(f) * gk;
}

}  // namespace ub
