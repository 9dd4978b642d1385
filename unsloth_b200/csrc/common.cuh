// Shared device/host helpers for the unsloth_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/unsloth_b200.h"

#define UB_SM_COUNT 148

// Launch-error check: never synchronises (SURVEY 8b conventions).
#define UB_RETURN_LAST()                                                      \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    return e__ == cudaSuccess ? UB200_OK : (int)e__;                          \
  } while (0)

namespace ub {

// ---- dtype traits -----------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float> {
  static constexpr int VEC = 4;  // elements per 16 B
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  __device__ static __forceinline__ float rnd(float v) { return v; }  // round-trip
};
template <> struct DT<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct DT<__half> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};

// round a float to the precision of dtype code `dt` (UB200_F32/F16/BF16)
__device__ __forceinline__ float round_to(int dt, float v) {
  if (dt == UB200_BF16) return DT<__nv_bfloat16>::rnd(v);
  if (dt == UB200_F16) return DT<__half>::rnd(v);
  return v;
}

// ---- 16-byte vector I/O -------------------------------------------------------------
template <typename T> struct alignas(16) Vec16 { T v[16 / sizeof(T)]; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[DT<T>::VEC]) {
  Vec16<T> r = *reinterpret_cast<const Vec16<T>*>(p);
#pragma unroll
  for (int i = 0; i < DT<T>::VEC; ++i) out[i] = DT<T>::to_f(r.v[i]);
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&in)[DT<T>::VEC]) {
  Vec16<T> r;
#pragma unroll
  for (int i = 0; i < DT<T>::VEC; ++i) r.v[i] = DT<T>::from_f(in[i]);
  *reinterpret_cast<Vec16<T>*>(p) = r;
}
// streaming (evict-first) variants for data touched once
template <typename T>
__device__ __forceinline__ void load_vec_cs(const T* p, float (&out)[DT<T>::VEC]) {
  int4 q = __ldcs(reinterpret_cast<const int4*>(p));
  Vec16<T> r = *reinterpret_cast<Vec16<T>*>(&q);
#pragma unroll
  for (int i = 0; i < DT<T>::VEC; ++i) out[i] = DT<T>::to_f(r.v[i]);
}
template <typename T>
__device__ __forceinline__ void store_vec_cs(T* p, const float (&in)[DT<T>::VEC]) {
  Vec16<T> r;
#pragma unroll
  for (int i = 0; i < DT<T>::VEC; ++i) r.v[i] = DT<T>::from_f(in[i]);
  __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<int4*>(&r));
}

// generic-dtype scalar load -> float (used for the small weight / table operands)
__device__ __forceinline__ float load_as_f(const void* p, int dt, int64_t i) {
  if (dt == UB200_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  if (dt == UB200_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}

// V consecutive weights/table entries starting at element `i0` (multiple of V) as floats.  One
// 16-byte (or two, for fp32 tables feeding 16-bit activations) vector load when alignment allows.
template <int V>
__device__ __forceinline__ void load_vec_as_f(const void* p, int dt, int64_t i0, float (&out)[V]) {
  if (dt == UB200_F32) {
    const float* f = reinterpret_cast<const float*>(p) + i0;
    if ((reinterpret_cast<uintptr_t>(f) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < V / 4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(f)[q];
        out[4 * q] = t.x; out[4 * q + 1] = t.y; out[4 * q + 2] = t.z; out[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) out[i] = f[i];
    }
  } else {
    const uint16_t* h = reinterpret_cast<const uint16_t*>(p) + i0;
    uint16_t raw[V];
    if (V == 8 && (reinterpret_cast<uintptr_t>(h) & 15) == 0) {
      const uint4 t = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(raw) = t;
    } else if (V == 4 && (reinterpret_cast<uintptr_t>(h) & 7) == 0) {
      const uint2 t = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(raw) = t;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) raw[i] = h[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i)
      out[i] = dt == UB200_BF16 ? __bfloat162float(__ushort_as_bfloat16(raw[i]))
                                : __half2float(__ushort_as_half(raw[i]));
  }
}

// ---- reductions ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is a __shared__ float[32]; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

inline int dtype_size(int dt) { return dt == UB200_F32 ? 4 : 2; }

}  // namespace ub
