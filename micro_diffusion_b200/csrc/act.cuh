// Activation storage type of the element-wise kernels.  The product path keeps GEMM operands / saved activations in
// bf16 (prec = 0, the reference's amp_bf16 regime, train.py:113); prec = 1 ("high precision", MD_PRECISION=high) keeps
// them in fp32 so that the SAME kernels and the same host sequencing can be gated against the fp32 oracle at the
// tolerance north_star states (1e-3 on loss and denoiser output).  All kernels work on groups of 8 (or 4) elements:
// one 16-byte load for bf16, two for fp32 -- loop structure and thread mapping are identical in both modes.
#pragma once
#include "common.cuh"

namespace md {

template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> { uint4 r; };
template <> struct V8<float> { float4 a, b; };

__device__ __forceinline__ V8<__nv_bfloat16> ldv8(const __nv_bfloat16* p) {
  V8<__nv_bfloat16> v;
  v.r = *reinterpret_cast<const uint4*>(p);
  return v;
}
__device__ __forceinline__ V8<float> ldv8(const float* p) {
  V8<float> v;
  v.a = *reinterpret_cast<const float4*>(p);
  v.b = *reinterpret_cast<const float4*>(p + 4);
  return v;
}
template <typename T> __device__ __forceinline__ V8<T> zerov8();
template <> __device__ __forceinline__ V8<__nv_bfloat16> zerov8<__nv_bfloat16>() {
  V8<__nv_bfloat16> v;
  v.r = make_uint4(0, 0, 0, 0);
  return v;
}
template <> __device__ __forceinline__ V8<float> zerov8<float>() {
  V8<float> v;
  v.a = v.b = make_float4(0.f, 0.f, 0.f, 0.f);
  return v;
}
__device__ __forceinline__ void unpackv8(const V8<__nv_bfloat16>& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v.r);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __low2float(h[e]);
    f[2 * e + 1] = __high2float(h[e]);
  }
}
__device__ __forceinline__ void unpackv8(const V8<float>& v, float (&f)[8]) {
  f[0] = v.a.x; f[1] = v.a.y; f[2] = v.a.z; f[3] = v.a.w;
  f[4] = v.b.x; f[5] = v.b.y; f[6] = v.b.z; f[7] = v.b.w;
}
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&f)[8]) { unpackv8(ldv8(p), f); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
// groups of 4
__device__ __forceinline__ float4 ld4a(const __nv_bfloat16* p) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ float4 ld4a(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4a(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}
__device__ __forceinline__ void st4a(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float2 ld2a(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ float2 ld2a(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float ld1a(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1a(const float* p) { return *p; }
__device__ __forceinline__ void st1a(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st1a(float* p, float v) { *p = v; }

}  // namespace md

// Runs `...` once with `AT` bound to the activation storage type selected by `prec` (0 bf16, 1 fp32).
#define MD_WITH_ACT(prec, ...)      \
  do {                              \
    if (prec) {                     \
      using AT = float;             \
      __VA_ARGS__;                  \
    } else {                        \
      using AT = __nv_bfloat16;     \
      __VA_ARGS__;                  \
    }                               \
  } while (0)
#define AP(T, p) reinterpret_cast<T*>(p)
#define CAP(T, p) reinterpret_cast<const T*>(p)
