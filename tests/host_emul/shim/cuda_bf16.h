// Host shim of the few bfloat16 pieces lrf_device.cuh uses (test infrastructure, see cuda_runtime.h here).
#pragma once
#include <cuda_runtime.h>   // the shim next to this file (float2, ...)
struct __nv_bfloat16 { uint16_t x; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline uint16_t shim_f2bf_rn(float f) {          // round to nearest even, like cvt.rn.bf16.f32
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float shim_bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
  __nv_bfloat162 r; r.x.x = shim_f2bf_rn(a); r.y.x = shim_f2bf_rn(b); return r;
}
static inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{shim_bf2f(h.x.x), shim_bf2f(h.y.x)}; }
