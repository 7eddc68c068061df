// Host shim (TEST INFRASTRUCTURE): lets g++ compile the device-side building blocks of
// localrf_b200/csrc/lrf_common.cuh and lrf_device.cuh as plain host functions, so the CPU test suite
// (no GPU in the build container) can run the kernels' per-sample arithmetic -- grid coordinates, contraction,
// VM gathers in fp32 and bf16 storage, alpha-mask lookup, ray generation -- against the oracle.
// Only what those two headers need; PTX helpers compile as never-expanded inline functions.
#pragma once
#define _GNU_SOURCE 1
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cstddef>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fdiv_rn(float a, float b) { return a / b; }       // IEEE division, round to nearest
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
using std::max;
using std::min;
// warp intrinsics: the shim runs one "lane" at a time; the warp helpers of lrf_device.cuh are not exercised
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline float __shfl_up_sync(unsigned, float v, int) { return v; }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
