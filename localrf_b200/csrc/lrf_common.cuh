// lrf_common.cuh -- device-side structures and small helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lrf {

constexpr int CD = 8;        // density components per plane   (opt.py:117 n_lamb_sigma)
constexpr int CA = 24;       // appearance components per plane (opt.py:118 n_lamb_sh)
constexpr int NF = 3 * CA;   // 72 appearance features per sample
constexpr int APP_DIM = 27;  // opt.py:119 data_dim_color
constexpr int FC = 128;      // opt.py:155 featureC

// matMode / vecMode of the reference (models/tensorBase.py:274-275)
__host__ __device__ constexpr int mat0(int i) { return i == 2 ? 1 : 0; }
__host__ __device__ constexpr int mat1(int i) { return i == 0 ? 1 : 2; }
__host__ __device__ constexpr int vecm(int i) { return 2 - i; }

// Layout of the per-field prepared block (floats).  All matrices are k-major ("transposed") so a
// thread block reads rows of consecutive outputs.
//   W1B [NF][FC]  : (mlp[0].weight @ basis_mat.weight)^T  -- basis (72->27) folded into layer 1
//   W2T [FC][FC]  : mlp[2].weight^T
//   B1 [FC], B2 [FC], W3 [3][FC+3] (row-major, padded to 3*132), B3 [4]
constexpr int PREP_W1B = 0;
constexpr int PREP_W2T = PREP_W1B + NF * FC;
constexpr int PREP_B1 = PREP_W2T + FC * FC;
constexpr int PREP_B2 = PREP_B1 + FC;
constexpr int PREP_W3 = PREP_B2 + FC;         // 3 rows of 132 floats (131 used)
constexpr int W3_LD = 132;
constexpr int PREP_B3 = PREP_W3 + 3 * W3_LD;
constexpr int PREP_FLOATS = PREP_B3 + 4;      // multiple of 4 -> 16-byte sized for bulk copies
static_assert(PREP_FLOATS % 4 == 0, "prepared block must be a multiple of 16 bytes");

struct FieldDev {
  int g[3];
  float amin[3], ainv[3];     // aabb min, 2/(max-min)   (tensorBase.py:321,342-345)
  const float* dplane[3];
  const float* dline[3];
  const float* aplane[3];
  const float* aline[3];
  const float* prep;
  const float* alpha_vol;     // may be null
  int ad[3];                  // alpha dims D,H,W
  float aamin[3], aainv[3];   // alpha aabb min, invgridSize (tensorBase.py:45)
  float density_shift, distance_scale, weight_thres;
  int act;
  const float* z;
  int S;
};

struct BatchDev {
  long long n_rays;
  const float* rays;
  const long long* ray_ids;
  int W, H, fov360;
  float focal, cx, cy;
  const float* intrinsics;
  const float* c2w;
  long long rays_per_view;
  const float* w2rf;
  const float* blend;
  long long blend_stride;
  const float* exposure;
  int accumulate, finalize, white_bg;
  float floater_thresh;
  float* rgb;
  float* depth;
  float* weights;
  float* dirs;
  unsigned long long* stats;
};

// ---- sampling helpers -------------------------------------------------------------------------

// ATen grid_sampler coordinate transform, align_corners=True, padding_mode="border"
// (call sites models/tensoRF.py:135-146): pixel = clamp((c+1)/2*(size-1), 0, size-1)
__device__ __forceinline__ void grid_coord(float c, int size, int& i0, int& i1, float& t) {
  float x = ((c + 1.0f) * 0.5f) * (float)(size - 1);
  x = fminf((float)(size - 1), fmaxf(x, 0.0f));
  float f = floorf(x);
  i0 = (int)f;
  i1 = min(i0 + 1, size - 1);
  t = x - f;
}

// utils/ray_utils.py:9-12
__device__ __forceinline__ void contract(float& x, float& y, float& z) {
  float n = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
  n = fmaxf(n, 1e-6f);
  if (!(n <= 1.0f)) {
    float s = __fdiv_rn(2.0f * n - 1.0f, n * n);
    x *= s; y *= s; z *= s;
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// models/tensorBase.py:495-499
__device__ __forceinline__ float feature2density(float f, float shift, int act) {
  if (act == 0) {
    float x = f + shift;
    return x > 20.0f ? x : log1pf(expf(x));  // F.softplus(beta=1, threshold=20)
  }
  return fmaxf(f, 0.0f);
}

}  // namespace lrf
