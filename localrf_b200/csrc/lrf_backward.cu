// lrf_backward.cu -- backward of the per-ray-batch render (TensorBase.forward with
// floater_thresh = 0; tensorBase.py:567-636): dL/d(rgb_map), dL/d(depth_map) -> dL/d(rays) and
// dL/d(every parameter of the field), in three launches that share a scratch block in HBM:
//
//   bwd_march_kernel    warp = ray, lane = sample.  Marches the ray again (feature, alpha,
//                       transmittance by a warp-shuffle scan, weight), stores the per-sample tables,
//                       seeds dL/dw with the depth / background terms and appends the shaded samples
//                       (weight > rayMarch_weight_thres) to one global list.
//   bwd_shade_kernel    persistent CTAs over 64-sample tiles of that list.  Recomputes the 72 plane x
//                       line products and the MLP (folded layer 1, layer 2, layer 3 + sigmoid) and runs
//                       its backward as register-tiled fp32 GEMMs out of shared memory; every thread
//                       owns a fixed block of dW2 / d(W1 @ basis) in registers for the whole launch
//                       (flushed once), product gradients scatter to the appearance planes / lines,
//                       the position gradient goes through the contraction Jacobian to the ray, and
//                       g . rgb is added to the sample's dL/dw.
//   bwd_density_kernel  warp = ray, lane = sample.  dL/dw -> dL/dalpha by a reverse (suffix) scan,
//                       through alpha = 1 - exp(-sigma dist scale) and softplus / relu to the density
//                       feature, whose lookup backward scatters to the density planes / lines (16-byte
//                       vector reductions); finishes d(rays) (vd = d/|d|, depth = sum(w z)/|d|).
//
// The math is the one pinned by oracle/lrf_oracle.c::orc_field_backward against the reference's
// autograd.  The shade step has two implementations: bwd_shade_tc_kernel (lrf_backward_tc.cuh, the six
// matrix products on tcgen05 with bf16 hi/lo operands and TMEM-resident weight gradients; default) and
// bwd_shade_kernel below (fp32 on the CUDA cores, LRF_BWD_TC=0).
#include "../../include/localrf_b200.h"
#include <cstdlib>

#include "lrf_device.cuh"

namespace lrf {

// prepared block of the backward (floats): folded layer 1 and W2, both transposed (k-major)
constexpr int BP_W1BT = 0;                      // [NF][FC]   (W1 @ basis)^T
constexpr int BP_W2T = BP_W1BT + NF * FC;       // [FC(k)][FC(n)]
constexpr int BP_B1 = BP_W2T + FC * FC;
constexpr int BP_B2 = BP_B1 + FC;
constexpr int BP_W3 = BP_B2 + FC;               // [3][W3_LD]
constexpr int BP_B3 = BP_W3 + 3 * W3_LD;
constexpr int BP_FLOATS = BP_B3 + 4;

struct BwdScratch {           // views into the caller's scratch block
  unsigned long long* count;  // number of shaded samples (zeroed per call)
  float* d_ovd;               // [n][6] position-gradient accumulators (zeroed per call)
  float *alpha, *T, *w, *fe, *gw;   // [n][S]
  unsigned char* valid;       // [n][S]
  int *app_ray, *app_k;       // [n*S]
};

struct BwdArgs {
  long long n_rays;
  const float* rays;          // [n][6]
  const float* g_rgb;         // [n][3]
  const float* g_depth;       // [n]
  const float* bp;            // prepared block above
  int white_bg;
  float* d_rays;              // [n][6]
  float* d_dplane[3]; float* d_dline[3]; float* d_aplane[3]; float* d_aline[3];
  float* d_w1b; float* d_b1; float* d_w2; float* d_b2; float* d_w3; float* d_b3;
  BwdScratch s;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ inline size_t bwd_scratch_layout(long long n, int S, BwdScratch* out, unsigned char* base,
                                          size_t* zero_bytes) {
  const size_t ns = (size_t)n * (size_t)S;
  size_t off = 0;
  const size_t o_count = off; off += 16;
  const size_t o_ovd = off;   off = align16(off + (size_t)n * 6 * sizeof(float));
  if (zero_bytes) *zero_bytes = off;
  const size_t o_alpha = off; off = align16(off + ns * 4);
  const size_t o_T = off;     off = align16(off + ns * 4);
  const size_t o_w = off;     off = align16(off + ns * 4);
  const size_t o_fe = off;    off = align16(off + ns * 4);
  const size_t o_gw = off;    off = align16(off + ns * 4);
  const size_t o_ray = off;   off = align16(off + ns * 4);
  const size_t o_k = off;     off = align16(off + ns * 4);
  const size_t o_valid = off; off = align16(off + ns);
  if (out && base) {
    out->count = reinterpret_cast<unsigned long long*>(base + o_count);
    out->d_ovd = reinterpret_cast<float*>(base + o_ovd);
    out->alpha = reinterpret_cast<float*>(base + o_alpha);
    out->T = reinterpret_cast<float*>(base + o_T);
    out->w = reinterpret_cast<float*>(base + o_w);
    out->fe = reinterpret_cast<float*>(base + o_fe);
    out->gw = reinterpret_cast<float*>(base + o_gw);
    out->app_ray = reinterpret_cast<int*>(base + o_ray);
    out->app_k = reinterpret_cast<int*>(base + o_k);
    out->valid = base + o_valid;
  }
  return off;
}

// ---- small device helpers --------------------------------------------------------------------------

// grid_sample border coordinate with derivative (zero when clipped)
__device__ __forceinline__ void coord_g(float c, int size, int& i0, int& i1, float& t, float& dc) {
  float x = ((c + 1.0f) * 0.5f) * (float)(size - 1);
  dc = (x <= 0.0f || x >= (float)(size - 1)) ? 0.0f : 0.5f * (float)(size - 1);
  x = fminf((float)(size - 1), fmaxf(x, 0.0f));
  const float f = floorf(x);
  i0 = (int)f;
  i1 = min(i0 + 1, size - 1);
  t = x - f;
}

// d contract(p)/dp applied to dpc (ray_utils.py:9-12; amax routes the norm's gradient to the
// component of largest magnitude)
__device__ __forceinline__ void contract_bwd(float p0, float p1, float p2, float g0, float g1,
                                             float g2, float& o0, float& o1, float& o2) {
  float n = fmaxf(fmaxf(fabsf(p0), fabsf(p1)), fabsf(p2));
  n = fmaxf(n, 1e-6f);
  if (n <= 1.0f) { o0 = g0; o1 = g1; o2 = g2; return; }
  const float s = __fdiv_rn(2.0f * n - 1.0f, n * n);
  const float ds = -2.0f / (n * n) + 2.0f / (n * n * n);
  const float dot = g0 * p0 + g1 * p1 + g2 * p2;
  int j = 0;
  float best = fabsf(p0);
  if (fabsf(p1) > best) { j = 1; best = fabsf(p1); }
  if (fabsf(p2) > best) { j = 2; }
  o0 = s * g0; o1 = s * g1; o2 = s * g2;
  const float pj = j == 0 ? p0 : (j == 1 ? p1 : p2);
  const float extra = (pj >= 0.0f ? 1.0f : -1.0f) * ds * dot;
  if (j == 0) o0 += extra; else if (j == 1) o1 += extra; else o2 += extra;
}

__device__ __forceinline__ void red4g(float* p, float x, float y, float z, float w) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(x, y, z, w));
}

__device__ __forceinline__ void load_ray(const float* rays, long long ray, RaySm& R) {
  const float* rp = rays + 6 * ray;
  const float d0 = rp[3], d1 = rp[4], d2 = rp[5];
  R.o[0] = rp[0]; R.o[1] = rp[1]; R.o[2] = rp[2];
  R.nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  R.vd[0] = __fdiv_rn(d0, R.nrm); R.vd[1] = __fdiv_rn(d1, R.nrm); R.vd[2] = __fdiv_rn(d2, R.nrm);
  R.blend = 1.0f;
}

// ====================================================================================================
// 1. march
// ====================================================================================================
constexpr int MARCH_THREADS = 256;

__global__ void __launch_bounds__(MARCH_THREADS)
bwd_march_kernel(const FieldDev F, const BwdArgs A) {
  const int S = F.S;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * MARCH_THREADS + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * MARCH_THREADS) >> 5;
  for (long long ray = warp0; ray < A.n_rays; ray += n_warps) {
    RaySm R;
    load_ray(A.rays, ray, R);
    const float gsum = A.white_bg ? -(A.g_rgb[3 * ray] + A.g_rgb[3 * ray + 1] + A.g_rgb[3 * ray + 2]) : 0.0f;
    const float gd = A.g_depth[ray];
    const size_t base = (size_t)ray * S;
    float carry = 1.0f;
    for (int k0 = 0; k0 < S; k0 += 32) {
      const int k = k0 + lane;
      float alpha = 0.0f, fe = 0.0f, z = 0.0f;
      bool valid = false;
      if (k < S) {
        z = __ldg(F.z + k);
        float p[3], q[3];
        sample_pos(F, R, z, p, q);
        valid = (k != S - 1);                                          // ray_valid[:, -1] = 0
        if (valid && F.alpha_vol) valid = alpha_mask(F, p) > 0.0f;     // tensorBase.py:593-598
        float sigma = 0.0f;
        if (valid) { fe = density_feature(F, q); sigma = feature2density(fe, F.density_shift, F.act); }
        const float znext = (k + 1 < S) ? __ldg(F.z + k + 1) : z;
        alpha = 1.0f - expf(-sigma * (znext - z) * F.distance_scale);   // tensorBase.py:610
        if (k == S - 1) alpha = 1.0f;                                  // alpha[:, -1] = 1
      }
      const float f = (k < S) ? (1.0f - alpha) + 1e-10f : 1.0f;
      const float inc = warp_scan_mul(f, lane);
      float exc = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) exc = 1.0f;
      const float Tk = carry * exc;
      const float wk = alpha * Tk;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
      const bool on = (k < S) && (wk > F.weight_thres);                // tensorBase.py:622
      if (k < S) {
        A.s.alpha[base + k] = alpha; A.s.T[base + k] = Tk; A.s.w[base + k] = wk;
        A.s.fe[base + k] = fe; A.s.valid[base + k] = valid ? 1 : 0;
        A.s.gw[base + k] = gsum + gd * __fdiv_rn(z, R.nrm);            // d rgb_map(white bg) + d depth
      }
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (m) {
        unsigned long long slot = 0;
        if (lane == 0) slot = atomicAdd(A.s.count, (unsigned long long)__popc(m));
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (on) {
          const unsigned long long e = slot + __popc(m & ((1u << lane) - 1u));
          A.s.app_ray[e] = (int)ray;
          A.s.app_k[e] = k;
        }
      }
    }
  }
}

// ====================================================================================================
// 2. shade: MLP forward + backward over tiles of 64 shaded samples
// ====================================================================================================
constexpr int SH_THREADS = 256;
constexpr int TS = 64;                 // samples per tile
constexpr int LDW = FC + 1;            // padded row of the transposed weights / activations
constexpr int LDX = NF + 1;            // padded row of the product tile

struct ShadeSmem {                     // float offsets
  static constexpr int w1bt = 0;                        // [NF][LDW]
  static constexpr int w2t = w1bt + NF * LDW;           // [FC][LDW]
  static constexpr int x = w2t + FC * LDW;              // [TS][LDX]  products, then their gradients
  static constexpr int h1 = x + TS * LDX;               // [TS][LDW]  h1, then dL/d(pre-activation 1)
  static constexpr int h2 = h1 + TS * LDW;              // [TS][LDW]  h2, then dL/d(pre-activation 2)
  static constexpr int w3 = h2 + TS * LDW;              // [3][W3_LD]
  static constexpr int b1 = w3 + 3 * W3_LD;
  static constexpr int b2 = b1 + FC;
  static constexpr int q = b2 + FC;                     // [TS][3] normalised grid coordinates
  static constexpr int praw = q + TS * 3;               // [TS][3] un-contracted sample position
  static constexpr int vd = praw + TS * 3;              // [TS][3]
  static constexpr int g = vd + TS * 3;                 // [TS][3] dL/d rgb_map of the sample's ray
  static constexpr int dpre = g + TS * 3;               // [TS][3] dL/d(layer-3 pre-sigmoid)
  static constexpr int wk = dpre + TS * 3;              // [TS]
  static constexpr int zk = wk + TS;                    // [TS]
  static constexpr int ray = zk + TS;                   // [TS] int
  static constexpr int kk = ray + TS;                   // [TS] int
  static constexpr int total = kk + TS;
};

__global__ void __launch_bounds__(SH_THREADS, 1)
bwd_shade_kernel(const FieldDev F, const BwdArgs A) {
  extern __shared__ __align__(16) float sm[];
  using L = ShadeSmem;
  float* W1BT_s = sm + L::w1bt;
  float* W2T_s = sm + L::w2t;
  float* X_s = sm + L::x;
  float* H1_s = sm + L::h1;
  float* H2_s = sm + L::h2;
  float* W3_s = sm + L::w3;
  float* b1_s = sm + L::b1;
  float* b2_s = sm + L::b2;
  float* q_s = sm + L::q;
  float* praw_s = sm + L::praw;
  float* vd_s = sm + L::vd;
  float* g_s = sm + L::g;
  float* dpre_s = sm + L::dpre;
  float* wk_s = sm + L::wk;
  float* zk_s = sm + L::zk;
  int* ray_s = reinterpret_cast<int*>(sm + L::ray);
  int* kk_s = reinterpret_cast<int*>(sm + L::kk);

  const int tid = threadIdx.x;
  const int tn = tid & 15, ts = tid >> 4;       // GEMM tiles: 4 samples x 8 units, or 8 x 8 weights
  const int S = F.S;

  // weights -> shared memory (padded rows: conflict-free both along n and along k)
  for (int e = tid; e < NF * FC; e += SH_THREADS) W1BT_s[(e / FC) * LDW + (e % FC)] = __ldg(A.bp + BP_W1BT + e);
  for (int e = tid; e < FC * FC; e += SH_THREADS) W2T_s[(e / FC) * LDW + (e % FC)] = __ldg(A.bp + BP_W2T + e);
  for (int e = tid; e < 3 * W3_LD; e += SH_THREADS) W3_s[e] = __ldg(A.bp + BP_W3 + e);
  for (int e = tid; e < FC; e += SH_THREADS) { b1_s[e] = __ldg(A.bp + BP_B1 + e); b2_s[e] = __ldg(A.bp + BP_B2 + e); }
  const float b3r[3] = {__ldg(A.bp + BP_B3), __ldg(A.bp + BP_B3 + 1), __ldg(A.bp + BP_B3 + 2)};

  // per-thread gradient accumulators, kept for the whole launch
  float accW2[8][8];      // dW2[n = 8 ts + i][k = tn + 16 j]
  float accW1[8][5];      // dW1B[n = 8 ts + i][t = tn + 16 j], t < 72
  float accW3[3] = {0.f, 0.f, 0.f};   // tid < 128: dW3[c][n = tid]; 128..130: dW3[c][128 + a]; 131: db3
  float accB1 = 0.0f, accB2 = 0.0f;   // tid < 128
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) accW2[i][j] = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j) accW1[i][j] = 0.0f;
  }
  __syncthreads();

  const long long n_app = (long long)*A.s.count;
  const long long n_tiles = (n_app + TS - 1) / TS;
  const int ls = tid >> 2, cq = tid & 3;        // gather / scatter: sample ls, channels cq*6 .. cq*6+5

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long e0 = tile * TS;
    const int n_valid = (int)((n_app - e0) < TS ? (n_app - e0) : TS);
    // ---- 2a. tile header: ray data of each sample ---------------------------------------------------
    if (tid < TS) {
      int ray = 0, k = 0;
      float wk = 0.0f;
      if (tid < n_valid) { ray = A.s.app_ray[e0 + tid]; k = A.s.app_k[e0 + tid]; wk = A.s.w[(size_t)ray * S + k]; }
      ray_s[tid] = ray; kk_s[tid] = k; wk_s[tid] = wk;
      RaySm R;
      load_ray(A.rays, ray, R);
      const float z = __ldg(F.z + k);
      zk_s[tid] = z;
      float p[3], q[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) praw_s[tid * 3 + a] = R.o[a] + R.vd[a] * z;
      sample_pos(F, R, z, p, q);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        q_s[tid * 3 + a] = q[a];
        vd_s[tid * 3 + a] = R.vd[a];
        g_s[tid * 3 + a] = tid < n_valid ? A.g_rgb[3 * (size_t)ray + a] : 0.0f;
      }
    }
    __syncthreads();
    // ---- 2b. the 72 products (compute_appfeature before basis_mat, tensoRF.py:153-195) --------------
    {
      const float qq[3] = {q_s[ls * 3], q_s[ls * 3 + 1], q_s[ls * 3 + 2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int W = F.g[mat0(i)], H = F.g[mat1(i)], Ln = F.g[vecm(i)];
        int x0, x1, y0, y1, l0, l1; float tx, ty, tl;
        grid_coord(qq[mat0(i)], W, x0, x1, tx);
        grid_coord(qq[mat1(i)], H, y0, y1, ty);
        grid_coord(qq[vecm(i)], Ln, l0, l1, tl);
        const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
        const float w10 = (1.0f - tx) * ty, w11 = tx * ty, u0 = 1.0f - tl;
        for (int grp = cq; grp < CA / 4; grp += 4) {                    // 4 components per 16-byte group
          const float* P = F.aplane[i] + grp * 4;
          const float* Lp = F.aline[i] + grp * 4;
          const float4 a = ldg4(P + ((size_t)y0 * W + x0) * CA), b = ldg4(P + ((size_t)y0 * W + x1) * CA);
          const float4 c = ldg4(P + ((size_t)y1 * W + x0) * CA), d = ldg4(P + ((size_t)y1 * W + x1) * CA);
          const float4 u = ldg4(Lp + (size_t)l0 * CA), v = ldg4(Lp + (size_t)l1 * CA);
          float* xo = X_s + ls * LDX + i * CA + grp * 4;
          const bool ok = ls < n_valid;
          xo[0] = ok ? (a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11) * (u.x * u0 + v.x * tl) : 0.0f;
          xo[1] = ok ? (a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11) * (u.y * u0 + v.y * tl) : 0.0f;
          xo[2] = ok ? (a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11) * (u.z * u0 + v.z * tl) : 0.0f;
          xo[3] = ok ? (a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11) * (u.w * u0 + v.w * tl) : 0.0f;
        }
      }
    }
    __syncthreads();
    // ---- 2c. layer 1 (basis folded in) and layer 2, forward ----------------------------------------
    {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = b1_s[tn + 16 * j];
      for (int k = 0; k < NF; ++k) {
        float xv[4], wv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = X_s[(4 * ts + i) * LDX + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = W1BT_s[k * LDW + tn + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) H1_s[(4 * ts + i) * LDW + tn + 16 * j] = fmaxf(acc[i][j], 0.0f);
    }
    __syncthreads();
    {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = b2_s[tn + 16 * j];
      for (int k = 0; k < FC; ++k) {
        float xv[4], wv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = H1_s[(4 * ts + i) * LDW + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = W2T_s[k * LDW + tn + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) H2_s[(4 * ts + i) * LDW + tn + 16 * j] = fmaxf(acc[i][j], 0.0f);
    }
    __syncthreads();
    // ---- 2d. layer 3 + sigmoid; dL/d(pre-sigmoid); g . rgb joins the sample's dL/dw ------------------
    {
      float s3[3] = {0.f, 0.f, 0.f};
      for (int m = 0; m < FC / 4; ++m) {
        const int n = cq + 4 * m;
        const float h = H2_s[ls * LDW + n];
#pragma unroll
        for (int c = 0; c < 3; ++c) s3[c] = fmaf(W3_s[c * W3_LD + n], h, s3[c]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s3[c] += __shfl_xor_sync(0xffffffffu, s3[c], 1);
        s3[c] += __shfl_xor_sync(0xffffffffu, s3[c], 2);
      }
      if (cq == 0) {
        float grgb = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float s = s3[c] + W3_s[c * W3_LD + FC] * vd_s[ls * 3] + W3_s[c * W3_LD + FC + 1] * vd_s[ls * 3 + 1] +
                          W3_s[c * W3_LD + FC + 2] * vd_s[ls * 3 + 2] + b3r[c];
          const float rgb = __fdiv_rn(1.0f, 1.0f + expf(-s));
          const float gc = g_s[ls * 3 + c];
          dpre_s[ls * 3 + c] = ls < n_valid ? gc * wk_s[ls] * rgb * (1.0f - rgb) : 0.0f;
          grgb += gc * rgb;
        }
        if (ls < n_valid) A.s.gw[(size_t)ray_s[ls] * S + kk_s[ls]] += grgb;
      }
    }
    __syncthreads();
    // ---- 2e. layer 3 backward --------------------------------------------------------------------------
    if (tid < FC) {
      for (int s = 0; s < TS; ++s) {
        const float h = H2_s[s * LDW + tid];
#pragma unroll
        for (int c = 0; c < 3; ++c) accW3[c] = fmaf(dpre_s[s * 3 + c], h, accW3[c]);
      }
    } else if (tid < FC + 3) {                   // view-direction columns (viewdirs detached, :628)
      const int a = tid - FC;
      for (int s = 0; s < TS; ++s) {
        const float v = vd_s[s * 3 + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) accW3[c] = fmaf(dpre_s[s * 3 + c], v, accW3[c]);
      }
    } else if (tid == FC + 3) {                  // bias
      for (int s = 0; s < TS; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) accW3[c] += dpre_s[s * 3 + c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = 4 * ts + i;
      const float d0 = dpre_s[s * 3], d1 = dpre_s[s * 3 + 1], d2 = dpre_s[s * 3 + 2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = tn + 16 * j;
        const float gsum = W3_s[n] * d0 + W3_s[W3_LD + n] * d1 + W3_s[2 * W3_LD + n] * d2;
        const float h = H2_s[s * LDW + n];
        H2_s[s * LDW + n] = h > 0.0f ? gsum : 0.0f;                     // dL/d(pre-activation 2)
      }
    }
    __syncthreads();
    // ---- 2f. layer 2 backward: db2, dW2 += dh2^T h1, dh1 = dh2 W2 ------------------------------------
    if (tid < FC) {
      float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll 4
      for (int s = 0; s < TS; s += 4) {
        p0 += H2_s[s * LDW + tid]; p1 += H2_s[(s + 1) * LDW + tid];
        p2 += H2_s[(s + 2) * LDW + tid]; p3 += H2_s[(s + 3) * LDW + tid];
      }
      accB2 += (p0 + p1) + (p2 + p3);
    }
    for (int s = 0; s < TS; ++s) {
      float dv[8], hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = H2_s[s * LDW + 8 * ts + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) hv[j] = H1_s[s * LDW + tn + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) accW2[i][j] = fmaf(dv[i], hv[j], accW2[i][j]);
    }
    {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      for (int n = 0; n < FC; ++n) {
        float dv[4], wv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = H2_s[(4 * ts + i) * LDW + n];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = W2T_s[(tn + 16 * j) * LDW + n];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(dv[i], wv[j], acc[i][j]);
      }
      __syncthreads();                                                   // every reader of h1 is done
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int o = (4 * ts + i) * LDW + tn + 16 * j;
          H1_s[o] = H1_s[o] > 0.0f ? acc[i][j] : 0.0f;                   // dL/d(pre-activation 1)
        }
    }
    __syncthreads();
    // ---- 2g. layer 1 backward: db1, dW1B += dh1^T x, dprod = dh1 W1B -----------------------------------
    if (tid < FC) {
      float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll 4
      for (int s = 0; s < TS; s += 4) {
        p0 += H1_s[s * LDW + tid]; p1 += H1_s[(s + 1) * LDW + tid];
        p2 += H1_s[(s + 2) * LDW + tid]; p3 += H1_s[(s + 3) * LDW + tid];
      }
      accB1 += (p0 + p1) + (p2 + p3);
    }
    for (int s = 0; s < TS; ++s) {
      float dv[8], xv[5];
#pragma unroll
      for (int i = 0; i < 8; ++i) dv[i] = H1_s[s * LDW + 8 * ts + i];
#pragma unroll
      for (int j = 0; j < 5; ++j) xv[j] = (tn + 16 * j < NF) ? X_s[s * LDX + tn + 16 * j] : 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) accW1[i][j] = fmaf(dv[i], xv[j], accW1[i][j]);
    }
    {
      float acc[4][5];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.0f;
      for (int n = 0; n < FC; ++n) {
        float dv[4], wv[5];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = H1_s[(4 * ts + i) * LDW + n];
#pragma unroll
        for (int j = 0; j < 5; ++j) wv[j] = (tn + 16 * j < NF) ? W1BT_s[(tn + 16 * j) * LDW + n] : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(dv[i], wv[j], acc[i][j]);
      }
      __syncthreads();                                                   // every reader of x is done
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (tn + 16 * j < NF) X_s[(4 * ts + i) * LDX + tn + 16 * j] = acc[i][j];
    }
    __syncthreads();
    // ---- 2h. products backward: appearance planes / lines (16-byte reductions), position -> ray -----
    {
      float dq0 = 0.0f, dq1 = 0.0f, dq2 = 0.0f;
      if (ls < n_valid) {
        const float qq[3] = {q_s[ls * 3], q_s[ls * 3 + 1], q_s[ls * 3 + 2]};
        float dq[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int W = F.g[mat0(i)], H = F.g[mat1(i)], Ln = F.g[vecm(i)];
          int x0, x1, y0, y1, l0, l1; float tx, ty, tl, dx, dy, dl;
          coord_g(qq[mat0(i)], W, x0, x1, tx, dx);
          coord_g(qq[mat1(i)], H, y0, y1, ty, dy);
          coord_g(qq[vecm(i)], Ln, l0, l1, tl, dl);
          const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
          const float w10 = (1.0f - tx) * ty, w11 = tx * ty, u0 = 1.0f - tl;
          float gx = 0.0f, gy = 0.0f, gl = 0.0f;
          for (int grp = cq; grp < CA / 4; grp += 4) {                  // 4 components per 16-byte group
            const size_t o00 = ((size_t)y0 * W + x0) * CA + grp * 4, o01 = ((size_t)y0 * W + x1) * CA + grp * 4;
            const size_t o10 = ((size_t)y1 * W + x0) * CA + grp * 4, o11 = ((size_t)y1 * W + x1) * CA + grp * 4;
            const size_t ol0 = (size_t)l0 * CA + grp * 4, ol1 = (size_t)l1 * CA + grp * 4;
            const float4 a4 = ldg4(F.aplane[i] + o00), b4 = ldg4(F.aplane[i] + o01);
            const float4 c4 = ldg4(F.aplane[i] + o10), d4 = ldg4(F.aplane[i] + o11);
            const float4 u4 = ldg4(F.aline[i] + ol0), v4 = ldg4(F.aline[i] + ol1);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const float uv[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
            float dP[4], dL[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float gc = X_s[ls * LDX + i * CA + grp * 4 + e];
              dP[e] = gc * (uv[e] * u0 + vv[e] * tl);
              dL[e] = gc * (av[e] * w00 + bv[e] * w01 + cv[e] * w10 + dv[e] * w11);
              gx += dP[e] * ((bv[e] - av[e]) * (1.0f - ty) + (dv[e] - cv[e]) * ty);
              gy += dP[e] * ((cv[e] - av[e]) * (1.0f - tx) + (dv[e] - bv[e]) * tx);
              gl += dL[e] * (vv[e] - uv[e]);
            }
            red4g(A.d_aplane[i] + o00, dP[0] * w00, dP[1] * w00, dP[2] * w00, dP[3] * w00);
            red4g(A.d_aplane[i] + o01, dP[0] * w01, dP[1] * w01, dP[2] * w01, dP[3] * w01);
            red4g(A.d_aplane[i] + o10, dP[0] * w10, dP[1] * w10, dP[2] * w10, dP[3] * w10);
            red4g(A.d_aplane[i] + o11, dP[0] * w11, dP[1] * w11, dP[2] * w11, dP[3] * w11);
            red4g(A.d_aline[i] + ol0, dL[0] * u0, dL[1] * u0, dL[2] * u0, dL[3] * u0);
            red4g(A.d_aline[i] + ol1, dL[0] * tl, dL[1] * tl, dL[2] * tl, dL[3] * tl);
          }
          dq[mat0(i)] += gx * dx; dq[mat1(i)] += gy * dy; dq[vecm(i)] += gl * dl;
        }
        dq0 = dq[0]; dq1 = dq[1]; dq2 = dq[2];
      }
      dq0 += __shfl_xor_sync(0xffffffffu, dq0, 1); dq0 += __shfl_xor_sync(0xffffffffu, dq0, 2);
      dq1 += __shfl_xor_sync(0xffffffffu, dq1, 1); dq1 += __shfl_xor_sync(0xffffffffu, dq1, 2);
      dq2 += __shfl_xor_sync(0xffffffffu, dq2, 1); dq2 += __shfl_xor_sync(0xffffffffu, dq2, 2);
      if (cq == 0 && ls < n_valid) {
        float dp0, dp1, dp2;
        contract_bwd(praw_s[ls * 3], praw_s[ls * 3 + 1], praw_s[ls * 3 + 2], dq0 * F.ainv[0],
                     dq1 * F.ainv[1], dq2 * F.ainv[2], dp0, dp1, dp2);
        float* acc = A.s.d_ovd + 6 * (size_t)ray_s[ls];
        const float z = zk_s[ls];
        atomicAdd(acc + 0, dp0); atomicAdd(acc + 1, dp1); atomicAdd(acc + 2, dp2);
        atomicAdd(acc + 3, dp0 * z); atomicAdd(acc + 4, dp1 * z); atomicAdd(acc + 5, dp2 * z);
      }
    }
    __syncthreads();
  }

  // ---- flush the per-thread accumulators -----------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = 8 * ts + i;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (accW2[i][j] != 0.0f) atomicAdd(A.d_w2 + n * FC + tn + 16 * j, accW2[i][j]);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (tn + 16 * j < NF && accW1[i][j] != 0.0f) atomicAdd(A.d_w1b + n * NF + tn + 16 * j, accW1[i][j]);
  }
  if (tid < FC + 3) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (accW3[c] != 0.0f) atomicAdd(A.d_w3 + c * (FC + 3) + tid, accW3[c]);
  } else if (tid == FC + 3) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (accW3[c] != 0.0f) atomicAdd(A.d_b3 + c, accW3[c]);
  }
  if (tid < FC) {
    if (accB1 != 0.0f) atomicAdd(A.d_b1 + tid, accB1);
    if (accB2 != 0.0f) atomicAdd(A.d_b2 + tid, accB2);
  }
}

}  // namespace lrf
#include "lrf_backward_tc.cuh"   // tensor-core shade step (the default; LRF_BWD_TC=0 selects bwd_shade_kernel)
namespace lrf {

// ====================================================================================================
// 3. density branch + rays
// ====================================================================================================
__global__ void __launch_bounds__(MARCH_THREADS)
bwd_density_kernel(const FieldDev F, const BwdArgs A) {
  const int S = F.S;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * MARCH_THREADS + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * MARCH_THREADS) >> 5;
  const int n_chunks = (S + 31) / 32;
  for (long long ray = warp0; ray < A.n_rays; ray += n_warps) {
    RaySm R;
    load_ray(A.rays, ray, R);
    const float gd = A.g_depth[ray];
    const size_t base = (size_t)ray * S;
    float d_o[3] = {0.0f, 0.0f, 0.0f}, d_vd[3] = {0.0f, 0.0f, 0.0f};   // lane-partial sums
    float dep_p = 0.0f;
    float suffix_carry = 0.0f;                                           // sum_{j in later chunks} gw_j w_j
    for (int ch = n_chunks - 1; ch >= 0; --ch) {
      const int k = ch * 32 + lane;
      const bool in = k < S;
      const float gwk = in ? A.s.gw[base + k] : 0.0f;
      const float wk = in ? A.s.w[base + k] : 0.0f;
      const float zk = in ? __ldg(F.z + k) : 0.0f;
      dep_p += wk * zk;
      const float term = gwk * wk;
      float inc = term;                                                  // inclusive suffix sum
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += u;
      }
      const float excl = inc - term + suffix_carry;                      // sum_{j > k} gw_j w_j
      suffix_carry += __shfl_sync(0xffffffffu, inc, 0);
      if (k < S - 1 && A.s.valid[base + k]) {                            // last / masked samples: constants
        const float al = A.s.alpha[base + k];
        const float dalpha = gwk * A.s.T[base + k] - __fdiv_rn(excl, (1.0f - al) + 1e-10f);
        const float dsigma = dalpha * (1.0f - al) * (__ldg(F.z + k + 1) - zk) * F.distance_scale;
        const float fe = A.s.fe[base + k];
        float df;
        if (F.act == 0) {
          const float x = fe + F.density_shift;
          df = dsigma * (x > 20.0f ? 1.0f : __fdiv_rn(1.0f, 1.0f + expf(-x)));
        } else {
          df = fe > 0.0f ? dsigma : 0.0f;
        }
        if (df != 0.0f) {
          float p[3], q[3], dq[3] = {0.0f, 0.0f, 0.0f};
          sample_pos(F, R, zk, p, q);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int W = F.g[mat0(i)], H = F.g[mat1(i)], Ln = F.g[vecm(i)];
            int x0, x1, y0, y1, l0, l1; float tx, ty, tl, dx, dy, dl;
            coord_g(q[mat0(i)], W, x0, x1, tx, dx);
            coord_g(q[mat1(i)], H, y0, y1, ty, dy);
            coord_g(q[vecm(i)], Ln, l0, l1, tl, dl);
            const size_t o00 = ((size_t)y0 * W + x0) * CD, o01 = ((size_t)y0 * W + x1) * CD;
            const size_t o10 = ((size_t)y1 * W + x0) * CD, o11 = ((size_t)y1 * W + x1) * CD;
            const size_t ol0 = (size_t)l0 * CD, ol1 = (size_t)l1 * CD;
            const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
            const float w10 = (1.0f - tx) * ty, w11 = tx * ty, u0 = 1.0f - tl;
            float gx = 0.0f, gy = 0.0f, gl = 0.0f;
#pragma unroll
            for (int c = 0; c < CD; c += 4) {
              const float4 a = ldg4(F.dplane[i] + o00 + c), b = ldg4(F.dplane[i] + o01 + c);
              const float4 cc = ldg4(F.dplane[i] + o10 + c), d = ldg4(F.dplane[i] + o11 + c);
              const float4 u = ldg4(F.dline[i] + ol0 + c), v = ldg4(F.dline[i] + ol1 + c);
              const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
              const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, dv[4] = {d.x, d.y, d.z, d.w};
              const float uv[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
              float dP[4], dL[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float P = av[e] * w00 + bv[e] * w01 + cv[e] * w10 + dv[e] * w11;
                const float Lv = uv[e] * u0 + vv[e] * tl;
                dP[e] = df * Lv;
                dL[e] = df * P;
                gx += dP[e] * ((bv[e] - av[e]) * (1.0f - ty) + (dv[e] - cv[e]) * ty);
                gy += dP[e] * ((cv[e] - av[e]) * (1.0f - tx) + (dv[e] - bv[e]) * tx);
                gl += dL[e] * (vv[e] - uv[e]);
              }
              red4g(A.d_dplane[i] + o00 + c, dP[0] * w00, dP[1] * w00, dP[2] * w00, dP[3] * w00);
              red4g(A.d_dplane[i] + o01 + c, dP[0] * w01, dP[1] * w01, dP[2] * w01, dP[3] * w01);
              red4g(A.d_dplane[i] + o10 + c, dP[0] * w10, dP[1] * w10, dP[2] * w10, dP[3] * w10);
              red4g(A.d_dplane[i] + o11 + c, dP[0] * w11, dP[1] * w11, dP[2] * w11, dP[3] * w11);
              red4g(A.d_dline[i] + ol0 + c, dL[0] * u0, dL[1] * u0, dL[2] * u0, dL[3] * u0);
              red4g(A.d_dline[i] + ol1 + c, dL[0] * tl, dL[1] * tl, dL[2] * tl, dL[3] * tl);
            }
            dq[mat0(i)] += gx * dx; dq[mat1(i)] += gy * dy; dq[vecm(i)] += gl * dl;
          }
          float dp0, dp1, dp2;
          contract_bwd(R.o[0] + R.vd[0] * zk, R.o[1] + R.vd[1] * zk, R.o[2] + R.vd[2] * zk,
                       dq[0] * F.ainv[0], dq[1] * F.ainv[1], dq[2] * F.ainv[2], dp0, dp1, dp2);
          d_o[0] += dp0; d_o[1] += dp1; d_o[2] += dp2;
          d_vd[0] += dp0 * zk; d_vd[1] += dp1 * zk; d_vd[2] += dp2 * zk;
        }
      }
    }
    // ---- rays: vd = d/|d| and depth = sum(w z)/|d| ----------------------------------------------------
#pragma unroll
    for (int a = 0; a < 3; ++a) { d_o[a] = warp_sum(d_o[a]); d_vd[a] = warp_sum(d_vd[a]); }
    const float depth = __fdiv_rn(warp_sum(dep_p), R.nrm);
    if (lane == 0) {
      const float* acc = A.s.d_ovd + 6 * (size_t)ray;                    // colour-branch part
#pragma unroll
      for (int a = 0; a < 3; ++a) { d_o[a] += acc[a]; d_vd[a] += acc[3 + a]; }
      const float dotv = R.vd[0] * d_vd[0] + R.vd[1] * d_vd[1] + R.vd[2] * d_vd[2];
      float* out = A.d_rays + 6 * ray;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        out[a] = d_o[a];
        out[3 + a] = __fdiv_rn(d_vd[a] - R.vd[a] * dotv, R.nrm) - gd * __fdiv_rn(depth, R.nrm) * R.vd[a];
      }
    }
  }
}

// prepared block of the backward: folded layer 1 and W2 (both transposed), biases, W3
__global__ void prepare_backward_kernel(const float* __restrict__ basis, const float* __restrict__ w1,
                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                        const float* __restrict__ b2, const float* __restrict__ w3,
                                        const float* __restrict__ b3, float* __restrict__ bp) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int e = t0; e < FC * NF; e += stride) {
    const int n = e / NF, t = e - n * NF;
    float s = 0.0f;
    for (int j = 0; j < APP_DIM; ++j) s = fmaf(w1[n * APP_DIM + j], basis[j * NF + t], s);
    bp[BP_W1BT + t * FC + n] = s;
  }
  for (int e = t0; e < FC * FC; e += stride) {
    const int n = e / FC, k = e - n * FC;
    bp[BP_W2T + k * FC + n] = w2[e];
  }
  for (int e = t0; e < FC; e += stride) { bp[BP_B1 + e] = b1[e]; bp[BP_B2 + e] = b2[e]; }
  for (int e = t0; e < 3 * W3_LD; e += stride) {
    const int c = e / W3_LD, n = e - c * W3_LD;
    bp[BP_W3 + e] = n < FC + 3 ? w3[c * (FC + 3) + n] : 0.0f;
  }
  for (int e = t0; e < 4; e += stride) bp[BP_B3 + e] = e < 3 ? b3[e] : 0.0f;
}

// ---- host-side launchers --------------------------------------------------------------------------
constexpr size_t BP_TC_OFFSET = ((size_t)BP_FLOATS * sizeof(float) + 1023) & ~(size_t)1023;   // forward operand block
size_t backward_prepared_bytes() { return BP_TC_OFFSET + (size_t)PREP_BYTES; }
cudaError_t launch_prepare(const float* basis, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, unsigned char* prep,
                           cudaStream_t stream);
size_t backward_scratch_bytes(long long n_rays, int S) {
  return bwd_scratch_layout(n_rays, S, nullptr, nullptr, nullptr);
}
size_t backward_shade_smem_bytes() { return (size_t)ShadeSmem::total * sizeof(float); }

cudaError_t launch_prepare_backward(const float* basis, const float* w1, const float* b1,
                                    const float* w2, const float* b2, const float* w3,
                                    const float* b3, float* bp, cudaStream_t stream) {
  prepare_backward_kernel<<<64, 256, 0, stream>>>(basis, w1, b1, w2, b2, w3, b3, bp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_prepare(basis, w1, b1, w2, b2, w3, b3, reinterpret_cast<unsigned char*>(bp) + BP_TC_OFFSET, stream);
}

cudaError_t launch_render_backward(const FieldDev& F, BwdArgs A, void* scratch, int n_sms,
                                   cudaStream_t stream) {
  size_t zero_bytes = 0;
  bwd_scratch_layout(A.n_rays, F.S, &A.s, static_cast<unsigned char*>(scratch), &zero_bytes);
  cudaError_t e = cudaMemsetAsync(scratch, 0, zero_bytes, stream);
  if (e != cudaSuccess) return e;
  const size_t smem = backward_shade_smem_bytes();
  static bool configured[64] = {false};
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    e = cudaFuncSetAttribute(bwd_shade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const long long warps_per_cta = MARCH_THREADS / 32;
  long long want = (A.n_rays + warps_per_cta - 1) / warps_per_cta;
  const long long cap = (long long)n_sms * 8;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  bwd_march_kernel<<<grid, MARCH_THREADS, 0, stream>>>(F, A);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // tcgen05 shade step by default; LRF_BWD_TC=0 keeps the CUDA-core kernel (the independent fp32
  // implementation the tensor-core one is tested against).  Read per call so tests can switch.
  const char* tc_env = getenv("LRF_BWD_TC");
  const bool use_tc = !(tc_env && tc_env[0] == '0');
  if (use_tc) {
    static bool configured_tc[64] = {false};
    if (dev >= 0 && dev < 64 && !configured_tc[dev]) {
      e = cudaFuncSetAttribute(bwd_shade_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::total);
      if (e != cudaSuccess) return e;
      configured_tc[dev] = true;
    }
    bwd_shade_tc_kernel<<<n_sms, TC_THREADS, TcSmem::total, stream>>>(
        F, A, reinterpret_cast<const unsigned char*>(A.bp) + BP_TC_OFFSET);
  } else {
    bwd_shade_kernel<<<n_sms, SH_THREADS, smem, stream>>>(F, A);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  bwd_density_kernel<<<grid, MARCH_THREADS, 0, stream>>>(F, A);
  return cudaGetLastError();
}

cudaError_t launch_render_backward_abi(const FieldDev& F, const float* rays, long long n_rays,
                                       int white_bg, const float* g_rgb, const float* g_depth,
                                       const float* bp, const LrfGradients& G, void* scratch,
                                       int n_sms, cudaStream_t stream) {
  BwdArgs A;
  A.n_rays = n_rays; A.rays = rays; A.g_rgb = g_rgb; A.g_depth = g_depth; A.bp = bp;
  A.white_bg = white_bg; A.d_rays = G.d_rays;
  for (int i = 0; i < 3; ++i) {
    A.d_dplane[i] = G.d_dplane[i]; A.d_dline[i] = G.d_dline[i];
    A.d_aplane[i] = G.d_aplane[i]; A.d_aline[i] = G.d_aline[i];
  }
  A.d_w1b = G.d_w1b; A.d_b1 = G.d_b1; A.d_w2 = G.d_w2; A.d_b2 = G.d_b2; A.d_w3 = G.d_w3; A.d_b3 = G.d_b3;
  return launch_render_backward(F, A, scratch, n_sms, stream);
}

}  // namespace lrf
