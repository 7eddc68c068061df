// lrf_render.cu -- the fused per-ray-batch render kernel (sm_100a).
//
// One launch = one (field, ray batch): ray generation -> contracted sampling -> VM density gather
// -> softplus/alpha/transmittance scan -> (floater filter) -> compaction of the samples whose
// weight exceeds the threshold -> VM appearance gather -> basis+MLP -> composite -> blend /
// exposure / clamp.  Replaces TensorBase.forward (models/tensorBase.py:567-636) and the per-field
// body of LocalTensorfs.forward (local_tensorfs.py:397-497).
//
// CTA = 8 warps, persistent over "ray tiles" of 8 rays.
//   phase 1 (density march): warp = ray, lane = sample.  32 consecutive samples of one ray are
//            ~14 voxels of path, so the warp's texel fetches fall into few 128-byte lines; the
//            transmittance is a warp-shuffle product scan with a carried prefix.
//   phase 2 (appearance): the tile's surviving samples are compacted (deterministically, per-ray
//            segments) and shaded in sub-tiles of TM samples: gather 72 features per sample into
//            shared memory, two CTA-wide GEMMs out of shared memory (weights staged once per CTA
//            by a TMA bulk copy), layer 3 + sigmoid, weighted accumulation per ray in sample order.
#include "lrf_common.cuh"

namespace lrf {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int RT = NWARPS;   // rays per tile
constexpr int TM = 64;       // appearance samples per MLP sub-tile
constexpr float T_EPS = 1e-10f;  // early-termination transmittance (error bound)

struct RaySm {
  float o[3];
  float vd[3];
  float nrm;
  float blend;
  float rgb[3];
  float depth;
  float acc;
  int count;
  int offset;
  int valid;
};

// ---- shared-memory carve-up (dynamic) -----------------------------------------------------------
struct SmemLayout {
  int prep, x, h, rgb, sray, sw, w, alpha, klist, ray, z, mbar, total;
};

__host__ __device__ inline SmemLayout smem_layout(int S, bool floater) {
  SmemLayout L;
  int off = 0;
  int Sp = (S + 3) & ~3;
  L.prep = off;  off += PREP_FLOATS * 4;
  L.x = off;     off += NF * TM * 4;          // also reused for the layer-3 partial sums
  L.h = off;     off += FC * TM * 4;
  L.rgb = off;   off += TM * 4 * 4;
  L.sray = off;  off += TM * 4;
  L.sw = off;    off += TM * 4;
  L.w = off;     off += RT * Sp * 4;
  L.alpha = off; off += floater ? RT * Sp * 4 : 0;
  L.klist = off; off += RT * Sp * 2;
  off = (off + 15) & ~15;
  L.ray = off;   off += RT * (int)sizeof(RaySm);
  off = (off + 15) & ~15;
  L.z = off;     off += (Sp + 4) * 4;
  L.mbar = off;  off += 16;
  L.total = off;
  return L;
}

// ---- PTX helpers: mbarrier + TMA 1-D bulk copy ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- ray setup (local_tensorfs.py:397-456 / tensorBase.py:578-580) ------------------------------
__device__ __forceinline__ void setup_ray(const BatchDev& B, long long r, RaySm& R) {
  float o[3], d[3];
  long long view = 0;
  if (B.rays) {
    const float* p = B.rays + 6 * r;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    d[0] = p[3]; d[1] = p[4]; d[2] = p[5];
    if (B.rays_per_view > 0) view = r / B.rays_per_view;
  } else {
    view = r / B.rays_per_view;
    long long id = B.ray_ids[r];
    long long col = id % B.W, row = (id / B.W) % B.H;            // ids2pixel
    float i = (float)col + 0.5f, j = (float)row + 0.5f;
    float dc[3];
    if (B.fov360) {                                               // get_ray_directions_360
      const float pi = 3.14159265358979323846f;
      float phi = j * pi / (float)B.H - pi / 2.0f;
      float theta = i * 2.0f * pi / (float)B.W + pi;
      float sp, cp, st, ct;
      sincosf(phi, &sp, &cp);
      sincosf(theta, &st, &ct);
      dc[0] = cp * st; dc[1] = sp; dc[2] = cp * ct;
    } else {                                                      // get_ray_directions_lean
      float focal = B.focal, cx = B.cx, cy = B.cy;
      if (B.intrinsics) { focal = B.intrinsics[0]; cx = B.intrinsics[1]; cy = B.intrinsics[2]; }
      dc[0] = __fdiv_rn(i - cx, focal);
      dc[1] = -__fdiv_rn(j - cy, focal);
      dc[2] = -1.0f;
    }
    if (B.dirs) { B.dirs[3 * r] = dc[0]; B.dirs[3 * r + 1] = dc[1]; B.dirs[3 * r + 2] = dc[2]; }
    const float* c = B.c2w + 12 * view;
#pragma unroll
    for (int a = 0; a < 3; ++a) {                                 // get_rays_lean
      o[a] = c[a * 4 + 3] + (B.w2rf ? B.w2rf[a] : 0.0f);
      d[a] = c[a * 4 + 0] * dc[0] + c[a * 4 + 1] * dc[1] + c[a * 4 + 2] * dc[2];
    }
  }
  float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  R.nrm = n;
#pragma unroll
  for (int a = 0; a < 3; ++a) { R.o[a] = o[a]; R.vd[a] = __fdiv_rn(d[a], n); }
  R.blend = B.blend ? B.blend[view * B.blend_stride] : 1.0f;
  R.rgb[0] = R.rgb[1] = R.rgb[2] = 0.0f;
  R.depth = 0.0f; R.acc = 0.0f; R.count = 0; R.offset = 0; R.valid = 1;
}

// sample position in the field's normalised [-1,1]^3 grid coordinates (tensorBase.py:438-440,602)
__device__ __forceinline__ void sample_pos(const FieldDev& F, const RaySm& R, float z, float* p,
                                           float* q) {
  p[0] = R.o[0] + R.vd[0] * z; p[1] = R.o[1] + R.vd[1] * z; p[2] = R.o[2] + R.vd[2] * z;
  contract(p[0], p[1], p[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a) q[a] = (p[a] - F.amin[a]) * F.ainv[a] - 1.0f;
}

// AlphaGridMask.sample_alpha (tensorBase.py:51-58): trilinear, zero padding, align_corners=True
__device__ __forceinline__ float alpha_mask(const FieldDev& F, const float* p) {
  int D = F.ad[0], H = F.ad[1], W = F.ad[2];
  float ix = (((p[0] - F.aamin[0]) * F.aainv[0] - 1.0f + 1.0f) * 0.5f) * (float)(W - 1);
  float iy = (((p[1] - F.aamin[1]) * F.aainv[1] - 1.0f + 1.0f) * 0.5f) * (float)(H - 1);
  float iz = (((p[2] - F.aamin[2]) * F.aainv[2] - 1.0f + 1.0f) * 0.5f) * (float)(D - 1);
  float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  float tx = ix - fx, ty = iy - fy, tz = iz - fz;
  float v = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int xx = x0 + (c & 1), yy = y0 + ((c >> 1) & 1), zz = z0 + (c >> 2);
    if (xx < 0 || xx >= W || yy < 0 || yy >= H || zz < 0 || zz >= D) continue;
    float w = ((c & 1) ? tx : 1.0f - tx) * ((c & 2) ? ty : 1.0f - ty) * ((c & 4) ? tz : 1.0f - tz);
    v += __ldg(F.alpha_vol + ((size_t)zz * H + yy) * W + xx) * w;
  }
  return v;
}

// compute_densityfeature for one point (tensoRF.py:112-151), channel-last planes/lines
__device__ __forceinline__ float density_feature(const FieldDev& F, const float* q) {
  float sigma = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int W = F.g[mat0(i)], H = F.g[mat1(i)], L = F.g[vecm(i)];
    int x0, x1, y0, y1, l0, l1;
    float tx, ty, tl;
    grid_coord(q[mat0(i)], W, x0, x1, tx);
    grid_coord(q[mat1(i)], H, y0, y1, ty);
    grid_coord(q[vecm(i)], L, l0, l1, tl);
    const float* P = F.dplane[i];
    const float* p00 = P + ((size_t)y0 * W + x0) * CD;
    const float* p01 = P + ((size_t)y0 * W + x1) * CD;
    const float* p10 = P + ((size_t)y1 * W + x0) * CD;
    const float* p11 = P + ((size_t)y1 * W + x1) * CD;
    const float* q0 = F.dline[i] + (size_t)l0 * CD;
    const float* q1 = F.dline[i] + (size_t)l1 * CD;
    float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
    float w10 = (1.0f - tx) * ty, w11 = tx * ty;
    float s = 0.0f;
#pragma unroll
    for (int h = 0; h < CD / 4; ++h) {
      float4 a = ldg4(p00 + 4 * h), b = ldg4(p01 + 4 * h), c = ldg4(p10 + 4 * h),
             d = ldg4(p11 + 4 * h);
      float4 u = ldg4(q0 + 4 * h), v = ldg4(q1 + 4 * h);
      float px = a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11;
      float py = a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11;
      float pz = a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11;
      float pw = a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11;
      s += px * (u.x * (1.0f - tl) + v.x * tl);
      s += py * (u.y * (1.0f - tl) + v.y * tl);
      s += pz * (u.z * (1.0f - tl) + v.z * tl);
      s += pw * (u.w * (1.0f - tl) + v.w * tl);
    }
    sigma += s;
  }
  return sigma;
}

// one plane's 24 appearance features of one point (tensoRF.py:153-194): out[c] = plane_c * line_c
__device__ __forceinline__ void app_plane_features(const FieldDev& F, int i, const float* q,
                                                   float* out /*[CA]*/) {
  const int W = F.g[mat0(i)], H = F.g[mat1(i)], L = F.g[vecm(i)];
  int x0, x1, y0, y1, l0, l1;
  float tx, ty, tl;
  grid_coord(q[mat0(i)], W, x0, x1, tx);
  grid_coord(q[mat1(i)], H, y0, y1, ty);
  grid_coord(q[vecm(i)], L, l0, l1, tl);
  const float* P = F.aplane[i];
  const float* p00 = P + ((size_t)y0 * W + x0) * CA;
  const float* p01 = P + ((size_t)y0 * W + x1) * CA;
  const float* p10 = P + ((size_t)y1 * W + x0) * CA;
  const float* p11 = P + ((size_t)y1 * W + x1) * CA;
  const float* q0 = F.aline[i] + (size_t)l0 * CA;
  const float* q1 = F.aline[i] + (size_t)l1 * CA;
  float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
  float w10 = (1.0f - tx) * ty, w11 = tx * ty;
#pragma unroll
  for (int h = 0; h < CA / 4; ++h) {
    float4 a = ldg4(p00 + 4 * h), b = ldg4(p01 + 4 * h), c = ldg4(p10 + 4 * h),
           d = ldg4(p11 + 4 * h);
    float4 u = ldg4(q0 + 4 * h), v = ldg4(q1 + 4 * h);
    out[4 * h + 0] = (a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11) * (u.x * (1.0f - tl) + v.x * tl);
    out[4 * h + 1] = (a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11) * (u.y * (1.0f - tl) + v.y * tl);
    out[4 * h + 2] = (a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11) * (u.z * (1.0f - tl) + v.z * tl);
    out[4 * h + 3] = (a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11) * (u.w * (1.0f - tl) + v.w * tl);
  }
}

// ---- warp primitives ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive product scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}

// weights of one ray from its alphas in shared memory: w[k] = alpha[k] * prod_{j<k}(1-alpha[j]+1e-10)
// (alpha2weights, tensorBase.py:23-32).  Returns nothing; writes w_s.
__device__ __forceinline__ void rescan_weights(const float* alpha_s, float* w_s, int S, int lane) {
  float carry = 1.0f;
  for (int k0 = 0; k0 < S; k0 += 32) {
    int k = k0 + lane;
    float a = (k < S) ? alpha_s[k] : 0.0f;
    if (k == S - 1) a = 1.0f;
    float f = (k < S) ? (1.0f - a) + 1e-10f : 1.0f;
    float inc = warp_scan_mul(f, lane);
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.0f;
    if (k < S) w_s[k] = a * (carry * exc);
    carry *= __shfl_sync(0xffffffffu, inc, 31);
  }
}

// ---- CTA-wide GEMM out of shared memory ---------------------------------------------------------
// acc[i][j] = sum_k A[k][m0+i] * Wt[k][n0+j],  A: [K][TM], Wt: [K][FC];  256 threads cover TM x FC
template <int K>
__device__ __forceinline__ void cta_gemm(const float* __restrict__ A, const float* __restrict__ Wt,
                                         int m0, int n0, float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float4 a = *reinterpret_cast<const float4*>(A + k * TM + m0);
    float4 w0 = *reinterpret_cast<const float4*>(Wt + k * FC + n0);
    float4 w1 = *reinterpret_cast<const float4*>(Wt + k * FC + n0 + 4);
    float av[4] = {a.x, a.y, a.z, a.w};
    float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}

// =================================================================================================
__global__ void __launch_bounds__(THREADS, 1)
render_kernel(const FieldDev F, const BatchDev B) {
  extern __shared__ __align__(128) unsigned char smem[];
  const bool floater = B.floater_thresh > 0.0f;
  const SmemLayout L = smem_layout(F.S, floater);
  float* prep_s = reinterpret_cast<float*>(smem + L.prep);
  float* x_s = reinterpret_cast<float*>(smem + L.x);
  float* part_s = x_s;  // layer-3 partial sums reuse the feature buffer
  float* h_s = reinterpret_cast<float*>(smem + L.h);
  float* rgb_s = reinterpret_cast<float*>(smem + L.rgb);
  int* sray_s = reinterpret_cast<int*>(smem + L.sray);
  float* sw_s = reinterpret_cast<float*>(smem + L.sw);
  float* w_all = reinterpret_cast<float*>(smem + L.w);
  float* alpha_all = reinterpret_cast<float*>(smem + L.alpha);
  unsigned short* klist_all = reinterpret_cast<unsigned short*>(smem + L.klist);
  RaySm* ray_s = reinterpret_cast<RaySm*>(smem + L.ray);
  float* z_s = reinterpret_cast<float*>(smem + L.z);
  const uint32_t mbar = smem_u32(smem + L.mbar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = F.S, Sp = (S + 3) & ~3;

  // -- stage the prepared MLP weights once per CTA with TMA bulk copies, and the z table ----------
  if (tid == 0) {
    mbar_init(mbar, 1);
    constexpr uint32_t bytes = PREP_FLOATS * 4;
    mbar_expect_tx(mbar, bytes);
    constexpr uint32_t CH = 32768;  // keep each bulk copy modest
    for (uint32_t o = 0; o < bytes; o += CH)
      tma_bulk_g2s(smem_u32(prep_s) + o, reinterpret_cast<const char*>(F.prep) + o,
                   min(CH, bytes - o), mbar);
  }
  for (int k = tid; k < S; k += THREADS) z_s[k] = F.z[k];
  if (tid == 0) z_s[S] = F.z[S - 1];  // dist of the last sample = 0 (tensorBase.py:584-587)
  __syncthreads();
  bool prep_ready = false;

  const float* W1B_s = prep_s + PREP_W1B;
  const float* W2T_s = prep_s + PREP_W2T;
  const float* b1_s = prep_s + PREP_B1;
  const float* b2_s = prep_s + PREP_B2;
  const float* W3_s = prep_s + PREP_W3;
  const float* b3_s = prep_s + PREP_B3;

  float* w_s = w_all + warp * Sp;
  float* alpha_s = alpha_all + warp * Sp;
  unsigned short* klist_s = klist_all + warp * Sp;

  const long long n_tiles = (B.n_rays + RT - 1) / RT;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ============================ phase 1: density march (warp = ray) ===========================
    const long long ray = tile * RT + warp;
    const bool have_ray = ray < B.n_rays;
    RaySm R;
    if (have_ray) {
      setup_ray(B, ray, R);
    } else {
      R.valid = 0; R.count = 0; R.offset = 0;
      R.rgb[0] = R.rgb[1] = R.rgb[2] = R.depth = R.acc = 0.0f; R.blend = 0.0f; R.nrm = 1.0f;
      R.o[0] = R.o[1] = R.o[2] = 0.0f; R.vd[0] = R.vd[1] = R.vd[2] = 0.0f;
    }
    int marched = 0;
    if (have_ray) {
      float carry = 1.0f, acc_p = 0.0f, dep_p = 0.0f, idx_p = 0.0f;
      int k0 = 0;
      for (; k0 < S; k0 += 32) {
        const int k = k0 + lane;
        float alpha = 0.0f;
        if (k < S) {
          const float z = z_s[k];
          float p[3], q[3];
          sample_pos(F, R, z, p, q);
          bool valid = (k != S - 1);                                  // ray_valid[:, -1] = 0
          if (valid && F.alpha_vol) valid = alpha_mask(F, p) > 0.0f;  // tensorBase.py:593-598
          float sigma = 0.0f;
          if (valid) {
            sigma = feature2density(density_feature(F, q), F.density_shift, F.act);
            ++marched;
          }
          const float dist = z_s[k + 1] - z;
          alpha = -expm1f(-sigma * dist * F.distance_scale);          // 1 - exp(-sigma*dist*scale)
          if (k == S - 1) alpha = 1.0f;                               // alpha[:, -1] = 1
        }
        const float f = (k < S) ? (1.0f - alpha) + 1e-10f : 1.0f;
        const float inc = warp_scan_mul(f, lane);
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.0f;
        const float wgt = alpha * (carry * exc);
        if (k < S) {
          w_s[k] = wgt;
          if (floater) alpha_s[k] = alpha;
          acc_p += wgt;
          dep_p += wgt * z_s[k];
          idx_p += wgt * (float)k;
        }
        carry *= __shfl_sync(0xffffffffu, inc, 31);
        if (!floater && carry < T_EPS) { k0 += 32; break; }           // early ray termination
      }
      for (int k = k0 + lane; k < S; k += 32) w_s[k] = 0.0f;          // terminated tail
      const float acc = warp_sum(acc_p), dep = warp_sum(dep_p);
      R.acc = acc;
      R.depth = __fdiv_rn(dep, R.nrm);                                // tensorBase.py:615
      if (floater) {                                                  // tensorBase.py:617-620
        const float lim = warp_sum(idx_p) * B.floater_thresh;
        __syncwarp();
        for (int k = lane; k < S; k += 32)
          if ((float)k < lim) alpha_s[k] = 0.0f;
        __syncwarp();
        rescan_weights(alpha_s, w_s, S, lane);
      }
      __syncwarp();
      // compact the samples with weight > threshold (tensorBase.py:622), in sample order
      int cnt = 0;
      for (int kb = 0; kb < S; kb += 32) {
        const int k = kb + lane;
        const bool on = (k < S) && (w_s[k] > F.weight_thres);
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) klist_s[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
        cnt += __popc(m);
      }
      R.count = cnt;
      if (B.weights) {
        float* wo = B.weights + (size_t)ray * S;
        for (int k = lane; k < S; k += 32) wo[k] = w_s[k];
      }
    }
    if (lane == 0) ray_s[warp] = R;
    __syncthreads();
    if (tid == 0) {
      int off = 0;
      for (int r = 0; r < RT; ++r) { ray_s[r].offset = off; off += ray_s[r].count; }
    }
    if (B.stats) {
      const int mt = (int)warp_sum((float)marched);   // <= 32 * S: exact in fp32
      if (lane == 0 && mt > 0) atomicAdd(B.stats, (unsigned long long)mt);
    }
    __syncthreads();
    const int total = ray_s[RT - 1].offset + ray_s[RT - 1].count;
    if (B.stats && tid == 0) atomicAdd(B.stats + 1, (unsigned long long)total);

    // ============================ phase 2: appearance + MLP ======================================
    if (total > 0 && !prep_ready) { mbar_wait(mbar, 0); prep_ready = true; }
    for (int j0 = 0; j0 < total; j0 += TM) {
      // -- gather: thread = (plane, sample) ------------------------------------------------------
      if (tid < 3 * TM) {
        const int pl = tid / TM, m = tid - pl * TM;
        const int j = j0 + m;
        float feat[CA];
        if (j < total) {
          int r = 0;
#pragma unroll
          for (int t = 1; t < RT; ++t) r += (j >= ray_s[t].offset) ? 1 : 0;
          const int k = klist_all[r * Sp + (j - ray_s[r].offset)];
          float p[3], q[3];
          sample_pos(F, ray_s[r], z_s[k], p, q);
          if (pl == 0) app_plane_features(F, 0, q, feat);
          else if (pl == 1) app_plane_features(F, 1, q, feat);
          else app_plane_features(F, 2, q, feat);
          if (pl == 0) { sray_s[m] = r; sw_s[m] = w_all[r * Sp + k]; }
        } else {
#pragma unroll
          for (int c = 0; c < CA; ++c) feat[c] = 0.0f;
          if (pl == 0) { sray_s[m] = -1; sw_s[m] = 0.0f; }
        }
#pragma unroll
        for (int c = 0; c < CA; ++c) x_s[(pl * CA + c) * TM + m] = feat[c];
      }
      __syncthreads();
      // -- layer 1 (basis folded in): h1 = relu(W1B^T x + b1) ------------------------------------
      const int m0 = (tid & 15) * 4, n0 = (tid >> 4) * 8;
      float acc[4][8];
      cta_gemm<NF>(x_s, W1B_s, m0, n0, acc);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const float b = b1_s[n0 + jn];
        float4 v = make_float4(fmaxf(acc[0][jn] + b, 0.0f), fmaxf(acc[1][jn] + b, 0.0f),
                               fmaxf(acc[2][jn] + b, 0.0f), fmaxf(acc[3][jn] + b, 0.0f));
        *reinterpret_cast<float4*>(h_s + (n0 + jn) * TM + m0) = v;
      }
      __syncthreads();
      // -- layer 2: h2 = relu(W2 h1 + b2), kept in registers; layer-3 partials --------------------
      cta_gemm<FC>(h_s, W2T_s, m0, n0, acc);
      {
        float part[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) part[i][0] = part[i][1] = part[i][2] = 0.0f;
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) {
          const float b = b2_s[n0 + jn];
          const float w30 = W3_s[0 * W3_LD + n0 + jn], w31 = W3_s[1 * W3_LD + n0 + jn],
                      w32 = W3_s[2 * W3_LD + n0 + jn];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float h = fmaxf(acc[i][jn] + b, 0.0f);
            part[i][0] = fmaf(w30, h, part[i][0]);
            part[i][1] = fmaf(w31, h, part[i][1]);
            part[i][2] = fmaf(w32, h, part[i][2]);
          }
        }
        const int tn = tid >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 3; ++c) part_s[(tn * TM + m0 + i) * 3 + c] = part[i][c];
      }
      __syncthreads();
      // -- layer 3 tail: + view direction, bias, sigmoid (tensorBase.py:126-133) ------------------
      if (tid < 3 * TM) {
        const int m = tid / 3, c = tid - 3 * m;
        float s = 0.0f;
#pragma unroll
        for (int tn = 0; tn < 16; ++tn) s += part_s[(tn * TM + m) * 3 + c];
        const int r = sray_s[m];
        if (r >= 0) {
          const RaySm& Rr = ray_s[r];
          s += W3_s[c * W3_LD + FC] * Rr.vd[0] + W3_s[c * W3_LD + FC + 1] * Rr.vd[1] +
               W3_s[c * W3_LD + FC + 2] * Rr.vd[2];
          s += b3_s[c];
          rgb_s[m * 4 + c] = __fdiv_rn(1.0f, 1.0f + expf(-s));
        } else {
          rgb_s[m * 4 + c] = 0.0f;
        }
      }
      __syncthreads();
      // -- composite: rgb_map += w * rgb, per ray in sample order (tensorBase.py:632) -------------
      if (tid < RT * 3) {
        const int r = tid / 3, c = tid - 3 * r;
        float a = ray_s[r].rgb[c];
        const int lo = max(ray_s[r].offset - j0, 0);
        const int hi = min(ray_s[r].offset + ray_s[r].count - j0, TM);
        for (int m = lo; m < hi; ++m) a = fmaf(sw_s[m], rgb_s[m * 4 + c], a);
        ray_s[r].rgb[c] = a;
      }
      __syncthreads();
    }

    // ============================ outputs (local_tensorfs.py:467-497) ===========================
    if (tid < RT) {
      const long long r = tile * RT + tid;
      if (r < B.n_rays) {
        const RaySm& Rr = ray_s[tid];
        float c[3];
        const float bg = B.white_bg ? (1.0f - Rr.acc) : 0.0f;       // tensorBase.py:633-634
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = (Rr.rgb[a] + bg) * Rr.blend;
        float dpt = Rr.depth * Rr.blend;
        if (B.accumulate) {
#pragma unroll
          for (int a = 0; a < 3; ++a) c[a] = B.rgb[3 * r + a] + c[a];
          dpt = B.depth[r] + dpt;
        }
        if (B.finalize) {
          if (B.exposure) {
            const long long view = B.rays_per_view > 0 ? r / B.rays_per_view : 0;
            const float* E = B.exposure + 9 * view;
            float o0 = E[0] * c[0] + E[1] * c[1] + E[2] * c[2];
            float o1 = E[3] * c[0] + E[4] * c[1] + E[5] * c[2];
            float o2 = E[6] * c[0] + E[7] * c[1] + E[8] * c[2];
            c[0] = o0; c[1] = o1; c[2] = o2;
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) c[a] = fminf(1.0f, fmaxf(0.0f, c[a]));
        }
        B.rgb[3 * r] = c[0]; B.rgb[3 * r + 1] = c[1]; B.rgb[3 * r + 2] = c[2];
        B.depth[r] = dpt;
      }
    }
    __syncthreads();
  }
  // never leave with the bulk copy still in flight
  if (!prep_ready) mbar_wait(mbar, 0);
}

// ---- small kernels ------------------------------------------------------------------------------
// prepared block: W1B = (W1 @ basis)^T, W2^T, biases, W3
__global__ void prepare_kernel(const float* __restrict__ basis, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ w2,
                               const float* __restrict__ b2, const float* __restrict__ w3,
                               const float* __restrict__ b3, float* __restrict__ prep) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int e = t; e < NF * FC; e += stride) {
    const int k = e / FC, n = e - k * FC;
    float s = 0.0f;
    for (int j = 0; j < APP_DIM; ++j) s = fmaf(w1[n * APP_DIM + j], basis[j * NF + k], s);
    prep[PREP_W1B + e] = s;
  }
  for (int e = t; e < FC * FC; e += stride) {
    const int k = e / FC, n = e - k * FC;
    prep[PREP_W2T + e] = w2[n * FC + k];
  }
  for (int e = t; e < FC; e += stride) { prep[PREP_B1 + e] = b1[e]; prep[PREP_B2 + e] = b2[e]; }
  for (int e = t; e < 3 * W3_LD; e += stride) {
    const int c = e / W3_LD, n = e - c * W3_LD;
    prep[PREP_W3 + e] = n < FC + 3 ? w3[c * (FC + 3) + n] : 0.0f;
  }
  for (int e = t; e < 4; e += stride) prep[PREP_B3 + e] = e < 3 ? b3[e] : 0.0f;
}

__global__ void density_feature_kernel(const FieldDev F, const float* __restrict__ xyz,
                                       long long M, float* __restrict__ out) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float q[3] = {xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2]};
  out[m] = density_feature(F, q);
}

// compute_appfeature: 72 products then basis_mat (unfolded: this entry returns the 27-vector)
__global__ void app_feature_kernel(const FieldDev F, const float* __restrict__ basis,
                                   const float* __restrict__ xyz, long long M,
                                   float* __restrict__ out) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float q[3] = {xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2]};
  float feat[NF];
  app_plane_features(F, 0, q, feat);
  app_plane_features(F, 1, q, feat + CA);
  app_plane_features(F, 2, q, feat + 2 * CA);
  for (int o = 0; o < APP_DIM; ++o) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NF; ++k) s = fmaf(__ldg(basis + o * NF + k), feat[k], s);
    out[m * APP_DIM + o] = s;
  }
}

__global__ void repack_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                              long long HW) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= HW * C) return;
  long long p = e / C;
  int c = (int)(e - p * C);
  dst[e] = src[(long long)c * HW + p];
}

// ---- host-side launchers (called from lrf_abi.cu) -----------------------------------------------
size_t render_smem_bytes(int S, bool floater) { return (size_t)smem_layout(S, floater).total; }

cudaError_t launch_render(const FieldDev& F, const BatchDev& B, int n_sms, cudaStream_t stream) {
  const bool floater = B.floater_thresh > 0.0f;
  const size_t smem = render_smem_bytes(F.S, floater);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  long long n_tiles = (B.n_rays + RT - 1) / RT;
  int grid = (int)(n_tiles < n_sms ? n_tiles : n_sms);
  if (grid < 1) grid = 1;
  render_kernel<<<grid, THREADS, smem, stream>>>(F, B);
  return cudaGetLastError();
}

cudaError_t launch_prepare(const float* basis, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, float* prep,
                           cudaStream_t stream) {
  prepare_kernel<<<64, 256, 0, stream>>>(basis, w1, b1, w2, b2, w3, b3, prep);
  return cudaGetLastError();
}

cudaError_t launch_density_feature(const FieldDev& F, const float* xyz, long long M, float* out,
                                   cudaStream_t stream) {
  if (M == 0) return cudaSuccess;
  density_feature_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(F, xyz, M, out);
  return cudaGetLastError();
}

cudaError_t launch_app_feature(const FieldDev& F, const float* basis, const float* xyz, long long M,
                               float* out, cudaStream_t stream) {
  if (M == 0) return cudaSuccess;
  app_feature_kernel<<<(unsigned)((M + 127) / 128), 128, 0, stream>>>(F, basis, xyz, M, out);
  return cudaGetLastError();
}

cudaError_t launch_repack(const float* src, float* dst, int C, long long HW, cudaStream_t stream) {
  long long n = HW * C;
  if (n == 0) return cudaSuccess;
  repack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, C, HW);
  return cudaGetLastError();
}

int render_threads() { return THREADS; }

}  // namespace lrf
