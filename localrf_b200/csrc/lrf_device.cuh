// lrf_device.cuh -- device-side building blocks shared by the kernels: PTX wrappers (mbarrier, TMA
// bulk copy, tcgen05), ray setup, the VM density / appearance gathers, warp scans.
#pragma once
#include <cuda_bf16.h>

#include "lrf_common.cuh"

namespace lrf {

// Experiment switch (DESIGN.md par. 10, code size): LRF_ROLL_PLANES=1 keeps the loop over the three planes of the
// VM gathers ROLLED (a third of the gather code, plane geometry indexed at run time); 0 = fully unrolled (default).
#ifndef LRF_ROLL_PLANES
#define LRF_ROLL_PLANES 0
#endif
#if LRF_ROLL_PLANES
#define LRF_PLANE_LOOP _Pragma("unroll 1")
#else
#define LRF_PLANE_LOOP _Pragma("unroll")
#endif

constexpr int TM = 128;      // appearance samples per MLP sub-tile (= UMMA M)
constexpr int TMEM_COLS = 512;   // power of two >= 384 used columns
constexpr int TM_ACC1 = 0;       // fp32 accumulator of layer 1   [0,128)
constexpr int TM_ACC2 = 128;     // fp32 accumulator of layer 2   [128,256)
constexpr int TM_A2HI = 256;     // layer-2 A operand, bf16 hi: 128 K-elements = 64 columns
constexpr int TM_A2LO = 320;     // layer-2 A operand, bf16 lo
constexpr int TM_ACC0 = 384;     // (fields with positional encodings) fp32 accumulator of basis_mat, 32 columns
// tcgen05 instruction descriptor, kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FC >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
constexpr float T_EPS = 1e-10f;  // early-termination transmittance (error bound)

struct RaySm {                     // one ray, as set up by setup_ray
  float o[3];                      // origin in the field's frame
  float vd[3];                     // normalised direction (tensorBase.py:578-580)
  float nrm;                       // |d| before normalisation (depth is divided by it, :615)
  float blend;                     // this field's blending weight for the ray's view
};

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
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- PTX helpers: tcgen05 (TMEM allocation, MMA, commit, TMEM load, fences) ----------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE (8x8 core matrices of 128 contiguous bytes)
// lbo = byte stride between the two K-adjacent core matrices, sbo = between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
         ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// same with the A operand in TMEM (lane = row, 32-bit column c = K elements 2c, 2c+1)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// the same two forms with an explicit instruction descriptor (other N, MN-major operands)
__host__ __device__ constexpr uint32_t umma_idesc(int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_ss_id(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_id(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns from registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with hi, lo bf16 (round-to-nearest): ~16 mantissa bits.  Packs two values per word.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
// stores 8 consecutive K elements (one 16-byte chunk) of row `row` into the hi and lo operands
__device__ __forceinline__ void store_chunk(unsigned char* hi_base, unsigned char* lo_base, int row,
                                            int kc, int chunks, const float* v) {
  uint4 h, l;
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
  const int off = (((row >> 3) * chunks + kc) * 8 + (row & 7)) * 16;
  *reinterpret_cast<uint4*>(hi_base + off) = h;
  *reinterpret_cast<uint4*>(lo_base + off) = l;
}

// One elected thread: D = Ahi*Bhi^T + Ahi*Blo^T + Alo*Bhi^T over `ksteps` K-steps of 16, then commit.
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo,
                                            uint32_t b_hi, uint32_t b_lo, int ksteps, int chunks,
                                            uint32_t bar) {
  const uint32_t sbo = (uint32_t)chunks * 128u;
  uint32_t acc = 0;
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t ko = (uint32_t)ks * 256u;   // two 128-byte core matrices per K-step
    const uint64_t ah = umma_desc(a_hi + ko, 128u, sbo), al = umma_desc(a_lo + ko, 128u, sbo);
    const uint64_t bh = umma_desc(b_hi + ko, 128u, sbo), bl = umma_desc(b_lo + ko, 128u, sbo);
    umma_bf16(d_tmem, ah, bh, acc);
    umma_bf16(d_tmem, ah, bl, 1u);
    umma_bf16(d_tmem, al, bh, 1u);
    acc = 1u;
  }
  umma_commit(bar);
}

// Layer 2: A (hi/lo) in TMEM, B in shared memory.
__device__ __forceinline__ void issue_layer_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo,
                                               uint32_t b_hi, uint32_t b_lo, int ksteps, int chunks,
                                               uint32_t bar) {
  const uint32_t sbo = (uint32_t)chunks * 128u;
  uint32_t acc = 0;
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t ko = (uint32_t)ks * 256u, kc = (uint32_t)ks * 8u;   // 16 bf16 = 8 TMEM columns
    const uint64_t bh = umma_desc(b_hi + ko, 128u, sbo), bl = umma_desc(b_lo + ko, 128u, sbo);
    umma_bf16_ts(d_tmem, a_hi + kc, bh, acc);
    umma_bf16_ts(d_tmem, a_hi + kc, bl, 1u);
    umma_bf16_ts(d_tmem, a_lo + kc, bh, 1u);
    acc = 1u;
  }
  umma_commit(bar);
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
    if (B.ij) { B.ij[2 * r] = col; B.ij[2 * r + 1] = row; }
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
LRF_PLANE_LOOP
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

// ---- 16-bit grid storage (LrfField.grid_dtype = LRF_GRID_BF16) ----------------------------------
// The same gathers reading bfloat16 texels in the same [H][W][C] / [L][C] layout: a density texel is ONE
// 16-byte load (8 components), an appearance texel three (24 components) -- half the sectors and half the
// load instructions of the fp32 grids.  bf16 -> fp32 is exact (a 16-bit shift) and the arithmetic is the fp32
// code's, expression for expression, so a field whose parameters are bf16-representable renders the same
// values from either storage.
struct Tex8 { float v[8]; };
__device__ __forceinline__ Tex8 ldg8_bf16(const __nv_bfloat16* p) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  Tex8 t;
  t.v[0] = __uint_as_float(r.x << 16); t.v[1] = __uint_as_float(r.x & 0xffff0000u);
  t.v[2] = __uint_as_float(r.y << 16); t.v[3] = __uint_as_float(r.y & 0xffff0000u);
  t.v[4] = __uint_as_float(r.z << 16); t.v[5] = __uint_as_float(r.z & 0xffff0000u);
  t.v[6] = __uint_as_float(r.w << 16); t.v[7] = __uint_as_float(r.w & 0xffff0000u);
  return t;
}

// compute_densityfeature for one point, bf16 texels (cf. density_feature)
__device__ __forceinline__ float density_feature_bf16(const FieldDev& F, const float* q) {
  static_assert(CD == 8, "one 16-byte load per density texel");
  float sigma = 0.0f;
LRF_PLANE_LOOP
  for (int i = 0; i < 3; ++i) {
    const int W = F.g[mat0(i)], H = F.g[mat1(i)], L = F.g[vecm(i)];
    int x0, x1, y0, y1, l0, l1;
    float tx, ty, tl;
    grid_coord(q[mat0(i)], W, x0, x1, tx);
    grid_coord(q[mat1(i)], H, y0, y1, ty);
    grid_coord(q[vecm(i)], L, l0, l1, tl);
    const __nv_bfloat16* P = reinterpret_cast<const __nv_bfloat16*>(F.dplane[i]);
    const __nv_bfloat16* Ln = reinterpret_cast<const __nv_bfloat16*>(F.dline[i]);
    const Tex8 a = ldg8_bf16(P + ((size_t)y0 * W + x0) * CD), b = ldg8_bf16(P + ((size_t)y0 * W + x1) * CD);
    const Tex8 c = ldg8_bf16(P + ((size_t)y1 * W + x0) * CD), d = ldg8_bf16(P + ((size_t)y1 * W + x1) * CD);
    const Tex8 u = ldg8_bf16(Ln + (size_t)l0 * CD), v = ldg8_bf16(Ln + (size_t)l1 * CD);
    float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
    float w10 = (1.0f - tx) * ty, w11 = tx * ty;
    float s = 0.0f;
#pragma unroll
    for (int e = 0; e < CD; ++e) {
      float px = a.v[e] * w00 + b.v[e] * w01 + c.v[e] * w10 + d.v[e] * w11;
      s += px * (u.v[e] * (1.0f - tl) + v.v[e] * tl);
    }
    sigma += s;
  }
  return sigma;
}

// one plane's 24 appearance features of one point, bf16 texels (cf. app_plane_features)
__device__ __forceinline__ void app_plane_features_bf16(const FieldDev& F, int i, const float* q,
                                                        float* out /*[CA]*/) {
  static_assert(CA % 8 == 0, "16-byte loads of 8 components");
  const int W = F.g[mat0(i)], H = F.g[mat1(i)], L = F.g[vecm(i)];
  int x0, x1, y0, y1, l0, l1;
  float tx, ty, tl;
  grid_coord(q[mat0(i)], W, x0, x1, tx);
  grid_coord(q[mat1(i)], H, y0, y1, ty);
  grid_coord(q[vecm(i)], L, l0, l1, tl);
  const __nv_bfloat16* P = reinterpret_cast<const __nv_bfloat16*>(F.aplane[i]);
  const __nv_bfloat16* p00 = P + ((size_t)y0 * W + x0) * CA;
  const __nv_bfloat16* p01 = P + ((size_t)y0 * W + x1) * CA;
  const __nv_bfloat16* p10 = P + ((size_t)y1 * W + x0) * CA;
  const __nv_bfloat16* p11 = P + ((size_t)y1 * W + x1) * CA;
  const __nv_bfloat16* q0 = reinterpret_cast<const __nv_bfloat16*>(F.aline[i]) + (size_t)l0 * CA;
  const __nv_bfloat16* q1 = reinterpret_cast<const __nv_bfloat16*>(F.aline[i]) + (size_t)l1 * CA;
  float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
  float w10 = (1.0f - tx) * ty, w11 = tx * ty;
#pragma unroll
  for (int c8 = 0; c8 < CA / 8; ++c8) {
    const Tex8 a = ldg8_bf16(p00 + 8 * c8), b = ldg8_bf16(p01 + 8 * c8), c = ldg8_bf16(p10 + 8 * c8),
               d = ldg8_bf16(p11 + 8 * c8);
    const Tex8 u = ldg8_bf16(q0 + 8 * c8), v = ldg8_bf16(q1 + 8 * c8);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      out[8 * c8 + e] = (a.v[e] * w00 + b.v[e] * w01 + c.v[e] * w10 + d.v[e] * w11) *
                        (u.v[e] * (1.0f - tl) + v.v[e] * tl);
  }
}

// storage-type dispatch (compile time): H16 = bf16 texels
template <bool H16>
__device__ __forceinline__ float density_feature_t(const FieldDev& F, const float* q) {
  if constexpr (H16) return density_feature_bf16(F, q);
  else return density_feature(F, q);
}
template <bool H16>
__device__ __forceinline__ void app_plane_features_t(const FieldDev& F, int i, const float* q, float* out) {
  if constexpr (H16) app_plane_features_bf16(F, i, q, out);
  else app_plane_features(F, i, q, out);
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



}  // namespace lrf
