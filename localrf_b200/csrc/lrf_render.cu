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
//            segments) and shaded in sub-tiles of 128 samples: gather the 72 plane x line products
//            per sample, split them into bf16 hi/lo and store them as the K-major A operand in
//            shared memory; layers 1 (basis folded in) and 2 run on the 5th-gen tensor cores
//            (tcgen05.mma kind::f16, M=128 N=128, three bf16 products hi*hi + hi*lo + lo*hi per
//            layer = ~16 mantissa bits, fp32 accumulators in TMEM); the epilogues read TMEM with
//            tcgen05.ld, apply bias/ReLU, re-split for the next layer, and finish layer 3 +
//            sigmoid + weighted accumulation per ray in sample order on the CUDA cores.  The
//            weight operands are staged once per CTA by TMA bulk copies.
#include <cuda_bf16.h>

#include "lrf_common.cuh"

namespace lrf {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int RT = NWARPS;   // rays per tile
constexpr int TM = 128;      // appearance samples per MLP sub-tile (= UMMA M)
constexpr int MLP_S = 192;       // pseudo sample count sizing mlp_kernel's scratch (>= TM*3*4 bytes)
constexpr int TMEM_COLS = 512;   // power of two >= 384 used columns
constexpr int TM_ACC1 = 0;       // fp32 accumulator of layer 1   [0,128)
constexpr int TM_ACC2 = 128;     // fp32 accumulator of layer 2   [128,256)
constexpr int TM_A2HI = 256;     // layer-2 A operand, bf16 hi: 128 K-elements = 64 columns
constexpr int TM_A2LO = 320;     // layer-2 A operand, bf16 lo
// tcgen05 instruction descriptor, kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FC >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
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
  int prep, a, part, rgb, sray, sw, w, alpha, klist, ray, z, mbar, total;
};

__host__ __device__ inline SmemLayout smem_layout(int S, bool floater) {
  SmemLayout L;
  int off = 0;
  int Sp = (S + 3) & ~3;
  L.prep = off;  off += PREP_BYTES;               // B operands (bf16 hi/lo) + fp32 tail
  off = (off + 1023) & ~1023;
  L.a = off;     off += 2 * OPER1_BYTES;          // A1 hi/lo (layer-2's A operand lives in TMEM)
  L.part = off;  off += TM * 3 * 4;               // layer-3 partial sums of the upper column half
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
  L.mbar = off;  off += 64;                       // 3 mbarriers + the TMEM base-address slot
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

// ---- shading of one sub-tile of TM samples whose A1 operand is already in shared memory ------------
// layer 1 + 2 on the tensor cores, layer 3 + sigmoid on the CUDA cores.  Must be called by all
// THREADS threads.  vd_of(m) gives the normalised view direction of sample row m (or nullptr).
struct ShadeSmem {
  unsigned char* prep;     // B operands + fp32 tail
  unsigned char* a;        // A1 (aliased) / A2 operands
  float* part;             // [TM][3]
  float* rgb;              // [TM][4]
  uint32_t bar1, bar2;     // mbarriers of the two MMA layers
  uint32_t tmem;           // TMEM base address
};

template <class ViewDir>
__device__ __forceinline__ void shade_tile(const ShadeSmem& sm, uint32_t phase, ViewDir vd_of) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* tail = reinterpret_cast<const float*>(sm.prep + PREP_TAIL);
  const float* b1_s = tail + TAIL_B1;
  const float* b2_s = tail + TAIL_B2;
  const float* W3_s = tail + TAIL_W3;
  const float* b3_s = tail + TAIL_B3;
  const uint32_t prep_a = smem_u32(sm.prep), a_a = smem_u32(sm.a);

  // A1 was written with generic-proxy stores: make it visible to the tensor core, then sync
  fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    issue_layer(sm.tmem + TM_ACC1, a_a, a_a + OPER1_BYTES, prep_a + PREP_B1HI, prep_a + PREP_B1LO,
                K1 / 16, K1_CHUNKS, sm.bar1);
  }
  mbar_wait(sm.bar1, phase);
  tc_fence_after();

  // -- epilogue 1: h1 = relu(acc1 + b1) -> bf16 hi/lo A operand of layer 2, written to TMEM ---------
  const int q = warp & 3, half = warp >> 2;       // TMEM lane quarter, column half
  const int row = q * 32 + lane;
  const uint32_t t_row = sm.tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c0 = half * 64 + cc * 32;
    float v[32];
    tmem_ld32(t_row + (uint32_t)(TM_ACC1 + c0), v);
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x0 = fmaxf(v[2 * j] + b1_s[c0 + 2 * j], 0.0f);
      const float x1 = fmaxf(v[2 * j + 1] + b1_s[c0 + 2 * j + 1], 0.0f);
      split2(x0, x1, hi[j], lo[j]);
    }
    tmem_st16(t_row + (uint32_t)(TM_A2HI + c0 / 2), hi);
    tmem_st16(t_row + (uint32_t)(TM_A2LO + c0 / 2), lo);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    issue_layer_ts(sm.tmem + TM_ACC2, sm.tmem + TM_A2HI, sm.tmem + TM_A2LO, prep_a + PREP_B2HI,
                   prep_a + PREP_B2LO, FC / 16, K2_CHUNKS, sm.bar2);
  }
  mbar_wait(sm.bar2, phase);
  tc_fence_after();

  // -- epilogue 2: h2 = relu(acc2 + b2); layer 3 (131 -> 3) + sigmoid (tensorBase.py:126-133) ------
  float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c0 = half * 64 + cc * 32;
    float v[32];
    tmem_ld32(t_row + (uint32_t)(TM_ACC2 + c0), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float h = fmaxf(v[j] + b2_s[c0 + j], 0.0f);
      p0 = fmaf(W3_s[0 * W3_LD + c0 + j], h, p0);
      p1 = fmaf(W3_s[1 * W3_LD + c0 + j], h, p1);
      p2 = fmaf(W3_s[2 * W3_LD + c0 + j], h, p2);
    }
  }
  tc_fence_before();
  if (half == 1) { sm.part[row * 3] = p0; sm.part[row * 3 + 1] = p1; sm.part[row * 3 + 2] = p2; }
  __syncthreads();
  if (half == 0) {
    const float* vd = vd_of(row);
    float s[3] = {p0 + sm.part[row * 3], p1 + sm.part[row * 3 + 1], p2 + sm.part[row * 3 + 2]};
    if (vd) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s[c] += W3_s[c * W3_LD + FC] * vd[0] + W3_s[c * W3_LD + FC + 1] * vd[1] +
                W3_s[c * W3_LD + FC + 2] * vd[2];
        s[c] += b3_s[c];
        sm.rgb[row * 4 + c] = __fdiv_rn(1.0f, 1.0f + expf(-s[c]));
      }
    } else {
      sm.rgb[row * 4] = sm.rgb[row * 4 + 1] = sm.rgb[row * 4 + 2] = 0.0f;
    }
  }
  __syncthreads();
}

// CTA prologue shared by the kernels that shade: mbarriers, TMEM allocation, weight staging.
// Returns the TMEM base address.  bars = {weights, layer 1, layer 2, tmem slot}.
__device__ __forceinline__ uint32_t shade_prologue(unsigned char* prep_s, const float* prep_g,
                                                   unsigned char* bars) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_w = smem_u32(bars), bar1 = bar_w + 8, bar2 = bar_w + 16, slot = bar_w + 24;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    constexpr uint32_t bytes = PREP_BYTES;
    mbar_expect_tx(bar_w, bytes);
    constexpr uint32_t CH = 32768;  // keep each bulk copy modest
    for (uint32_t o = 0; o < bytes; o += CH)
      tma_bulk_g2s(smem_u32(prep_s) + o, reinterpret_cast<const char*>(prep_g) + o,
                   min(CH, bytes - o), bar_w);
  }
  if (warp == 0) tmem_alloc(slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(bars + 24);
}

// =================================================================================================
__global__ void __launch_bounds__(THREADS, 1)
render_kernel(const FieldDev F, const BatchDev B) {
  extern __shared__ __align__(128) unsigned char smem[];
  const bool floater = B.floater_thresh > 0.0f;
  const SmemLayout L = smem_layout(F.S, floater);
  unsigned char* prep_s = smem + L.prep;
  unsigned char* a_s = smem + L.a;
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

  // -- mbarriers, TMEM, TMA bulk staging of the weight operands (once per CTA), the z table --------
  ShadeSmem sm;
  sm.prep = prep_s;
  sm.a = a_s;
  sm.part = reinterpret_cast<float*>(smem + L.part);
  sm.rgb = reinterpret_cast<float*>(smem + L.rgb);
  sm.bar1 = mbar + 8;
  sm.bar2 = mbar + 16;
  for (int k = tid; k < S; k += THREADS) z_s[k] = F.z[k];
  if (tid == 0) z_s[S] = F.z[S - 1];  // dist of the last sample = 0 (tensorBase.py:584-587)
  sm.tmem = shade_prologue(prep_s, F.prep, smem + L.mbar);
  bool prep_ready = false;
  uint32_t mma_phase = 0;
  float* rgb_s = sm.rgb;

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
      // -- gather: work item = (plane, sample); 24 products -> three 16-byte chunks of A1 hi/lo ----
      for (int item = tid; item < 3 * TM; item += THREADS) {
        const int pl = item / TM, m = item - pl * TM;
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
        for (int c8 = 0; c8 < CA / 8; ++c8)
          store_chunk(a_s, a_s + OPER1_BYTES, m, pl * (CA / 8) + c8, K1_CHUNKS, feat + 8 * c8);
        if (pl == 0) {   // K padding 72..79 (the region is reused by A2, so re-zero every tile)
          const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          store_chunk(a_s, a_s + OPER1_BYTES, m, K1_CHUNKS - 1, K1_CHUNKS, zero);
        }
      }
      // -- MLP on the tensor cores + layer 3 / sigmoid ---------------------------------------------
      shade_tile(sm, mma_phase, [&](int m) -> const float* {
        const int r = sray_s[m];
        return r >= 0 ? ray_s[r].vd : nullptr;
      });
      mma_phase ^= 1u;
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
  // never leave with the bulk copy still in flight; release TMEM
  if (!prep_ready) mbar_wait(mbar, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(sm.tmem, TMEM_COLS);
}

// fused basis_mat + MLPRender_Fea_late_view on explicit plane x line products (one CTA per 128 rows)
__global__ void __launch_bounds__(THREADS, 1)
mlp_kernel(const float* __restrict__ prep_g, const float* __restrict__ feats,
           const float* __restrict__ viewdirs, long long M, float* __restrict__ rgb) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout L = smem_layout(MLP_S, false);
  ShadeSmem sm;
  sm.prep = smem + L.prep;
  sm.a = smem + L.a;
  sm.part = reinterpret_cast<float*>(smem + L.part);
  sm.rgb = reinterpret_cast<float*>(smem + L.rgb);
  const uint32_t mbar = smem_u32(smem + L.mbar);
  sm.bar1 = mbar + 8;
  sm.bar2 = mbar + 16;
  float* vd_s = reinterpret_cast<float*>(smem + L.w);   // [TM][3] view directions (fits: MLP_S)
  sm.tmem = shade_prologue(sm.prep, prep_g, smem + L.mbar);
  mbar_wait(mbar, 0);
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t phase = 0;
  const long long n_tiles = (M + TM - 1) / TM;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long base = t * TM;
    for (int item = tid; item < 3 * TM; item += THREADS) {
      const int pl = item / TM, m = item - pl * TM;
      float feat[CA];
#pragma unroll
      for (int c = 0; c < CA; ++c)
        feat[c] = (base + m < M) ? feats[(base + m) * NF + pl * CA + c] : 0.0f;
#pragma unroll
      for (int c8 = 0; c8 < CA / 8; ++c8)
        store_chunk(sm.a, sm.a + OPER1_BYTES, m, pl * (CA / 8) + c8, K1_CHUNKS, feat + 8 * c8);
      if (pl == 0) {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(sm.a, sm.a + OPER1_BYTES, m, K1_CHUNKS - 1, K1_CHUNKS, zero);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          vd_s[m * 3 + c] = (base + m < M) ? viewdirs[(base + m) * 3 + c] : 0.0f;
      }
    }
    shade_tile(sm, phase, [&](int m) -> const float* { return vd_s + m * 3; });
    phase ^= 1u;
    for (int e = tid; e < TM * 3; e += THREADS) {
      const int m = e / 3, c = e - 3 * m;
      if (base + m < M) rgb[(base + m) * 3 + c] = sm.rgb[m * 4 + c];
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(sm.tmem, TMEM_COLS);
}

// ---- small kernels ------------------------------------------------------------------------------
// prepared block: bf16 hi/lo operand images of W1B = W1 @ basis and W2, fp32 biases and W3
__global__ void prepare_kernel(const float* __restrict__ basis, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ w2,
                               const float* __restrict__ b2, const float* __restrict__ w3,
                               const float* __restrict__ b3, unsigned char* __restrict__ prep) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int e = t; e < FC * K1; e += stride) {
    const int n = e / K1, k = e - n * K1;
    float s = 0.0f;
    if (k < NF)
      for (int j = 0; j < APP_DIM; ++j) s = fmaf(w1[n * APP_DIM + j], basis[j * NF + k], s);
    const __nv_bfloat16 hi = __float2bfloat16_rn(s);
    const __nv_bfloat16 lo = __float2bfloat16_rn(s - __bfloat162float(hi));
    const int off = oper_offset(n, k, K1_CHUNKS);
    *reinterpret_cast<__nv_bfloat16*>(prep + PREP_B1HI + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(prep + PREP_B1LO + off) = lo;
  }
  for (int e = t; e < FC * FC; e += stride) {
    const int n = e / FC, k = e - n * FC;
    const float s = w2[n * FC + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(s);
    const __nv_bfloat16 lo = __float2bfloat16_rn(s - __bfloat162float(hi));
    const int off = oper_offset(n, k, K2_CHUNKS);
    *reinterpret_cast<__nv_bfloat16*>(prep + PREP_B2HI + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(prep + PREP_B2LO + off) = lo;
  }
  float* tail = reinterpret_cast<float*>(prep + PREP_TAIL);
  for (int e = t; e < FC; e += stride) { tail[TAIL_B1 + e] = b1[e]; tail[TAIL_B2 + e] = b2[e]; }
  for (int e = t; e < 3 * W3_LD; e += stride) {
    const int c = e / W3_LD, n = e - c * W3_LD;
    tail[TAIL_W3 + e] = n < FC + 3 ? w3[c * (FC + 3) + n] : 0.0f;
  }
  for (int e = t; e < 4; e += stride) tail[TAIL_B3 + e] = e < 3 ? b3[e] : 0.0f;
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

cudaError_t launch_mlp(const float* prep, const float* feats, const float* viewdirs, long long M,
                       float* rgb, int n_sms, cudaStream_t stream) {
  if (M == 0) return cudaSuccess;
  const size_t smem = render_smem_bytes(MLP_S, false);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  long long n_tiles = (M + TM - 1) / TM;
  int grid = (int)(n_tiles < n_sms ? n_tiles : n_sms);
  mlp_kernel<<<grid, THREADS, smem, stream>>>(prep, feats, viewdirs, M, rgb);
  return cudaGetLastError();
}

cudaError_t launch_prepare(const float* basis, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, unsigned char* prep,
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
