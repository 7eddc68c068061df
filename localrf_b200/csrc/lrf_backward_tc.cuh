// lrf_backward_tc.cuh -- the shade step of the backward with its matrix products on tcgen05
// (MLP backward).  Included by lrf_backward.cu after bwd_shade_kernel; the default shade
// path (LRF_BWD_TC=0 selects the CUDA-core kernel, kept as the independent implementation).
// tools/umma_probe.py established on the B200 that the no-swizzle K-major images read MN-major with
// LBO = stride between core matrices along K, SBO = stride along MN, for A and B, smem and TMEM A.
//
// Tile = 128 shaded samples, 256 threads (8 warps): warp w reads TMEM lanes 32 (w & 3) .. +31
// (row = sample for the sample-row products, = output unit for the weight-gradient accumulators),
// column half w >> 2.  All steps are CTA-synchronous (one elected thread issues, one mbarrier).
// EVERY product is three bf16 MMAs on hi/lo-split fp32 operands (hi.hi + hi.lo + lo.hi, ~16 mantissa
// bits, as in the forward); products with the exact `ones` image need two.
//
//   step  product (D in TMEM)                       A                          B
//   L1    acc   = x  W1B^T          [s][n] K=80     TMEM hi/lo (x)             W1B image, K-major
//   L2    acc   = h1 W2^T           [s][n] K=128    TMEM hi/lo (h1)            W2 image,  K-major
//   G3    dW3^T += h2^T dpre        [n][c] K=s      h2 images, MN-major        dpre images, MN-major
//   G2a   acc   = dh2 W2            [s][k] K=128    TMEM hi/lo (dh2)           W2 image,  MN-major
//   G2b   dW2   += dh2^T h1         [n][k] K=s      dh2 images, MN-major       h1 images, MN-major
//   G2c   db2   += dh2^T 1          [n][.] K=s      dh2 images, MN-major       ones image
//   G1a   acc   = dh1 W1B           [s][t] K=128    TMEM hi/lo (dh1)           W1B image, MN-major
//   G1c   db1   += dh1^T 1          [n][.] K=s      dh1 images, MN-major       ones image
//   G1b   dW1B  += dh1^T x          [n][t] K=s      dh1 images, MN-major       x images,  MN-major
//
// Shared memory cannot hold the hi AND lo images of x, h1/dh1, h2/dh2 (168 KB) next to both weight
// operands (104 KB), so (1) only ONE layer's weight operand is resident at a time: a 64 KB region is
// refilled by TMA bulk copies (W1B -> W2 -> W1B per tile, 144 KB of L2 traffic per 128 samples, issued
// as soon as the previous occupant's last MMA has completed and hidden behind the epilogues), and
// (2) the x images are not kept across the tile: the scatter phase at the end of the tile, which
// re-reads the same texels for the bilinear derivatives anyway, recomputes the products and writes the
// x images into the (then idle) weight region; G1b is the tile's last product.
#pragma once

namespace lrf {

constexpr int TT = 128;                 // samples per tile
constexpr int TC_THREADS = 256;
// TMEM columns
constexpr int TC_ACC = 0;               // 128: working accumulator of the sample-row products
constexpr int TC_DW2 = 128;             // 128: dW2[n][k], resident for the whole launch
constexpr int TC_DW1 = 256;             // 80 : dW1B[n][t]
constexpr int TC_DW3 = 336;             // 16 : dW3^T[n][c]
constexpr int TC_DB2 = 352;             // 16 : column 0 = db2[n]
constexpr int TC_DB1 = 368;             // 16 : column 0 = db1[n]
constexpr int TC_AHI = 384;             // 64 : A operand, bf16 hi (K <= 128)
constexpr int TC_ALO = 448;             // 64 : A operand, bf16 lo
// PROBE: which descriptor field carries the stride between core matrices along K for an MN-major
// operand in the no-swizzle layout (CUTLASS make_umma_desc<Major::MN>, INTERLEAVE: LBO = K stride).
constexpr bool MN_LBO_IS_K_STRIDE = true;

__host__ __device__ constexpr uint32_t tc_idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TT >> 4) << 24);
}
// descriptor of an operand read MN-major out of a K-major image with `chunks` 16-byte chunks per row:
// core matrices are 128 B apart along MN and chunks*128 B apart along K
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, int chunks) {
  const uint32_t k_stride = (uint32_t)chunks * 128u, mn_stride = 128u;
  return MN_LBO_IS_K_STRIDE ? umma_desc(saddr, k_stride, mn_stride) : umma_desc(saddr, mn_stride, k_stride);
}
__device__ __forceinline__ void mma_ss_i(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_i(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 8 consecutive K elements of row `row`, bf16 hi only
__device__ __forceinline__ void store_chunk_hi(unsigned char* base, int row, int kc, int chunks, const float* v) {
  uint4 h;
  uint32_t lo;
  split2(v[0], v[1], h.x, lo); split2(v[2], v[3], h.y, lo);
  split2(v[4], v[5], h.z, lo); split2(v[6], v[7], h.w, lo);
  *reinterpret_cast<uint4*>(base + (((row >> 3) * chunks + kc) * 8 + (row & 7)) * 16) = h;
}

constexpr int W1B_BYTES = 2 * OPER1_BYTES;       // hi + lo images of W1B (40 KB), contiguous in the prepared block
constexpr int W2_BYTES = 2 * OPER2_BYTES;        // hi + lo images of W2 (64 KB)
struct TcSmem {                          // byte offsets
  static constexpr int wreg = 0;                                   // weight region: W1B | W2 | (tile end) x hi/lo images
  static constexpr int h1h = W2_BYTES;                             // h1 hi/lo images [128][128] x2, then dh1
  static constexpr int d2h = h1h + 2 * TT * FC * 2;                // h2 hi/lo images, then dh2
  static constexpr int dph = d2h + 2 * TT * FC * 2;                // dpre hi/lo images [128][16] x2
  static constexpr int ones = dph + 2 * TT * 16 * 2;               // ones image [128][16] (column 0 = 1)
  static constexpr int tail = ones + TT * 16 * 2;                  // fp32 b1, b2, W3, b3 of the prepared block
  static constexpr int q = tail + TAIL_FLOATS * 4;                 // per-sample fp32: q, praw, vd, g (x3)
  static constexpr int praw = q + TT * 12, vd = praw + TT * 12, g = vd + TT * 12;
  static constexpr int dpre = g + TT * 12;                         // [128][3]
  static constexpr int part = dpre + TT * 12;                      // [128][3] layer-3 partial sums
  static constexpr int wk = part + TT * 12, zk = wk + TT * 4, ray = zk + TT * 4, kk = ray + TT * 4;
  static constexpr int small = kk + TT * 4;                        // dW3 view-dir columns [3][3], db3 [3] (+pad)
  static constexpr int bars = small + 64;
  static constexpr int total = bars + 64;
  // fp32 staging [128][LDXF]: x before step L1 (in the h1 region), dprod after step G1a (in the h2 region)
  static constexpr int xf = h1h, dpf = d2h;
};
constexpr int LDXF = 81;
static_assert(TT * LDXF * 4 <= 2 * TT * FC * 2, "fp32 staging must fit one pair of activation images");
static_assert(W1B_BYTES <= W2_BYTES && PREP_B2HI == W1B_BYTES && PREP_B1LO == OPER1_BYTES &&
              PREP_B2LO == PREP_B2HI + OPER2_BYTES, "weight operands must be contiguous hi/lo pairs");
static_assert(TcSmem::total <= 232448, "shared-memory budget");

// three products of hi/lo operand pairs read from shared memory (both MN-major): D (+)= A^T-ish x B
__device__ __forceinline__ void mma3_ss(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl,
                                        uint32_t idesc, uint32_t first_acc) {
  mma_ss_i(d, ah, bh, idesc, first_acc);
  mma_ss_i(d, ah, bl, idesc, 1u);
  mma_ss_i(d, al, bh, idesc, 1u);
}
// TMA bulk load of `bytes` (multiple of 16) into the weight region, completion on `bar`
__device__ __forceinline__ void load_weights(uint32_t dst, const unsigned char* src, uint32_t bytes, uint32_t bar) {
  mbar_expect_tx(bar, bytes);
  constexpr uint32_t CH = 32768;
  for (uint32_t o = 0; o < bytes; o += CH) tma_bulk_g2s(dst + o, src + o, min(CH, bytes - o), bar);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
bwd_shade_tc_kernel(const FieldDev F, const BwdArgs A, const unsigned char* __restrict__ prep_g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using L = TcSmem;
  unsigned char* WREG = smem + L::wreg;                 // W1B hi|lo, or W2 hi|lo, or x hi|lo images
  unsigned char* H1H = smem + L::h1h;
  unsigned char* H1L = H1H + TT * FC * 2;
  unsigned char* D2H = smem + L::d2h;
  unsigned char* D2L = D2H + TT * FC * 2;
  unsigned char* DPH = smem + L::dph;
  unsigned char* DPL = DPH + TT * 16 * 2;
  unsigned char* ONES = smem + L::ones;
  unsigned char* XH = WREG;                             // x images live in the weight region at the tile's end
  unsigned char* XL = WREG + OPER1_BYTES;
  float* XF = reinterpret_cast<float*>(smem + L::xf);   // x staging (h1 region)
  float* DPF = reinterpret_cast<float*>(smem + L::dpf); // dprod staging (h2 region)
  float* tail_s = reinterpret_cast<float*>(smem + L::tail);
  float* q_s = reinterpret_cast<float*>(smem + L::q);
  float* praw_s = reinterpret_cast<float*>(smem + L::praw);
  float* vd_s = reinterpret_cast<float*>(smem + L::vd);
  float* g_s = reinterpret_cast<float*>(smem + L::g);
  float* dpre_s = reinterpret_cast<float*>(smem + L::dpre);
  float* part_s = reinterpret_cast<float*>(smem + L::part);
  float* wk_s = reinterpret_cast<float*>(smem + L::wk);
  float* zk_s = reinterpret_cast<float*>(smem + L::zk);
  int* ray_s = reinterpret_cast<int*>(smem + L::ray);
  int* kk_s = reinterpret_cast<int*>(smem + L::kk);
  float* small_s = reinterpret_cast<float*>(smem + L::small);       // [0..8] dW3[c][128+a], [9..11] db3
  unsigned char* bars = smem + L::bars;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qd = warp & 3, half = warp >> 2, row = qd * 32 + lane;
  const int S = F.S;
  const uint32_t bar_w = smem_u32(bars), bar = bar_w + 8, slot = bar_w + 24;
  const uint32_t wreg_a = smem_u32(WREG);
  const unsigned char* W1B_g = prep_g + PREP_B1HI;      // hi | lo, contiguous
  const unsigned char* W2_g = prep_g + PREP_B2HI;

  const long long n_app = (long long)*A.s.count;
  const long long n_tiles = (n_app + TT - 1) / TT;
  const bool has_work = (long long)blockIdx.x < n_tiles;

  // ---- prologue: barriers, first weight load, TMEM, constant images, fp32 tail ----------------------
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar, 1);
    if (has_work) load_weights(wreg_a, W1B_g, W1B_BYTES, bar_w);
  }
  for (int e = tid; e < TT * 16 * 2 / 4; e += TC_THREADS) {
    reinterpret_cast<uint32_t*>(ONES)[e] = 0u;
    reinterpret_cast<uint32_t*>(DPH)[e] = 0u;
    reinterpret_cast<uint32_t*>(DPL)[e] = 0u;
  }
  {
    const float* tail_g = reinterpret_cast<const float*>(prep_g + PREP_TAIL);
    for (int e = tid; e < TAIL_FLOATS; e += TC_THREADS) tail_s[e] = __ldg(tail_g + e);
  }
  if (tid < 16) small_s[tid] = 0.0f;
  if (warp == 0) tmem_alloc(slot, TMEM_COLS);
  __syncthreads();
  if (tid < TT)                                                       // element (row = tid, k = 0) = 1.0 (bf16 0x3F80)
    *reinterpret_cast<unsigned short*>(ONES + (((tid >> 3) * 2 + 0) * 8 + (tid & 7)) * 16) = 0x3F80;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bars + 24);
  const uint32_t t_row = tmem + ((uint32_t)(qd * 32) << 16);
  const float* b1_s = tail_s + TAIL_B1;
  const float* b2_s = tail_s + TAIL_B2;
  const float* W3_s = tail_s + TAIL_W3;
  const float* b3_s = tail_s + TAIL_B3;
  uint32_t phase = 0, wphase = 0;                                     // parities of `bar` / `bar_w`

  const int ls = tid >> 2, cq = tid & 3;
  bool first = true;                                                  // first tile of this CTA: accumulators start at 0

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long e0 = tile * TT;
    const int n_valid = (int)((n_app - e0) < TT ? (n_app - e0) : TT);
    // ---- header -------------------------------------------------------------------------------------
    if (tid < TT) {
      int ray = 0, k = 0;
      float wk = 0.0f;
      if (tid < n_valid) { ray = A.s.app_ray[e0 + tid]; k = A.s.app_k[e0 + tid]; wk = A.s.w[(size_t)ray * S + k]; }
      ray_s[tid] = ray; kk_s[tid] = k; wk_s[tid] = wk;
      RaySm R;
      load_ray(A.rays, ray, R);
      const float z = __ldg(F.z + k);
      zk_s[tid] = z;
      float p[3], qn[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) praw_s[tid * 3 + a] = R.o[a] + R.vd[a] * z;
      sample_pos(F, R, z, p, qn);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        q_s[tid * 3 + a] = qn[a];
        vd_s[tid * 3 + a] = R.vd[a];
        g_s[tid * 3 + a] = tid < n_valid ? A.g_rgb[3 * (size_t)ray + a] : 0.0f;
      }
    }
    __syncthreads();
    // ---- gather: the 72 products -> fp32 staging (two passes of 64 samples) -----------------------------
    for (int pass = 0; pass < 2; ++pass) {
      const int sm_ = pass * 64 + ls;
      const float qq[3] = {q_s[sm_ * 3], q_s[sm_ * 3 + 1], q_s[sm_ * 3 + 2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int W = F.g[mat0(i)], H = F.g[mat1(i)], Ln = F.g[vecm(i)];
        int x0, x1, y0, y1, l0, l1; float tx, ty, tl;
        grid_coord(qq[mat0(i)], W, x0, x1, tx);
        grid_coord(qq[mat1(i)], H, y0, y1, ty);
        grid_coord(qq[vecm(i)], Ln, l0, l1, tl);
        const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
        const float w10 = (1.0f - tx) * ty, w11 = tx * ty, u0 = 1.0f - tl;
        for (int grp = cq; grp < CA / 4; grp += 4) {
          const float* P = F.aplane[i] + grp * 4;
          const float* Lp = F.aline[i] + grp * 4;
          const float4 a = ldg4(P + ((size_t)y0 * W + x0) * CA), b = ldg4(P + ((size_t)y0 * W + x1) * CA);
          const float4 c = ldg4(P + ((size_t)y1 * W + x0) * CA), d = ldg4(P + ((size_t)y1 * W + x1) * CA);
          const float4 u = ldg4(Lp + (size_t)l0 * CA), v = ldg4(Lp + (size_t)l1 * CA);
          float* xo = XF + sm_ * LDXF + i * CA + grp * 4;
          const bool ok = sm_ < n_valid;
          xo[0] = ok ? (a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11) * (u.x * u0 + v.x * tl) : 0.0f;
          xo[1] = ok ? (a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11) * (u.y * u0 + v.y * tl) : 0.0f;
          xo[2] = ok ? (a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11) * (u.z * u0 + v.z * tl) : 0.0f;
          xo[3] = ok ? (a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11) * (u.w * u0 + v.w * tl) : 0.0f;
        }
      }
      if (cq == 0) {
#pragma unroll
        for (int k = NF; k < LDXF; ++k) XF[sm_ * LDXF + k] = 0.0f;
      }
    }
    __syncthreads();
    // ---- x -> TMEM A operand (hi/lo): half 0 = k 0..63, half 1 = k 64..95 (zeros >= 72) -----------------
    {
      const int k0 = half * 64, nk = half ? 32 : 64;
      float xv[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) xv[j] = (j < nk && k0 + j < K1) ? XF[row * LDXF + k0 + j] : 0.0f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (cc * 32 < nk) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split2(xv[cc * 32 + 2 * j], xv[cc * 32 + 2 * j + 1], hi[j], lo[j]);
          tmem_st16(t_row + (uint32_t)(TC_AHI + k0 / 2 + cc * 16), hi);
          tmem_st16(t_row + (uint32_t)(TC_ALO + k0 / 2 + cc * 16), lo);
        }
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();                                                  // (also: every reader of the x staging is done)
    // ---- L1 (W1B resident) ------------------------------------------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_w, wphase);
      tc_fence_after();
      const uint32_t id = tc_idesc(FC, 0, 0), sbo = (uint32_t)K1_CHUNKS * 128u;
      uint32_t acc = 0;
      for (int ks = 0; ks < K1 / 16; ++ks) {
        const uint64_t bh = umma_desc(wreg_a + ks * 256, 128u, sbo);
        const uint64_t bl = umma_desc(wreg_a + OPER1_BYTES + ks * 256, 128u, sbo);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bh, id, acc);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bl, id, 1u);
        mma_ts_i(tmem + TC_ACC, tmem + TC_ALO + ks * 8, bh, id, 1u);
        acc = 1u;
      }
      umma_commit(bar);
    }
    wphase ^= 1u;
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    if (tid == 0) load_weights(wreg_a, W2_g, W2_BYTES, bar_w);       // W1B has been read: W2 streams in behind epilogue 1
    // ---- epilogue 1: h1 -> TMEM A (hi/lo) + hi/lo images; keep the sign mask -----------------------------
    uint32_t m1[2];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      float v[32];
      tmem_ld32(t_row + (uint32_t)(TC_ACC + c0), v);
      uint32_t hi[16], lo[16], mk = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j] + b1_s[c0 + j], 0.0f);
        mk |= (v[j] > 0.0f ? 1u : 0u) << j;
      }
      m1[cc] = mk;
#pragma unroll
      for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
      tmem_st16(t_row + (uint32_t)(TC_AHI + c0 / 2), hi);
      tmem_st16(t_row + (uint32_t)(TC_ALO + c0 / 2), lo);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) store_chunk(H1H, H1L, row, c0 / 8 + kc, K2_CHUNKS, v + kc * 8);
    }
    tmem_st_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- L2 (W2 resident) -------------------------------------------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_w, wphase);
      tc_fence_after();
      const uint32_t id = tc_idesc(FC, 0, 0), sbo = (uint32_t)K2_CHUNKS * 128u;
      uint32_t acc = 0;
      for (int ks = 0; ks < FC / 16; ++ks) {
        const uint64_t bh = umma_desc(wreg_a + ks * 256, 128u, sbo);
        const uint64_t bl = umma_desc(wreg_a + OPER2_BYTES + ks * 256, 128u, sbo);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bh, id, acc);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bl, id, 1u);
        mma_ts_i(tmem + TC_ACC, tmem + TC_ALO + ks * 8, bh, id, 1u);
        acc = 1u;
      }
      umma_commit(bar);
    }
    wphase ^= 1u;
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    // ---- epilogue 2a: h2 -> hi/lo images (operand of dW3), layer-3 partial sums, sign mask -----------------
    uint32_t m2[2];
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      float v[32];
      tmem_ld32(t_row + (uint32_t)(TC_ACC + c0), v);
      uint32_t mk = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j] + b2_s[c0 + j], 0.0f);
        mk |= (v[j] > 0.0f ? 1u : 0u) << j;
        p0 = fmaf(W3_s[0 * W3_LD + c0 + j], v[j], p0);
        p1 = fmaf(W3_s[1 * W3_LD + c0 + j], v[j], p1);
        p2 = fmaf(W3_s[2 * W3_LD + c0 + j], v[j], p2);
      }
      m2[cc] = mk;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) store_chunk(D2H, D2L, row, c0 / 8 + kc, K2_CHUNKS, v + kc * 8);
    }
    if (half == 1) { part_s[row * 3] = p0; part_s[row * 3 + 1] = p1; part_s[row * 3 + 2] = p2; }
    __syncthreads();
    if (half == 0) {                                                   // sigmoid, dL/d(pre-sigmoid), g.rgb -> dL/dw
      const float ps[3] = {p0 + part_s[row * 3], p1 + part_s[row * 3 + 1], p2 + part_s[row * 3 + 2]};
      float grgb = 0.0f, dp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float s = ps[c] + W3_s[c * W3_LD + FC] * vd_s[row * 3] + W3_s[c * W3_LD + FC + 1] * vd_s[row * 3 + 1] +
                        W3_s[c * W3_LD + FC + 2] * vd_s[row * 3 + 2] + b3_s[c];
        const float rgb = __fdiv_rn(1.0f, 1.0f + expf(-s));
        const float gc = g_s[row * 3 + c];
        dp[c] = row < n_valid ? gc * wk_s[row] * rgb * (1.0f - rgb) : 0.0f;
        dpre_s[row * 3 + c] = dp[c];
        grgb += gc * rgb;
      }
      if (row < n_valid) A.s.gw[(size_t)ray_s[row] * S + kk_s[row]] += grgb;
      // dpre images [s][16] (chunks = 2): columns 0..2 of chunk 0, hi and lo
      uint32_t h01, h2x, l01, l2x;
      split2(dp[0], dp[1], h01, l01); split2(dp[2], 0.0f, h2x, l2x);
      const int off = (((row >> 3) * 2 + 0) * 8 + (row & 7)) * 16;
      *reinterpret_cast<uint2*>(DPH + off) = make_uint2(h01, h2x);
      *reinterpret_cast<uint2*>(DPL + off) = make_uint2(l01, l2x);
      // view-direction columns of dW3 and db3: warp-reduce over the 32 rows, one shared atomic per warp
      float red[12];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int a = 0; a < 3; ++a) red[c * 3 + a] = dp[c] * vd_s[row * 3 + a];
        red[9 + c] = dp[c];
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const float s = warp_sum(red[e]);
        if (lane == 0) atomicAdd(small_s + e, s);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- G3: dW3^T[n][c] += sum_s h2[s][n] dpre[s][c] ------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id = tc_idesc(16, 1, 1);
      for (int ks = 0; ks < TT / 16; ++ks) {
        const uint32_t ao = ks * 2 * K2_CHUNKS * 128, bo = ks * 2 * 2 * 128;
        mma3_ss(tmem + TC_DW3, mn_desc(smem_u32(D2H) + ao, K2_CHUNKS), mn_desc(smem_u32(D2L) + ao, K2_CHUNKS),
                mn_desc(smem_u32(DPH) + bo, 2), mn_desc(smem_u32(DPL) + bo, 2), id, (first && ks == 0) ? 0u : 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    // ---- epilogue 2b: dh2 = (h2 > 0) W3^T dpre -> TMEM A (hi/lo) + hi/lo images --------------------------------
    {
      const float d0 = dpre_s[row * 3], d1 = dpre_s[row * 3 + 1], d2 = dpre_s[row * 3 + 2];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = c0 + j;
          const float gs = W3_s[n] * d0 + W3_s[W3_LD + n] * d1 + W3_s[2 * W3_LD + n] * d2;
          v[j] = ((m2[cc] >> j) & 1u) ? gs : 0.0f;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        tmem_st16(t_row + (uint32_t)(TC_AHI + c0 / 2), hi);
        tmem_st16(t_row + (uint32_t)(TC_ALO + c0 / 2), lo);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) store_chunk(D2H, D2L, row, c0 / 8 + kc, K2_CHUNKS, v + kc * 8);
      }
      tmem_st_wait();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- G2: dh1 = dh2 W2 ; dW2 += dh2^T h1 ; db2 += dh2^T 1 -------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      uint32_t acc = 0;
      const uint32_t id_row = tc_idesc(FC, 0, 1);                      // A from TMEM, B = W2 image read MN-major
      for (int ks = 0; ks < FC / 16; ++ks) {
        const uint64_t bh = mn_desc(wreg_a + ks * 2 * K2_CHUNKS * 128, K2_CHUNKS);
        const uint64_t bl = mn_desc(wreg_a + OPER2_BYTES + ks * 2 * K2_CHUNKS * 128, K2_CHUNKS);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bh, id_row, acc);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bl, id_row, 1u);
        mma_ts_i(tmem + TC_ACC, tmem + TC_ALO + ks * 8, bh, id_row, 1u);
        acc = 1u;
      }
      const uint32_t id_w = tc_idesc(FC, 1, 1), id_b = tc_idesc(16, 1, 1);
      for (int ks = 0; ks < TT / 16; ++ks) {
        const uint32_t go = (first && ks == 0) ? 0u : 1u;
        const uint32_t ao = ks * 2 * K2_CHUNKS * 128;
        const uint64_t ah = mn_desc(smem_u32(D2H) + ao, K2_CHUNKS), al = mn_desc(smem_u32(D2L) + ao, K2_CHUNKS);
        mma3_ss(tmem + TC_DW2, ah, al, mn_desc(smem_u32(H1H) + ao, K2_CHUNKS), mn_desc(smem_u32(H1L) + ao, K2_CHUNKS),
                id_w, go);
        const uint64_t od = mn_desc(smem_u32(ONES) + ks * 2 * 2 * 128, 2);
        mma_ss_i(tmem + TC_DB2, ah, od, id_b, go);
        mma_ss_i(tmem + TC_DB2, al, od, id_b, 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    if (tid == 0) load_weights(wreg_a, W1B_g, W1B_BYTES, bar_w);      // W2 has been read: W1B streams in behind the epilogue
    // ---- epilogue G2: dh1 = (h1 > 0) acc -> TMEM A (hi/lo) + hi/lo images (over the h1 images) ---------------------
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      float v[32];
      tmem_ld32(t_row + (uint32_t)(TC_ACC + c0), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ((m1[cc] >> j) & 1u) ? v[j] : 0.0f;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
      tmem_st16(t_row + (uint32_t)(TC_AHI + c0 / 2), hi);
      tmem_st16(t_row + (uint32_t)(TC_ALO + c0 / 2), lo);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) store_chunk(H1H, H1L, row, c0 / 8 + kc, K2_CHUNKS, v + kc * 8);
    }
    tmem_st_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- G1a: dprod = dh1 W1B ; G1c: db1 += dh1^T 1 ------------------------------------------------------------
    if (tid == 0) {
      mbar_wait(bar_w, wphase);
      tc_fence_after();
      uint32_t acc = 0;
      const uint32_t id_row = tc_idesc(K1, 0, 1);                      // N = 80 (t), B = W1B image read MN-major
      for (int ks = 0; ks < FC / 16; ++ks) {
        const uint64_t bh = mn_desc(wreg_a + ks * 2 * K1_CHUNKS * 128, K1_CHUNKS);
        const uint64_t bl = mn_desc(wreg_a + OPER1_BYTES + ks * 2 * K1_CHUNKS * 128, K1_CHUNKS);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bh, id_row, acc);
        mma_ts_i(tmem + TC_ACC, tmem + TC_AHI + ks * 8, bl, id_row, 1u);
        mma_ts_i(tmem + TC_ACC, tmem + TC_ALO + ks * 8, bh, id_row, 1u);
        acc = 1u;
      }
      const uint32_t id_b = tc_idesc(16, 1, 1);
      for (int ks = 0; ks < TT / 16; ++ks) {
        const uint32_t go = (first && ks == 0) ? 0u : 1u;
        const uint32_t ao = ks * 2 * K2_CHUNKS * 128;
        const uint64_t od = mn_desc(smem_u32(ONES) + ks * 2 * 2 * 128, 2);
        mma_ss_i(tmem + TC_DB1, mn_desc(smem_u32(H1H) + ao, K2_CHUNKS), od, id_b, go);
        mma_ss_i(tmem + TC_DB1, mn_desc(smem_u32(H1L) + ao, K2_CHUNKS), od, id_b, 1u);
      }
      umma_commit(bar);
    }
    wphase ^= 1u;
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    // ---- dprod -> fp32 staging (h2 region: dh2 images are dead); clear the x images (weight region: W1B is dead) ---
    {
      float v[32];
      if (half == 0) {
        tmem_ld32(t_row + (uint32_t)(TC_ACC + 0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) DPF[row * LDXF + j] = v[j];
        tmem_ld32(t_row + (uint32_t)(TC_ACC + 64), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) DPF[row * LDXF + 64 + j] = v[j];
      } else {
        tmem_ld32(t_row + (uint32_t)(TC_ACC + 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) DPF[row * LDXF + 32 + j] = v[j];
      }
      for (int e = tid; e < W1B_BYTES / 16; e += TC_THREADS)          // rows >= n_valid and k 72..79 stay zero
        reinterpret_cast<uint4*>(WREG)[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc_fence_before();
    __syncthreads();
    // ---- products backward (as in bwd_shade_kernel) + the x hi/lo images, two passes of 64 samples -------------------
    for (int pass = 0; pass < 2; ++pass) {
      const int sm_ = pass * 64 + ls;
      float dq0 = 0.0f, dq1 = 0.0f, dq2 = 0.0f;
      if (sm_ < n_valid) {
        const float qq[3] = {q_s[sm_ * 3], q_s[sm_ * 3 + 1], q_s[sm_ * 3 + 2]};
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
          for (int grp = cq; grp < CA / 4; grp += 4) {
            const size_t o00 = ((size_t)y0 * W + x0) * CA + grp * 4, o01 = ((size_t)y0 * W + x1) * CA + grp * 4;
            const size_t o10 = ((size_t)y1 * W + x0) * CA + grp * 4, o11 = ((size_t)y1 * W + x1) * CA + grp * 4;
            const size_t ol0 = (size_t)l0 * CA + grp * 4, ol1 = (size_t)l1 * CA + grp * 4;
            const float4 a4 = ldg4(F.aplane[i] + o00), b4 = ldg4(F.aplane[i] + o01);
            const float4 c4 = ldg4(F.aplane[i] + o10), d4 = ldg4(F.aplane[i] + o11);
            const float4 u4 = ldg4(F.aline[i] + ol0), v4 = ldg4(F.aline[i] + ol1);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const float uv[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
            float dP[4], dL[4], xr[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float gc = DPF[sm_ * LDXF + i * CA + grp * 4 + e];
              const float pi = av[e] * w00 + bv[e] * w01 + cv[e] * w10 + dv[e] * w11;   // the gather's expressions:
              const float li = uv[e] * u0 + vv[e] * tl;                                // bit-identical products
              xr[e] = pi * li;
              dP[e] = gc * li;
              dL[e] = gc * pi;
              gx += dP[e] * ((bv[e] - av[e]) * (1.0f - ty) + (dv[e] - cv[e]) * ty);
              gy += dP[e] * ((cv[e] - av[e]) * (1.0f - tx) + (dv[e] - bv[e]) * tx);
              gl += dL[e] * (vv[e] - uv[e]);
            }
            {  // x images: 4 consecutive k of row sm_ = one 8-byte piece of a 16-byte chunk, hi and lo
              const int k = i * CA + grp * 4;
              uint32_t h0, h1, l0_, l1_;
              split2(xr[0], xr[1], h0, l0_); split2(xr[2], xr[3], h1, l1_);
              const int off = (((sm_ >> 3) * K1_CHUNKS + (k >> 3)) * 8 + (sm_ & 7)) * 16 + (k & 7) * 2;
              *reinterpret_cast<uint2*>(XH + off) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(XL + off) = make_uint2(l0_, l1_);
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
      if (cq == 0 && sm_ < n_valid) {
        float dp0, dp1, dp2;
        contract_bwd(praw_s[sm_ * 3], praw_s[sm_ * 3 + 1], praw_s[sm_ * 3 + 2], dq0 * F.ainv[0],
                     dq1 * F.ainv[1], dq2 * F.ainv[2], dp0, dp1, dp2);
        float* accp = A.s.d_ovd + 6 * (size_t)ray_s[sm_];
        const float z = zk_s[sm_];
        atomicAdd(accp + 0, dp0); atomicAdd(accp + 1, dp1); atomicAdd(accp + 2, dp2);
        atomicAdd(accp + 3, dp0 * z); atomicAdd(accp + 4, dp1 * z); atomicAdd(accp + 5, dp2 * z);
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- G1b: dW1B += dh1^T x (the tile's last product; then the next tile's W1B streams in) ------------------------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id_w = tc_idesc(K1, 1, 1);
      for (int ks = 0; ks < TT / 16; ++ks) {
        const uint32_t ao = ks * 2 * K2_CHUNKS * 128, bo = ks * 2 * K1_CHUNKS * 128;
        mma3_ss(tmem + TC_DW1, mn_desc(smem_u32(H1H) + ao, K2_CHUNKS), mn_desc(smem_u32(H1L) + ao, K2_CHUNKS),
                mn_desc(smem_u32(XH) + bo, K1_CHUNKS), mn_desc(smem_u32(XL) + bo, K1_CHUNKS), id_w,
                (first && ks == 0) ? 0u : 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    tc_fence_after();
    first = false;
    if (tid == 0 && tile + gridDim.x < n_tiles) load_weights(wreg_a, W1B_g, W1B_BYTES, bar_w);
  }

  // ---- flush the TMEM-resident weight gradients (rows = output unit n) ------------------------------------------------
  if (!first) {
    tc_fence_after();
    const int n = row;
    float v[32];
    for (int cc = 0; cc < 2; ++cc) {                                   // dW2[n][k], this thread's column half
      const int c0 = half * 64 + cc * 32;
      tmem_ld32(t_row + (uint32_t)(TC_DW2 + c0), v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (v[j] != 0.0f) atomicAdd(A.d_w2 + n * FC + c0 + j, v[j]);
    }
    if (half == 0) {                                                   // dW1B[n][t], t < 72 (80 columns held)
      for (int c0 = 0; c0 < 96; c0 += 32) {
        tmem_ld32(t_row + (uint32_t)(TC_DW1 + c0), v);                 // reads past column 80 into dW3 / db: ignored
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < NF && v[j] != 0.0f) atomicAdd(A.d_w1b + n * NF + c0 + j, v[j]);
      }
    } else {                                                           // dW3^T[n][c], db2[n], db1[n]: columns 336..383
      tmem_ld32(t_row + (uint32_t)(TC_DW3), v);                        // 336..367: dW3 (16) + db2 (16)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (v[c] != 0.0f) atomicAdd(A.d_w3 + c * (FC + 3) + n, v[c]);
      if (v[16] != 0.0f) atomicAdd(A.d_b2 + n, v[16]);
      tmem_ld32(t_row + (uint32_t)(TC_DB1 - 16), v);                   // 352..383: db2 (16) + db1 (16)
      if (v[16] != 0.0f) atomicAdd(A.d_b1 + n, v[16]);
    }
    if (tid < 9) {
      const int c = tid / 3, a = tid - 3 * c;
      if (small_s[tid] != 0.0f) atomicAdd(A.d_w3 + c * (FC + 3) + FC + a, small_s[tid]);
    } else if (tid < 12) {
      if (small_s[tid] != 0.0f) atomicAdd(A.d_b3 + (tid - 9), small_s[tid]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace lrf
