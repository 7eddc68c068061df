// lrf_render.cu -- the fused per-ray-batch render kernel (sm_100a), streaming / warp-specialised.
//
// One launch = one (field, ray batch): ray generation -> contracted sampling -> VM density gather
// -> softplus/alpha/transmittance scan -> (floater filter) -> selection of the samples whose weight
// exceeds the threshold -> VM appearance gather -> basis+MLP -> composite -> blend / exposure /
// clamp.  Replaces TensorBase.forward (models/tensorBase.py:567-636) and the per-field body of
// LocalTensorfs.forward (local_tensorfs.py:397-497).
//
// One persistent CTA per SM, 16 warps (512 threads; LRF_THREADS) in three roles that only meet through
// mbarriers and a few shared-memory counters:
//
//   producers (warps 0 .. W_ISSUE-1: 11)  each takes one ray at a time from a global counter (dynamic
//       scheduling: no wave quantisation of a 4096-ray batch).  March: lane = sample, 32 samples
//       per step, transmittance by a warp-shuffle product scan with a carried prefix.  Samples
//       with w > threshold go to a small per-warp queue whose entries carry their ray slot (so samples of
//       consecutive rays share a gather); every 32 of them are gathered (lane = sample: 3 planes x 4 texels
//       x 96 B + 3 lines, 72 products), split into bf16 hi/lo and written as one ROW of the current 128-row
//       A tile in shared memory (K-major 8x8 core matrices).  Rows are handed out by an atomic cursor, so
//       rays of all producers interleave in a tile; each row carries (ray slot, weight).  Every row arrives
//       once on the tile's "full" mbarrier (128 arrivals = tile ready).
//   MMA issuer (warp W_ISSUE, one elected lane)  per tile: layer 1 (A from shared memory, basis folded
//       into W1, K = 80) and layer 2 (A from TMEM, K = 128) as tcgen05.mma kind::f16 M=128 N=128,
//       three bf16 products hi*hi + hi*lo + lo*hi per K-step, fp32 accumulators in TMEM;
//       tcgen05.commit signals the consumers and frees the A tile for the producers.
//   consumers (the last 4 warps, thread = tile row)  epilogue 1: tcgen05.ld acc1, bias + ReLU, re-split,
//       tcgen05.st as layer 2's A operand; epilogue 2: tcgen05.ld acc2, bias + ReLU, layer 3
//       (131 -> 3) + sigmoid; composite: w * rgb is added to the row's ray slot as 32-bit FIXED
//       POINT with shared-memory atomics (integer adds are associative, so results are
//       bit-identical whatever the interleaving of rows, tiles and threads); the last contributor
//       of a ray writes its output (white background, blend, accumulate, exposure, clamp) and, in a
//       multi-GPU run, stores the finished pixel into every peer's gathered buffer.
//
// render_kernel_t<true> (fields with positional encodings) inserts a layer 0 (basis_mat as its own product),
// builds the encoded layer-1 input in TMEM 64 columns at a time and streams layer 1's weight chunks by TMA.
// Compile-time experiments kept as switches (DESIGN.md par. 10): LRF_THREADS, LRF_CONS_WARPS, LRF_SPIN_NS,
// LRF_STEAL.  The weight operands are staged once per CTA by TMA bulk copies (cp.async.bulk -> UBLKCP).
#include <cstdlib>

#include "lrf_device.cuh"

namespace lrf {

#ifndef LRF_THREADS
#define LRF_THREADS 512             // 16 warps (128 registers each); 640 = 20 warps at 96 registers
#endif
#ifndef LRF_CONS_WARPS
#define LRF_CONS_WARPS 4            // 4: one consumer warp per TMEM lane quarter; 8: two per quarter (64 columns each)
#endif
constexpr int THREADS = LRF_THREADS;
constexpr int N_CONS = LRF_CONS_WARPS;
constexpr int W_CONS = THREADS / 32 - N_CONS;   // the last warps: consumers (warp % 4 = TMEM lane quarter)
constexpr int W_ISSUE = W_CONS - 1;             // the warp before them: MMA issuer
constexpr int MAX_PROD = W_ISSUE;               // producer warps 0 .. W_ISSUE-1 (11 at 512 threads / 4 consumers)
constexpr int CONS_COLS = FC * 4 / N_CONS;      // columns of the accumulators one consumer thread handles
static_assert(THREADS % 128 == 0 && W_ISSUE >= 1 && (N_CONS == 4 || N_CONS == 8), "warp roles");
constexpr int NSLOT = 3;            // ray slots per producer warp
constexpr int QCAP = 64;            // per-warp queue of selected samples
constexpr int SPIN_PAD = 2048;      // polls before a blocked producer pads the open tile

// colour accumulators are 32-bit fixed point (2^-30 units, 9.3e-10: finer than an fp32 ulp of the
// sums, which are <= 1): integer adds are associative, so the rows of a ray can be composited in any
// order, by any thread, and the result is bit-identical.  (32-bit shared-memory atomics are native;
// 64-bit ones turn into contended CAS loops.)
constexpr float FIX_SCALE = 1073741824.0f;              // 2^30
constexpr float FIX_INV = 1.0f / 1073741824.0f;

struct Slot {                       // one ray in flight between a producer and the consumers
  float o[3];                       // ray origin / |d| (the gathers of queued samples read the ray from its slot)
  float nrm;
  float vd[3];
  float blend;
  unsigned int rgb[3];              // sum of w * rgb, fixed point
  float acc;
  float depth;                      // final depth of the ray (for the fused pixel exchange)
  float vterm[3];                   // (fields with view encodings) W3[:, 128:] . [vd, PE(vd)]
  long long ray;
  int pending;                      // 1 (open token) + rows submitted and not yet composited
  int state;                        // 0 free, 1 active
};

#ifndef LRF_STEAL
#define LRF_STEAL 0                 // (measured: correct but 10 % slower) idle producers help the in-flight rays of their CTA, one 32-sample step at a time
#endif
constexpr int MAXST = 32;           // steps a ray may have for step stealing (S <= 1024); longer rays march sequentially

// One ray being marched by its owner producer and, once the global ray queue is empty, by the idle
// producers of the CTA: steps are claimed from `next_step`; step s publishes the product P[s] of its
// 32 transmittance factors right after its density gathers, then waits for P[0..s-1] and multiplies
// them IN ORDER (bit-identical to the sequential carry), so every weight is independent of who
// computed which step; per-step partial sums are added in step order by the owner.
struct RayRec {
  unsigned int next_step;           // claim counter (>= n_steps between rays: nothing to claim)
  unsigned int stop_at;             // steps >= stop_at contribute nothing (early ray termination, T < 1e-10)
  unsigned int ready, done;         // bit s: P[s] published / step s finished (sums stored, selected samples counted)
  int slot_id;
  long long ray;
  float P[MAXST], accs[MAXST], deps[MAXST];
  int marched[MAXST];
};

struct Ctrl {                       // CTA control block in shared memory
  unsigned long long bar_w, full[2], mma1, a2rdy, mma2;
  unsigned long long mma0, xrdy[2], xfree[2], wrdy;   // fields with positional encodings (PE pipeline)
  uint32_t tmem;
  unsigned int released;            // tiles whose A1 buffer + row metadata may be overwritten
  unsigned int cursor;              // rows handed out so far (tile = cursor >> 7)
  unsigned int prod_done;
  unsigned int draw_done;           // producers whose draw from the global ray queue has failed (they help from then on)
  unsigned int n_tiles_final;
  unsigned int done;
};

struct SmemV3 {
  int prep, a1, mslot, mw, part, slots, recs, q, alpha, z, ctrl, total;
  int per_prod;                     // bytes of one producer's queue (+ alpha table)
};

constexpr int PE_WBUF = (PE_RESIDENT + 1023) & ~1023;     // offset of the streamed layer-1 chunk inside the prep region
__host__ __device__ inline SmemV3 smem_v3(int S, bool floater, int nprod, bool pe = false) {
  SmemV3 L;
  int off = 0;
  const int Sp = (S + 3) & ~3;
  L.prep = off;   off += pe ? PE_WBUF + 2 * PE_WC_BYTES : PREP_BYTES;
  off = (off + 1023) & ~1023;
  L.a1 = off;     off += 2 * 2 * OPER1_BYTES;          // two A1 tiles (hi + lo each)
  L.mslot = off;  off += 2 * TM;                        // per-row slot id (u8), two tiles
  L.mw = off;     off += 2 * TM * 4;                    // per-row weight
  L.part = off;   off += (N_CONS == 8 ? TM * 4 * 4 : 0); // layer-3 partial sums of the upper column half
  off = (off + 15) & ~15;
  L.slots = off;  off += MAX_PROD * NSLOT * (int)sizeof(Slot);
  off = (off + 15) & ~15;
  L.recs = off;   off += (LRF_STEAL ? nprod * (int)sizeof(RayRec) : 0);
  L.q = off;
  L.per_prod = QCAP * 2 + QCAP * 4 + QCAP + (floater ? Sp * 4 : 0);   // sample index, weight, slot id per entry
  L.per_prod = (L.per_prod + 15) & ~15;
  off += nprod * L.per_prod;
  L.alpha = QCAP * 7 + (16 - (QCAP * 7) % 16) % 16;     // offset of the alpha table inside a producer block
  L.z = off;      off += (Sp + 4) * 4;
  off = (off + 15) & ~15;
  L.ctrl = off;   off += (int)sizeof(Ctrl);
  L.total = (off + 15) & ~15;
  return L;
}

// ---- outputs (local_tensorfs.py:467-497) --------------------------------------------------------
__device__ __forceinline__ void write_rgb(const BatchDev& B, long long r,
                                          const unsigned int* rgb_fix, float acc, float blend,
                                          float depth) {
  float c[3];
  const float bg = B.white_bg ? (1.0f - acc) : 0.0f;              // tensorBase.py:633-634
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const unsigned int v = *reinterpret_cast<const volatile unsigned int*>(rgb_fix + a);
    c[a] = (__uint2float_rn(v) * FIX_INV + bg) * blend;           // tensorBase.py:632
  }
  if (B.accumulate) {
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = B.rgb[B.rgb_stride * r + a] + c[a];
  }
  if (B.finalize) {
    if (B.exposure) {
      const long long view = B.rays_per_view > 0 ? r / B.rays_per_view : 0;
      const float* E = B.exposure + 9 * view;
      const float o0 = E[0] * c[0] + E[1] * c[1] + E[2] * c[2];
      const float o1 = E[3] * c[0] + E[4] * c[1] + E[5] * c[2];
      const float o2 = E[6] * c[0] + E[7] * c[1] + E[8] * c[2];
      c[0] = o0; c[1] = o1; c[2] = o2;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = fminf(1.0f, fmaxf(0.0f, c[a]));
  }
  float* o = B.rgb + B.rgb_stride * r;
  o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
  if (B.n_peers > 0 && B.finalize) {
    // fused pixel exchange: this ray's (r,g,b,depth) goes straight into every peer's gathered buffer
    if (B.mc_pix) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(B.mc_pix + 4 * r),
                   "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(depth)
                   : "memory");
    } else {
      const float4 v = make_float4(c[0], c[1], c[2], depth);
      for (int p = 0; p < B.n_peers; ++p) *reinterpret_cast<float4*>(B.peer_pix[p] + 4 * r) = v;
    }
  }
}

__device__ __forceinline__ void finalize_slot(const BatchDev& B, Slot* s) {
  write_rgb(B, s->ray, s->rgb, s->acc, s->blend, s->depth);
  __threadfence_block();
  *reinterpret_cast<volatile int*>(&s->state) = 0;
}

// 72 appearance products of one sample -> one row of the A1 tile (bf16 hi/lo, 9 chunks of 8)
__device__ __forceinline__ void app_row(const FieldDev& F, const float* q, unsigned char* a_hi,
                                        unsigned char* a_lo, int row) {
LRF_PLANE_LOOP
  for (int i = 0; i < 3; ++i) {
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
    const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
    const float w10 = (1.0f - tx) * ty, w11 = tx * ty;
    const float u0 = 1.0f - tl;
#pragma unroll
    for (int c8 = 0; c8 < CA / 8; ++c8) {
      float v[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = c8 * 8 + h * 4;
        const float4 a = ldg4(p00 + o), b = ldg4(p01 + o), c = ldg4(p10 + o), d = ldg4(p11 + o);
        const float4 u = ldg4(q0 + o), w = ldg4(q1 + o);
        v[4 * h + 0] = (a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11) * (u.x * u0 + w.x * tl);
        v[4 * h + 1] = (a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11) * (u.y * u0 + w.y * tl);
        v[4 * h + 2] = (a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11) * (u.z * u0 + w.z * tl);
        v[4 * h + 3] = (a.w * w00 + b.w * w01 + c.w * w10 + d.w * w11) * (u.w * u0 + w.w * tl);
      }
      store_chunk(a_hi, a_lo, row, i * (CA / 8) + c8, K1_CHUNKS, v);
    }
  }
}

// the same row from bf16 texels (LrfField.grid_dtype = LRF_GRID_BF16): three 16-byte loads per texel
__device__ __forceinline__ void app_row_bf16(const FieldDev& F, const float* q, unsigned char* a_hi,
                                             unsigned char* a_lo, int row) {
LRF_PLANE_LOOP
  for (int i = 0; i < 3; ++i) {
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
    const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty);
    const float w10 = (1.0f - tx) * ty, w11 = tx * ty;
    const float u0 = 1.0f - tl;
#pragma unroll
    for (int c8 = 0; c8 < CA / 8; ++c8) {
      const Tex8 a = ldg8_bf16(p00 + 8 * c8), b = ldg8_bf16(p01 + 8 * c8), c = ldg8_bf16(p10 + 8 * c8),
                 d = ldg8_bf16(p11 + 8 * c8);
      const Tex8 u = ldg8_bf16(q0 + 8 * c8), w = ldg8_bf16(q1 + 8 * c8);
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        v[e] = (a.v[e] * w00 + b.v[e] * w01 + c.v[e] * w10 + d.v[e] * w11) * (u.v[e] * u0 + w.v[e] * tl);
      store_chunk(a_hi, a_lo, row, i * (CA / 8) + c8, K1_CHUNKS, v);
    }
  }
}

struct ProdCtx {
  unsigned char* smem;
  const SmemV3* L;
  Ctrl* ctrl;
  int lane;
};

// wait until buffer (T & 1) may be rewritten for tile T: tile T-2 has been read by the tensor core
// and its row metadata by the consumers.  A monotonic counter, not an mbarrier parity: producers
// can be more than two tiles ahead of the consumers, which a 1-bit phase cannot express.
__device__ __forceinline__ void wait_tile_free(Ctrl* ctrl, unsigned int T) {
  if (T >= 2) {
#ifndef LRF_SPIN_NS
#define LRF_SPIN_NS 200
#endif
    while (*reinterpret_cast<volatile unsigned int*>(&ctrl->released) + 1u < T) {
      if (LRF_SPIN_NS) __nanosleep(LRF_SPIN_NS);   // back off: a spinning producer takes issue slots from the gathers
    }
    __threadfence_block();
  }
}

// completes the currently open tile with invalid rows (slot 0xFF) so the consumers can drain it
__device__ __noinline__ void pad_open_tile(const ProdCtx& P) {
  unsigned int start = 0, n = 0;
  if (P.lane == 0) {
    unsigned int old = *reinterpret_cast<volatile unsigned int*>(&P.ctrl->cursor);
    for (;;) {
      const unsigned int rem = old & (TM - 1);
      if (rem == 0) { n = 0; break; }
      const unsigned int seen = atomicCAS(&P.ctrl->cursor, old, old + (TM - rem));
      if (seen == old) { start = old; n = TM - rem; break; }
      old = seen;
    }
  }
  start = __shfl_sync(0xffffffffu, start, 0);
  n = __shfl_sync(0xffffffffu, n, 0);
  if (n == 0) return;
  const unsigned int T = start >> 7;
  wait_tile_free(P.ctrl, T);
  unsigned char* mslot = P.smem + P.L->mslot + (T & 1) * TM;
  float* mw = reinterpret_cast<float*>(P.smem + P.L->mw) + (T & 1) * TM;
  const uint32_t full = smem_u32(&P.ctrl->full[T & 1]);
  for (unsigned int g = start + P.lane; g < start + n; g += 32) {
    mslot[g & (TM - 1)] = 0xFF;
    mw[g & (TM - 1)] = 0.0f;
    __threadfence_block();
    mbar_arrive(full);
  }
  __syncwarp();
}

// per-warp queue of selected samples: (sample index, weight, ray slot); entries of different rays mix
struct WarpQ {
  unsigned short* qk;
  float* qw;
  unsigned char* qs;
  int qn;
  unsigned long long n_app;
};

// gathers the first n (<= 32) queued samples (each with its own ray, read from its slot) and submits
// them as tile rows.  Their slots' `pending` counts were raised when the samples were selected.
template <bool H16>
__device__ __forceinline__ void flush_rows(const ProdCtx& P, const FieldDev& F, const float* z_s,
                                           const Slot* slots, const WarpQ& Q, int n) {
  unsigned int start = 0;
  if (P.lane == 0) start = atomicAdd(&P.ctrl->cursor, (unsigned int)n);
  start = __shfl_sync(0xffffffffu, start, 0);
  if (P.lane < n) {
    const unsigned int g = start + P.lane, T = g >> 7;
    const int row = (int)(g & (TM - 1)), b = (int)(T & 1);
    const int k = Q.qk[P.lane], sid = Q.qs[P.lane];
    const Slot* sl = slots + sid;
    RaySm R;
    R.o[0] = sl->o[0]; R.o[1] = sl->o[1]; R.o[2] = sl->o[2];
    R.vd[0] = sl->vd[0]; R.vd[1] = sl->vd[1]; R.vd[2] = sl->vd[2];
    float p[3], q[3];
    sample_pos(F, R, z_s[k], p, q);
    wait_tile_free(P.ctrl, T);
    unsigned char* a_hi = P.smem + P.L->a1 + b * (2 * OPER1_BYTES);
    if constexpr (H16) app_row_bf16(F, q, a_hi, a_hi + OPER1_BYTES, row);
    else app_row(F, q, a_hi, a_hi + OPER1_BYTES, row);
    P.smem[P.L->mslot + b * TM + row] = (unsigned char)sid;
    reinterpret_cast<float*>(P.smem + P.L->mw)[b * TM + row] = Q.qw[P.lane];
    fence_async_smem();                  // generic-proxy writes -> visible to the tensor core
    mbar_arrive(smem_u32(&P.ctrl->full[b]));
  }
  __syncwarp();
}

// appends the lanes with on == true (mask m = their ballot) to the warp's queue; flushes 32 rows when it
// holds >= 32.  The caller has already added popc(m) to the slot's pending count.
template <bool H16>
__device__ __forceinline__ void queue_push(const ProdCtx& P, const FieldDev& F, const float* z_s,
                                           const Slot* slots, WarpQ& Q, unsigned m, bool on, int k,
                                           float wgt, int slot_id) {
  const int lane = P.lane;
  if (on) {
    const int pos = Q.qn + __popc(m & ((1u << lane) - 1u));
    Q.qk[pos] = (unsigned short)k; Q.qw[pos] = wgt; Q.qs[pos] = (unsigned char)slot_id;
  }
  Q.qn += __popc(m);
  Q.n_app += __popc(m);
  __syncwarp();
  if (Q.qn >= 32) {
    flush_rows<H16>(P, F, z_s, slots, Q, 32);
    unsigned short tk = 0; float tw = 0.0f; unsigned char ts = 0;
    if (lane < Q.qn - 32) { tk = Q.qk[32 + lane]; tw = Q.qw[32 + lane]; ts = Q.qs[32 + lane]; }
    __syncwarp();
    if (lane < Q.qn - 32) { Q.qk[lane] = tk; Q.qw[lane] = tw; Q.qs[lane] = ts; }
    Q.qn -= 32;
    __syncwarp();
  }
}

// alpha of sample k of the ray in slot-resident form (tensorBase.py:581-610); counts valid samples
template <bool H16>
__device__ __forceinline__ float sample_alpha(const FieldDev& F, const RaySm& R, const float* z_s, int k,
                                              int S, int& marched) {
  const float z = z_s[k];
  float p[3], q[3];
  sample_pos(F, R, z, p, q);
  bool valid = (k != S - 1);                                  // ray_valid[:, -1] = 0
  if (valid && F.alpha_vol) valid = alpha_mask(F, p) > 0.0f;  // tensorBase.py:593-598
  float sigma = 0.0f;
  if (valid) {
    sigma = feature2density(density_feature_t<H16>(F, q), F.density_shift, F.act);
    ++marched;
  }
  const float dist = z_s[k + 1] - z;
  float alpha = 1.0f - expf(-sigma * dist * F.distance_scale);   // tensorBase.py:610, same form
  if (k == S - 1) alpha = 1.0f;                                  // alpha[:, -1] = 1
  return alpha;
}

// One claimed step (32 samples) of the ray described by `rec`: see RayRec.
template <bool H16>
__device__ __forceinline__ void process_step(const ProdCtx& P, const FieldDev& F, const BatchDev& B,
                                             const float* z_s, Slot* slots, RayRec* rec, int s, WarpQ& Q) {
  const int lane = P.lane, S = F.S, k = s * 32 + lane;
  const int slot_id = *reinterpret_cast<volatile int*>(&rec->slot_id);
  const long long ray = *reinterpret_cast<volatile long long*>(&rec->ray);
  Slot* slot = slots + slot_id;
  RaySm R;
  R.o[0] = slot->o[0]; R.o[1] = slot->o[1]; R.o[2] = slot->o[2];
  R.vd[0] = slot->vd[0]; R.vd[1] = slot->vd[1]; R.vd[2] = slot->vd[2];
  const bool skip = (unsigned)s >= *reinterpret_cast<volatile unsigned int*>(&rec->stop_at);
  float alpha = 0.0f;
  int marched = 0;
  if (!skip && k < S) alpha = sample_alpha<H16>(F, R, z_s, k, S, marched);
  const float f = (k < S) ? (1.0f - alpha) + 1e-10f : 1.0f;
  const float inc = warp_scan_mul(f, lane);
  float exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) exc = 1.0f;
  const float Ploc = __shfl_sync(0xffffffffu, inc, 31);
  if (lane == 0) {
    *reinterpret_cast<volatile float*>(&rec->P[s]) = skip ? 0.0f : Ploc;
    __threadfence_block();
    atomicOr(&rec->ready, 1u << s);
  }
  // carry = ((1 * P[0]) * P[1]) ... * P[s-1], the sequential order
  float carry = 1.0f;
  if (!skip) {
    const unsigned need = s ? (0xffffffffu >> (32 - s)) : 0u;
    while ((*reinterpret_cast<volatile unsigned int*>(&rec->ready) & need) != need) { }
    __threadfence_block();
    for (int j = 0; j < s; ++j) carry *= *reinterpret_cast<volatile float*>(&rec->P[j]);
  }
  const bool live = !skip && !(carry < T_EPS);                  // a terminated ray's later steps contribute nothing
  const float wgt = live ? alpha * (carry * exc) : 0.0f;
  const float zk = (k < S) ? z_s[k] : 0.0f;
  const float acc_s = warp_sum((live && k < S) ? wgt : 0.0f);
  const float dep_s = warp_sum((live && k < S) ? wgt * zk : 0.0f);
  const int marched_s = (int)warp_sum(live ? (float)marched : 0.0f);
  if (B.weights && k < S) B.weights[(size_t)ray * S + k] = wgt;
  if (live && carry * Ploc < T_EPS && lane == 0) atomicMin(&rec->stop_at, (unsigned)(s + 1));
  const bool on = live && (k < S) && (wgt > F.weight_thres);    // tensorBase.py:622
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (lane == 0) {
    rec->accs[s] = acc_s; rec->deps[s] = dep_s; rec->marched[s] = marched_s;
    if (m) atomicAdd(&slot->pending, __popc(m));                // before `done`: the owner drops its token after it
    __threadfence_block();
    atomicOr(&rec->done, 1u << s);
  }
  queue_push<H16>(P, F, z_s, slots, Q, m, on, k, wgt, slot_id);
}

// =================================================================================================
// PE = the field has positional encodings (fea_pe > 0 or view_pe > 0): basis_mat is its own tensor-core
// product (layer 0), the consumers build the encoded layer-1 input chunk by chunk in TMEM, layer 1's weight
// operand is streamed by TMA in K-chunks of 64.  PE = false is the reference-default fast path (folded basis).
// H16 = the grids are stored as bfloat16 (LrfField.grid_dtype): the gathers of the producers read 16-byte
// texel pieces of 8 components; everything downstream of the gathers is the same code.
template <bool PE, bool H16 = false>
__global__ void __launch_bounds__(THREADS, 1)
render_kernel_t(const FieldDev F, const BatchDev B, const int nprod) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const bool floater = B.floater_thresh > 0.0f;
  const SmemV3 L = smem_v3(F.S, floater, nprod, PE);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + L.ctrl);
  Slot* slots = reinterpret_cast<Slot*>(smem + L.slots);
  float* z_s = reinterpret_cast<float*>(smem + L.z);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = F.S;

  // ---- prologue ---------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(smem_u32(&ctrl->bar_w), 1);
    mbar_init(smem_u32(&ctrl->full[0]), TM);
    mbar_init(smem_u32(&ctrl->full[1]), TM);
    mbar_init(smem_u32(&ctrl->mma1), 1);
    mbar_init(smem_u32(&ctrl->a2rdy), TM * N_CONS / 4);
    mbar_init(smem_u32(&ctrl->mma2), 1);
    if (PE) {
      mbar_init(smem_u32(&ctrl->mma0), 1);
      mbar_init(smem_u32(&ctrl->xrdy[0]), TM); mbar_init(smem_u32(&ctrl->xrdy[1]), TM);
      mbar_init(smem_u32(&ctrl->xfree[0]), 1); mbar_init(smem_u32(&ctrl->xfree[1]), 1);
      mbar_init(smem_u32(&ctrl->wrdy), 1);
    }
    ctrl->cursor = 0; ctrl->prod_done = 0; ctrl->draw_done = 0; ctrl->n_tiles_final = 0; ctrl->done = 0;
    ctrl->released = 0;
    constexpr uint32_t bytes = PE ? PE_RESIDENT : PREP_BYTES;
    mbar_expect_tx(smem_u32(&ctrl->bar_w), bytes);
    constexpr uint32_t CH = 32768;
    for (uint32_t o = 0; o < bytes; o += CH)
      tma_bulk_g2s(smem_u32(smem + L.prep) + o, reinterpret_cast<const char*>(F.prep) + o,
                   min(CH, bytes - o), smem_u32(&ctrl->bar_w));
  }
  for (int k = tid; k < S; k += THREADS) z_s[k] = F.z[k];
  if (tid == 0) z_s[S] = F.z[S - 1];       // dist of the last sample = 0 (tensorBase.py:584-587)
  for (int e = tid; e < MAX_PROD * NSLOT * (int)sizeof(Slot) / 4; e += THREADS)
    reinterpret_cast<int*>(slots)[e] = 0;
  if (LRF_STEAL && tid < nprod)                                      // no ray open: nothing to claim
    reinterpret_cast<RayRec*>(smem + L.recs)[tid].next_step = 0x40000000u;
  {  // K padding 72..79 of both A1 tiles (hi and lo) is zero for the whole launch
    const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = tid; e < 2 * TM; e += THREADS) {
      unsigned char* a_hi = smem + L.a1 + (e / TM) * (2 * OPER1_BYTES);
      store_chunk(a_hi, a_hi + OPER1_BYTES, e % TM, K1_CHUNKS - 1, K1_CHUNKS, zero);
    }
  }
  if (warp == W_CONS) tmem_alloc(smem_u32(&ctrl->tmem), TMEM_COLS);
  if (B.signal_seq && B.wait_seq && tid < B.n_peers) {
    // fused pixel exchange: the peers must be done with the gathered buffer this launch overwrites
    const unsigned long long* local = B.peer_flags[B.rank] + tid;
    unsigned long long v = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local) : "memory");
    } while (v < B.wait_seq);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctrl->tmem);

  const int prod_id = warp < W_ISSUE ? warp : -1;

  if (prod_id >= 0 && prod_id < nprod) {
    // ======================================= PRODUCER ============================================
    ProdCtx P{smem, &L, ctrl, lane};
    unsigned char* mine = smem + L.q + prod_id * L.per_prod;
    WarpQ Q;
    Q.qk = reinterpret_cast<unsigned short*>(mine);
    Q.qw = reinterpret_cast<float*>(mine + QCAP * 2);
    Q.qs = mine + QCAP * 6;
    Q.qn = 0; Q.n_app = 0;
    float* alpha_s = reinterpret_cast<float*>(mine + L.alpha);
    RayRec* recs = reinterpret_cast<RayRec*>(smem + L.recs);
    RayRec* rec = recs + prod_id;
    const int n_steps = (S + 31) >> 5;
    const bool stepwise = LRF_STEAL && !floater && n_steps <= MAXST;   // else: the owner marches its ray alone
    const unsigned full_mask = n_steps >= 32 ? 0xffffffffu : ((1u << n_steps) - 1u);
    unsigned long long n_march = 0;
    int next_slot = 0;
    for (;;) {
      long long ray = 0;
      if (lane == 0) ray = (long long)atomicAdd(B.sched, 1ull);
      ray = __shfl_sync(0xffffffffu, ray, 0);
      if (ray >= B.n_rays) break;
      RaySm R;
      setup_ray(B, ray, R);
      // -- acquire a ray slot (if the consumers starve: submit what this warp holds, pad the open tile) --
      const int slot_id = prod_id * NSLOT + next_slot;
      next_slot = (next_slot + 1 == NSLOT) ? 0 : next_slot + 1;
      Slot* slot = slots + slot_id;
      for (int spins = 0;; ++spins) {
        int st = 0;
        if (lane == 0) st = *reinterpret_cast<volatile int*>(&slot->state);
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st == 0) break;
        if (spins == SPIN_PAD) {
          if (Q.qn > 0) { flush_rows<H16>(P, F, z_s, slots, Q, Q.qn); Q.qn = 0; }
          pad_open_tile(P);
          spins = 0;
        }
      }
      if (lane == 0) {
        slot->o[0] = R.o[0]; slot->o[1] = R.o[1]; slot->o[2] = R.o[2]; slot->nrm = R.nrm;
        slot->vd[0] = R.vd[0]; slot->vd[1] = R.vd[1]; slot->vd[2] = R.vd[2];
        slot->blend = R.blend;
        slot->rgb[0] = slot->rgb[1] = slot->rgb[2] = 0u;
        slot->ray = ray;
        slot->pending = 1;
      }
      if (PE) {
        // view term of layer 3 (tensorBase.py:126-131): W3[:, 128:] . [vd, sin(vd 2^f), cos(vd 2^f)], d-major f-minor
        const int V = F.view_pe, nv = 3 * (1 + 2 * V), ld = FC + nv;
        float vt[3] = {0.0f, 0.0f, 0.0f};
        for (int j = lane; j < nv; j += 32) {
          float v;
          if (j < 3) v = R.vd[j];
          else {
            const int e = (j - 3) % (3 * V), d = e / V, f = e - d * V;
            const float ang = R.vd[d] * (float)(1 << f);
            v = (j - 3) < 3 * V ? sinf(ang) : cosf(ang);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) vt[c] = fmaf(__ldg(F.w3 + c * ld + FC + j), v, vt[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) vt[c] = warp_sum(vt[c]);
        if (lane == 0) { slot->vterm[0] = vt[0]; slot->vterm[1] = vt[1]; slot->vterm[2] = vt[2]; }
      }
      if (lane == 0) {
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&slot->state) = 1;
      }
      __syncwarp();
      float acc = 0.0f, dep = 0.0f;
      if (stepwise) {
        // -- march by claimed steps (tensorBase.py:581-615); idle producers of this CTA may take steps --
        if (lane == 0) {
          rec->slot_id = slot_id; rec->ray = ray;
          rec->stop_at = (unsigned)n_steps; rec->ready = 0u; rec->done = 0u;
          __threadfence_block();
          *reinterpret_cast<volatile unsigned int*>(&rec->next_step) = 0u;     // opens the ray for claims
        }
        __syncwarp();
        for (;;) {
          unsigned int st = 0;
          if (lane == 0) st = atomicAdd(&rec->next_step, 1u);
          st = __shfl_sync(0xffffffffu, st, 0);
          if (st >= (unsigned)n_steps) break;
          if (st >= *reinterpret_cast<volatile unsigned int*>(&rec->stop_at)) {
            // the ray has terminated (T < 1e-10): close it -- this step and every unclaimed one contribute
            // nothing (steps a helper has already claimed past this point skip themselves)
            unsigned int rem = 0;
            if (lane == 0) rem = atomicExch(&rec->next_step, 0x40000000u);
            rem = min(__shfl_sync(0xffffffffu, rem, 0), (unsigned)n_steps);
            unsigned int mask = 1u << st;
            for (unsigned int j = rem; j < (unsigned)n_steps; ++j) mask |= 1u << j;
            if (lane < n_steps && ((mask >> lane) & 1u)) {
              rec->accs[lane] = 0.0f; rec->deps[lane] = 0.0f; rec->marched[lane] = 0;
              *reinterpret_cast<volatile float*>(&rec->P[lane]) = 0.0f;
            }
            if (B.weights)
              for (int k = (int)st * 32 + lane; k < S; k += 32)
                if ((mask >> (k >> 5)) & 1u) B.weights[(size_t)ray * S + k] = 0.0f;
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicOr(&rec->ready, mask); atomicOr(&rec->done, mask); }
            break;
          }
          process_step<H16>(P, F, B, z_s, slots, rec, (int)st, Q);
        }
        while (*reinterpret_cast<volatile unsigned int*>(&rec->done) != full_mask) { }   // helpers' steps
        __threadfence_block();
        int marched = 0;
        for (int j = 0; j < n_steps; ++j) { acc += rec->accs[j]; dep += rec->deps[j]; marched += rec->marched[j]; }
        n_march += (unsigned long long)marched;
      } else {
        // -- sequential march of the whole ray by its owner (floater filter, very long sample tables) ----
        float carry = 1.0f, acc_p = 0.0f, dep_p = 0.0f, idx_p = 0.0f;
        int k0 = 0, marched = 0;
        float* wo = B.weights ? B.weights + (size_t)ray * S : nullptr;
        for (; k0 < S; k0 += 32) {
          const int k = k0 + lane;
          float alpha = 0.0f;
          if (k < S) alpha = sample_alpha<H16>(F, R, z_s, k, S, marched);
          const float f = (k < S) ? (1.0f - alpha) + 1e-10f : 1.0f;
          const float inc = warp_scan_mul(f, lane);
          float exc = __shfl_up_sync(0xffffffffu, inc, 1);
          if (lane == 0) exc = 1.0f;
          const float wgt = alpha * (carry * exc);
          if (k < S) {
            acc_p += wgt;
            dep_p += wgt * z_s[k];
            idx_p += wgt * (float)k;
            if (floater) alpha_s[k] = alpha;
            else if (wo) wo[k] = wgt;
          }
          carry *= __shfl_sync(0xffffffffu, inc, 31);
          if (!floater) {
            const bool on = (k < S) && (wgt > F.weight_thres);          // tensorBase.py:622
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (lane == 0 && m) atomicAdd(&slot->pending, __popc(m));
            queue_push<H16>(P, F, z_s, slots, Q, m, on, k, wgt, slot_id);
            if (carry < T_EPS) { k0 += 32; break; }                      // early ray termination
          }
        }
        if (!floater && wo)
          for (int k = k0 + lane; k < S; k += 32) wo[k] = 0.0f;          // terminated tail
        acc = warp_sum(acc_p); dep = warp_sum(dep_p);
        if (floater) {                                                   // tensorBase.py:617-620
          const float lim = warp_sum(idx_p) * B.floater_thresh;
          __syncwarp();
          float c2 = 1.0f;
          for (int kb = 0; kb < S; kb += 32) {
            const int k = kb + lane;
            float a = (k < S) ? alpha_s[k] : 0.0f;
            if ((float)k < lim) a = 0.0f;
            if (k == S - 1) a = 1.0f;
            const float f = (k < S) ? (1.0f - a) + 1e-10f : 1.0f;
            const float inc = warp_scan_mul(f, lane);
            float exc = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) exc = 1.0f;
            const float wgt = a * (c2 * exc);
            c2 *= __shfl_sync(0xffffffffu, inc, 31);
            if (k < S && wo) wo[k] = wgt;
            const bool on = (k < S) && (wgt > F.weight_thres);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (lane == 0 && m) atomicAdd(&slot->pending, __popc(m));
            queue_push<H16>(P, F, z_s, slots, Q, m, on, k, wgt, slot_id);
          }
        }
        n_march += (unsigned long long)warp_sum((float)marched);
      }
      // -- depth now, colour when the last row of this ray has been composited ---------------------
      if (lane == 0) {
        float dpt = __fdiv_rn(dep, R.nrm) * R.blend;                   // tensorBase.py:615
        if (B.accumulate) dpt = B.depth[B.depth_stride * ray] + dpt;
        B.depth[B.depth_stride * ray] = dpt;
        slot->acc = acc;
        slot->depth = dpt;
        __threadfence_block();
        const int old = atomicSub(&slot->pending, 1);                  // drop the open token
        if (old == 1) { __threadfence_block(); finalize_slot(B, slot); }
      }
      __syncwarp();
    }
    // -- the global ray queue is empty: help the rays still in flight in this CTA, step by step ------
    if (lane == 0) { __threadfence_block(); atomicAdd(&ctrl->draw_done, 1u); }
    if (stepwise) {
      for (;;) {
        bool any = false;
        for (int r = 1; r < nprod; ++r) {
          RayRec* other = recs + ((prod_id + r) % nprod);
          for (;;) {
            if (*reinterpret_cast<volatile unsigned int*>(&other->next_step) >= (unsigned)n_steps) break;
            unsigned int st = 0;
            if (lane == 0) st = atomicAdd(&other->next_step, 1u);
            st = __shfl_sync(0xffffffffu, st, 0);
            if (st >= (unsigned)n_steps) break;
            __threadfence_block();
            process_step<H16>(P, F, B, z_s, slots, other, (int)st, Q);
            any = true;
          }
        }
        if (!any) {
          if (*reinterpret_cast<volatile unsigned int*>(&ctrl->draw_done) >= (unsigned)nprod) break;
          __nanosleep(200);
        }
      }
    }
    if (Q.qn > 0) { flush_rows<H16>(P, F, z_s, slots, Q, Q.qn); Q.qn = 0; }
    // -- the last producer completes the open tile and publishes the tile count ----------------------
    if (lane == 0 && B.stats) { atomicAdd(B.stats, n_march); atomicAdd(B.stats + 1, Q.n_app); }
    unsigned int d = 0;
    if (lane == 0) { __threadfence_block(); d = atomicAdd(&ctrl->prod_done, 1u); }
    d = __shfl_sync(0xffffffffu, d, 0);
    if (d == (unsigned int)nprod - 1) {
      pad_open_tile(P);
      if (lane == 0) {
        *reinterpret_cast<volatile unsigned int*>(&ctrl->n_tiles_final) =
            *reinterpret_cast<volatile unsigned int*>(&ctrl->cursor) >> 7;
        __threadfence_block();
        *reinterpret_cast<volatile unsigned int*>(&ctrl->done) = 1u;
      }
    }
  } else if (warp == W_ISSUE) {
    // ======================================= MMA ISSUER ==========================================
    const uint32_t prep_a = smem_u32(smem + L.prep);
    uint32_t wpar = 0, xuse[2] = {0u, 0u};            // PE pipeline: parity of wrdy, uses of the two input-chunk buffers
    mbar_wait(smem_u32(&ctrl->bar_w), 0);
    for (unsigned int T = 0;; ++T) {
      const int b = (int)(T & 1);
      bool stop = false;
      while (!mbar_try(smem_u32(&ctrl->full[b]), (T >> 1) & 1)) {
        if (*reinterpret_cast<volatile unsigned int*>(&ctrl->done) &&
            T >= *reinterpret_cast<volatile unsigned int*>(&ctrl->n_tiles_final)) { stop = true; break; }
      }
      if (stop) break;
      if (PE) {
        // layer 0: acc0[128 x 32] = products x basis^T (three bf16 products), then layer 1 chunk by chunk
        const int nch = pe_chunks(F.fea_pe);
        const uint32_t wbuf = prep_a + PE_WBUF;
        const unsigned char* w1_g = reinterpret_cast<const unsigned char*>(F.prep) + PE_W1;
        if (lane == 0) {
          mbar_expect_tx(smem_u32(&ctrl->wrdy), 2 * PE_WC_BYTES);                 // chunk 0 streams in under layer 0
          tma_bulk_g2s(wbuf, w1_g, 2 * PE_WC_BYTES, smem_u32(&ctrl->wrdy));
          tc_fence_after();
          const uint32_t a1 = smem_u32(smem + L.a1 + b * (2 * OPER1_BYTES));
          const uint32_t sbo = (uint32_t)K1_CHUNKS * 128u, id0 = umma_idesc(PE_N0);
          uint32_t accf = 0;
          for (int ks = 0; ks < K1 / 16; ++ks) {
            const uint32_t ko = (uint32_t)ks * 256u;
            const uint64_t ah = umma_desc(a1 + ko, 128u, sbo), al = umma_desc(a1 + OPER1_BYTES + ko, 128u, sbo);
            const uint64_t bh = umma_desc(prep_a + PE_B0HI + ko, 128u, sbo), bl = umma_desc(prep_a + PE_B0LO + ko, 128u, sbo);
            umma_ss_id(tmem + TM_ACC0, ah, bh, id0, accf);
            umma_ss_id(tmem + TM_ACC0, ah, bl, id0, 1u);
            umma_ss_id(tmem + TM_ACC0, al, bh, id0, 1u);
            accf = 1u;
          }
          umma_commit(smem_u32(&ctrl->mma0));
        }
        __syncwarp();
        for (int c = 0; c < nch; ++c) {
          const int xb = c & 1;
          mbar_wait(smem_u32(&ctrl->wrdy), wpar); wpar ^= 1u;                      // weight chunk c is in shared memory
          mbar_wait(smem_u32(&ctrl->xrdy[xb]), xuse[xb] & 1u);                     // input chunk c is in TMEM
          if (lane == 0) {
            tc_fence_after();
            const uint32_t sbo = (uint32_t)(PE_KC / 8) * 128u, id1 = umma_idesc(FC);
            const uint32_t a_hi = tmem + TM_A2HI + xb * 32, a_lo = tmem + TM_A2LO + xb * 32;
            for (int ks = 0; ks < PE_KC / 16; ++ks) {
              const uint32_t ko = (uint32_t)ks * 256u, kc = (uint32_t)ks * 8u;
              const uint64_t bh = umma_desc(wbuf + ko, 128u, sbo), bl = umma_desc(wbuf + PE_WC_BYTES + ko, 128u, sbo);
              const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
              umma_ts_id(tmem + TM_ACC1, a_hi + kc, bh, id1, first);
              umma_ts_id(tmem + TM_ACC1, a_hi + kc, bl, id1, 1u);
              umma_ts_id(tmem + TM_ACC1, a_lo + kc, bh, id1, 1u);
            }
            umma_commit(smem_u32(&ctrl->xfree[xb]));
          }
          __syncwarp();
          mbar_wait(smem_u32(&ctrl->xfree[xb]), xuse[xb] & 1u);                    // chunk c consumed: buffers reusable
          xuse[xb]++;
          if (c + 1 < nch && lane == 0) {
            mbar_expect_tx(smem_u32(&ctrl->wrdy), 2 * PE_WC_BYTES);
            tma_bulk_g2s(wbuf, w1_g + (size_t)(c + 1) * 2 * PE_WC_BYTES, 2 * PE_WC_BYTES, smem_u32(&ctrl->wrdy));
          }
        }
        if (lane == 0) { tc_fence_after(); umma_commit(smem_u32(&ctrl->mma1)); }   // layer 1 complete
      } else if (lane == 0) {
        tc_fence_after();
        const uint32_t a1 = smem_u32(smem + L.a1 + b * (2 * OPER1_BYTES));
        const uint32_t sbo = (uint32_t)K1_CHUNKS * 128u;
        uint32_t accf = 0;
        for (int ks = 0; ks < K1 / 16; ++ks) {
          const uint32_t ko = (uint32_t)ks * 256u;
          const uint64_t ah = umma_desc(a1 + ko, 128u, sbo);
          const uint64_t al = umma_desc(a1 + OPER1_BYTES + ko, 128u, sbo);
          const uint64_t bh = umma_desc(prep_a + PREP_B1HI + ko, 128u, sbo);
          const uint64_t bl = umma_desc(prep_a + PREP_B1LO + ko, 128u, sbo);
          umma_bf16(tmem + TM_ACC1, ah, bh, accf);
          umma_bf16(tmem + TM_ACC1, ah, bl, 1u);
          umma_bf16(tmem + TM_ACC1, al, bh, 1u);
          accf = 1u;
        }
        umma_commit(smem_u32(&ctrl->mma1));
      }
      __syncwarp();
      mbar_wait(smem_u32(&ctrl->a2rdy), T & 1);       // layer-2 A operand is in TMEM
      if (lane == 0) {
        tc_fence_after();
        issue_layer_ts(tmem + TM_ACC2, tmem + TM_A2HI, tmem + TM_A2LO, prep_a + (PE ? PE_B2HI : PREP_B2HI),
                       prep_a + (PE ? PE_B2LO : PREP_B2LO), FC / 16, K2_CHUNKS, smem_u32(&ctrl->mma2));
      }
      __syncwarp();
    }
  } else if (warp >= W_CONS) {
    // ======================================= CONSUMERS ===========================================
    const int q4 = warp & 3, row = q4 * 32 + lane, ctid = tid - W_CONS * 32;
    const int chalf = (warp - W_CONS) >> 2, col0 = chalf * CONS_COLS;     // this thread's accumulator columns
    float* part_s = reinterpret_cast<float*>(smem + L.part);
    const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
    const float* tail = reinterpret_cast<const float*>(smem + L.prep + (PE ? PE_TAIL : PREP_TAIL));
    uint32_t xuse[2] = {0u, 0u};
    const float* b1_s = tail + TAIL_B1;
    const float* b2_s = tail + TAIL_B2;
    const float* W3_s = tail + TAIL_W3;
    const float* b3_s = tail + TAIL_B3;
    mbar_wait(smem_u32(&ctrl->bar_w), 0);
    for (unsigned int T = 0;; ++T) {
      const int b = (int)(T & 1);
      bool stop = false;
      while (!mbar_try(smem_u32(PE ? &ctrl->mma0 : &ctrl->mma1), T & 1)) {
        if (*reinterpret_cast<volatile unsigned int*>(&ctrl->done) &&
            T >= *reinterpret_cast<volatile unsigned int*>(&ctrl->n_tiles_final)) { stop = true; break; }
      }
      if (stop) break;
      tc_fence_after();
      // row metadata -> registers, then the A1 tile (and its metadata) may be rewritten
      const int my_slot = smem[L.mslot + b * TM + row];
      const float my_w = reinterpret_cast<const float*>(smem + L.mw)[b * TM + row];
      asm volatile("bar.sync 2, %0;" ::"n"(N_CONS * 32) : "memory");   // first MMA of tile T done + all metadata read ->
      if (ctid == 0) {                                  // the A1 buffer of tile T is free again
        __threadfence_block();
        *reinterpret_cast<volatile unsigned int*>(&ctrl->released) = T + 1u;
      }
      if (PE) {
        // -- layer-0 epilogue: the 27 basis features; encoded layer-1 input, 64 columns at a time ------------
        // x = [a, sin(a_d 2^f) (d-major, f-minor), cos(...)]   (tensorBase.py:14-21,115-125); zeros when !refine
        float a[32];
        tmem_ld32(t_row + (uint32_t)TM_ACC0, a);
        const int Fp = F.fea_pe, nsin = APP_DIM * Fp, nch = pe_chunks(Fp);
        if (chalf == 0) {
          for (int c = 0; c < nch; ++c) {
            const int xb = c & 1;
            if (xuse[xb] > 0) mbar_wait(smem_u32(&ctrl->xfree[xb]), (xuse[xb] - 1u) & 1u);   // previous occupant consumed
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {                       // 32 columns -> 16 packed words hi, 16 lo
              uint32_t hi[16], lo[16];
#pragma unroll 1
              for (int j2 = 0; j2 < 16; ++j2) {
                float v2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int j = c * PE_KC + h * 32 + 2 * j2 + u;
                  float v = 0.0f;
                  if (j < APP_DIM) v = a[j];
                  else if (B.refine && j < APP_DIM + 2 * nsin) {
                    const int e = (j - APP_DIM) % nsin, d = e / Fp, f = e - d * Fp;
                    const float ang = a[d] * (float)(1 << f);
                    v = (j - APP_DIM) < nsin ? sinf(ang) : cosf(ang);
                  }
                  v2[u] = v;
                }
                split2(v2[0], v2[1], hi[j2], lo[j2]);
              }
              tmem_st16(t_row + (uint32_t)(TM_A2HI + xb * 32 + h * 16), hi);
              tmem_st16(t_row + (uint32_t)(TM_A2LO + xb * 32 + h * 16), lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(smem_u32(&ctrl->xrdy[xb]));
            xuse[xb]++;
          }
        }
        mbar_wait(smem_u32(&ctrl->mma1), T & 1);         // every chunk of layer 1 has been accumulated
        tc_fence_after();
      }
      // -- epilogue 1: h1 = relu(acc1 + b1) -> bf16 hi/lo, layer 2's A operand in TMEM -------------
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + CONS_COLS; c0 += 32) {
        float v[32];
        tmem_ld32(t_row + (uint32_t)(TM_ACC1 + c0), v);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 bb = *reinterpret_cast<const float4*>(b1_s + c0 + 2 * j);
          split2(fmaxf(v[2 * j] + bb.x, 0.0f), fmaxf(v[2 * j + 1] + bb.y, 0.0f), hi[j], lo[j]);
          split2(fmaxf(v[2 * j + 2] + bb.z, 0.0f), fmaxf(v[2 * j + 3] + bb.w, 0.0f), hi[j + 1],
                 lo[j + 1]);
        }
        tmem_st16(t_row + (uint32_t)(TM_A2HI + c0 / 2), hi);
        tmem_st16(t_row + (uint32_t)(TM_A2LO + c0 / 2), lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&ctrl->a2rdy));
      // -- epilogue 2: h2 = relu(acc2 + b2); layer 3 + sigmoid (tensorBase.py:126-133) ---------------
      mbar_wait(smem_u32(&ctrl->mma2), T & 1);
      tc_fence_after();
      float pa[3] = {0.0f, 0.0f, 0.0f}, pb[3] = {0.0f, 0.0f, 0.0f};   // two chains per channel (ILP)
#pragma unroll 1
      for (int c0 = col0; c0 < col0 + CONS_COLS; c0 += 32) {
        float v[32];
        tmem_ld32(t_row + (uint32_t)(TM_ACC2 + c0), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(b2_s + c0 + j);
          const float h0 = fmaxf(v[j] + bb.x, 0.0f), h1 = fmaxf(v[j + 1] + bb.y, 0.0f);
          const float h2 = fmaxf(v[j + 2] + bb.z, 0.0f), h3 = fmaxf(v[j + 3] + bb.w, 0.0f);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float4 ww = *reinterpret_cast<const float4*>(W3_s + c * W3_LD + c0 + j);
            pa[c] = fmaf(ww.x, h0, pa[c]); pb[c] = fmaf(ww.y, h1, pb[c]);
            pa[c] = fmaf(ww.z, h2, pa[c]); pb[c] = fmaf(ww.w, h3, pb[c]);
          }
        }
      }
      tc_fence_before();
      if (N_CONS == 8) {                     // the two column halves of a row meet through shared memory
        if (chalf == 1) {
#pragma unroll
          for (int c = 0; c < 3; ++c) part_s[row * 4 + c] = pa[c] + pb[c];
        }
        asm volatile("bar.sync 3, %0;" ::"n"(N_CONS * 32) : "memory");
        if (chalf == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { pa[c] += pb[c]; pb[c] = part_s[row * 4 + c]; }
        }
      }
      // -- composite (tensorBase.py:632): w * rgb into the ray's fixed-point accumulators -------------
      if (my_slot != 0xFF && chalf == 0) {
        Slot* sl = slots + my_slot;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sc = pa[c] + pb[c];
          if (PE) sc += sl->vterm[c];
          else sc += W3_s[c * W3_LD + FC] * sl->vd[0] + W3_s[c * W3_LD + FC + 1] * sl->vd[1] +
                     W3_s[c * W3_LD + FC + 2] * sl->vd[2];
          sc += b3_s[c];
          sc = __fdiv_rn(1.0f, 1.0f + expf(-sc));
          atomicAdd(&sl->rgb[c], __float2uint_rn(my_w * sc * FIX_SCALE));
        }
        __threadfence_block();
        const int old = atomicSub(&sl->pending, 1);
        if (old == 1) { __threadfence_block(); finalize_slot(B, sl); }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == W_CONS) tmem_dealloc(tmem, TMEM_COLS);
  if (B.signal_seq && tid == 0) {
    // every pixel store of this CTA is ordered before the barrier above; the last CTA of the launch
    // publishes the step to all peers (release at system scope, after a cumulative fence)
    __threadfence_system();
    const unsigned int old = atomicAdd(B.done_ctr, 1u);
    if (old == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < B.n_peers; ++p) {
        unsigned long long* remote = B.peer_flags[p] + B.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(B.signal_seq) : "memory");
      }
    }
  }
}

// ---- host-side launcher --------------------------------------------------------------------------
int render_threads() { return THREADS; }

static int pick_nprod(int S, bool floater, int max_smem, bool pe = false) {
  int cap = MAX_PROD;
  if (const char* e = getenv("LRF_NPROD")) {      // tuning / debugging knob
    const int v = atoi(e);
    if (v >= 1 && v < cap) cap = v;
  }
  for (int n = cap; n >= 1; --n)
    if (smem_v3(S, floater, n, pe).total <= max_smem) return n;
  return 0;
}

size_t render_smem_bytes(int S, bool floater, int max_smem, bool pe) {
  const int n = pick_nprod(S, floater, max_smem, pe);
  return (size_t)smem_v3(S, floater, n ? n : 1, pe).total;
}

cudaError_t launch_render(const FieldDev& F, const BatchDev& B, int n_sms, int max_smem,
                          cudaStream_t stream) {
  const bool floater = B.floater_thresh > 0.0f;
  const bool pe = F.fea_pe > 0 || F.view_pe > 0;
  const int nprod = pick_nprod(F.S, floater, max_smem, pe);
  if (nprod < 1) return cudaErrorInvalidConfiguration;
  const size_t smem = (size_t)smem_v3(F.S, floater, nprod, pe).total;
  const int variant = pe ? 1 : (F.grid16 ? 2 : 0);   // (bf16 grids with positional encodings are refused by the ABI)
  static size_t configured[3][64] = {{0}, {0}, {0}};   // per kernel, per device: the attribute belongs to the device's context
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && smem > configured[variant][dev]) {
    e = variant == 1 ? cudaFuncSetAttribute(render_kernel_t<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
      : variant == 2 ? cudaFuncSetAttribute(render_kernel_t<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                     : cudaFuncSetAttribute(render_kernel_t<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[variant][dev] = smem;
  }
  e = cudaMemsetAsync(B.sched, 0, 2 * sizeof(unsigned long long), stream);   // ray counter + finished-CTA counter
  if (e != cudaSuccess) return e;
  // a CTA per SM, but never more CTAs than there are groups of 8 rays
  long long want = (B.n_rays + 7) / 8;
  int grid = (int)(want < n_sms ? want : n_sms);
  if (grid < 1) grid = 1;
  if (variant == 1) render_kernel_t<true, false><<<grid, THREADS, smem, stream>>>(F, B, nprod);
  else if (variant == 2) render_kernel_t<false, true><<<grid, THREADS, smem, stream>>>(F, B, nprod);
  else render_kernel_t<false, false><<<grid, THREADS, smem, stream>>>(F, B, nprod);
  return cudaGetLastError();
}

}  // namespace lrf
