// umma_probe.cu -- a one-CTA experiment, NOT part of the product library: runs
//     D[128][N] = A[128][K] * B[N][K]^T      (bf16 operands, fp32 accumulate in TMEM)
// with tcgen05.mma.kind::f16 for operand placements the render kernel does not use yet, so that the
// MLP-backward design of DESIGN.md par. 9.3 can be settled in a single GPU call:
//   * A from shared memory, K-major (the validated baseline) or MN-major (bit 15 of the instruction
//     descriptor): the MN-major read of an activation image [sample][feature] gives A^T without a
//     transposed copy -- what dW2 += dh2^T h1 needs;
//   * B from shared memory, K-major or MN-major (bit 16): the MN-major read of the forward weight
//     image [n][k] gives the B operand of dh1 = dh2 W2 without a second weight image;
//   * A from tensor memory (the validated .ts form) against an MN-major B.
// The host (tools/umma_probe.py) lays out the byte images and passes the descriptor fields
// (leading / stride byte offsets, per-K-step address advance) it wants to try; the kernel just issues
// the MMAs and returns the accumulator.  Build: see umma_probe.py.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../localrf_b200/csrc/lrf_device.cuh"

using namespace lrf;

struct ProbeParams {
  uint32_t idesc;                 // full instruction descriptor (formats, majors, N, M)
  int a_in_tmem;                  // 1: A rows are uint32 bf16 pairs, stored to TMEM by the CTA
  int n_ksteps;                   // K / 16
  int N;                          // accumulator columns to read back (multiple of 32 here)
  uint32_t a_bytes, b_bytes;      // image sizes
  uint32_t a_lbo, a_sbo, a_kstep; // shared-memory descriptor fields for A, byte advance per K-step
  uint32_t b_lbo, b_sbo, b_kstep;
};

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int PROBE_THREADS = 128;
constexpr int IMG_MAX = 65536;
constexpr int TM_A = 256;         // TMEM columns of the A operand (.ts form)

__global__ void __launch_bounds__(PROBE_THREADS, 1)
umma_probe_kernel(const ProbeParams P, const unsigned char* __restrict__ a_img,
                  const unsigned char* __restrict__ b_img, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_s = smem;
  unsigned char* b_s = smem + IMG_MAX;
  unsigned char* bars = smem + 2 * IMG_MAX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = smem_u32(bars), slot = bar + 8;
  if (tid == 0) mbar_init(bar, 1);
  if (!P.a_in_tmem)
    for (uint32_t o = tid * 16; o < P.a_bytes; o += PROBE_THREADS * 16)
      *reinterpret_cast<uint4*>(a_s + o) = *reinterpret_cast<const uint4*>(a_img + o);
  for (uint32_t o = tid * 16; o < P.b_bytes; o += PROBE_THREADS * 16)
    *reinterpret_cast<uint4*>(b_s + o) = *reinterpret_cast<const uint4*>(b_img + o);
  if (warp == 0) tmem_alloc(slot, TMEM_COLS);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bars + 8);
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  if (P.a_in_tmem) {                               // row tid: K/2 words, 16 at a time
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_img) + (size_t)tid * (P.n_ksteps * 8);
    for (int c = 0; c < P.n_ksteps * 8; c += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = src[c + j];
      tmem_st16(t_row + (uint32_t)(TM_A + c), r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
  }
  if (tid == 0) {
    tc_fence_after();
    uint32_t acc = 0;
    for (int ks = 0; ks < P.n_ksteps; ++ks) {
      const uint64_t bd = umma_desc(smem_u32(b_s) + ks * P.b_kstep, P.b_lbo, P.b_sbo);
      if (P.a_in_tmem) {
        mma_ts(tmem, tmem + TM_A + ks * 8, bd, P.idesc, acc);
      } else {
        const uint64_t ad = umma_desc(smem_u32(a_s) + ks * P.a_kstep, P.a_lbo, P.a_sbo);
        mma_ss(tmem, ad, bd, P.idesc, acc);
      }
      acc = 1u;
    }
    umma_commit(bar);
  }
  bool done = false;                               // bounded wait: a setting the hardware rejects must not hang the box
  for (int it = 0; it < (1 << 22) && !done; ++it) done = mbar_try(bar, 0);
  tc_fence_after();
  if (!done) {
    if (tid == 0) out[0] = -12345.0f;              // sentinel: the MMAs never completed
  } else
  for (int c0 = 0; c0 < P.N; c0 += 32) {
    float v[32];
    tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * P.N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

extern "C" int umma_probe_run(const ProbeParams* p, const void* a_img, const void* b_img, float* out,
                              void* stream) {
  if (!p || p->a_bytes > IMG_MAX || p->b_bytes > IMG_MAX || (p->a_bytes | p->b_bytes) & 15) return -1;
  if (p->N % 32 || p->N < 32 || p->N > 256 || p->n_ksteps < 1 || (p->a_in_tmem && p->n_ksteps % 2)) return -1;
  const int smem = 2 * IMG_MAX + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return -3;
  umma_probe_kernel<<<1, PROBE_THREADS, smem, (cudaStream_t)stream>>>(
      *p, static_cast<const unsigned char*>(a_img), static_cast<const unsigned char*>(b_img), out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
