// lrf_abi.cu -- the C ABI declared in include/localrf_b200.h (validation + launches).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/localrf_b200.h"
#include "lrf_common.cuh"

namespace lrf {
size_t render_smem_bytes(int S, bool floater, int max_smem, bool pe);
int render_threads();
cudaError_t launch_render(const FieldDev& F, const BatchDev& B, int n_sms, int max_smem,
                          cudaStream_t stream);
cudaError_t launch_prepare(const float* basis, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, unsigned char* prep,
                           cudaStream_t stream);
cudaError_t launch_prepare_pe(const float* basis, const float* w1, const float* b1, const float* w2,
                              const float* b2, const float* w3, const float* b3, int fea_pe, int view_pe,
                              unsigned char* prep, cudaStream_t stream);
cudaError_t launch_mlp(const float* prep, const float* feats, const float* viewdirs, long long M,
                       float* rgb, int n_sms, cudaStream_t stream);
cudaError_t launch_density_feature(const FieldDev& F, const float* xyz, long long M, float* out,
                                   cudaStream_t stream);
cudaError_t launch_app_feature(const FieldDev& F, const float* basis, const float* xyz,
                               long long M, float* out, cudaStream_t stream);
cudaError_t launch_repack(const float* src, float* dst, int C, long long HW, cudaStream_t stream);
cudaError_t launch_pack_bf16(const float* src, void* dst, long long n, int n_sms, cudaStream_t stream);
cudaError_t launch_alpha_mask_build(const FieldDev& F, const float* aabb_max, const int* dims, float length,
                                    float thres, float* alpha_scratch, float* mask,
                                    unsigned long long* kept, int n_sms, cudaStream_t stream);
cudaError_t launch_upsample(const float* src, int H, int W, float* dst, int H2, int W2, int C, int n_sms,
                            cudaStream_t stream);
cudaError_t launch_density_l1(const FieldDev& F, double* sum, int n_sms, cudaStream_t stream);
cudaError_t launch_density_l1_backward(const FieldDev& F, const float* gout, float* const* d_plane,
                                       float* const* d_line, cudaStream_t stream);
cudaError_t launch_tv(const float* x, int H, int W, int C, double* sums, int n_sms, cudaStream_t stream);
cudaError_t launch_tv_backward(const float* x, int H, int W, int C, const float* gout, float kh, float kw,
                               float* dx, int n_sms, cudaStream_t stream);
cudaError_t launch_sample_ray(const float* rays, const float* jitter, long long N, int S, const float* aabb,
                              float near, float far, float step, float* pts, float* z,
                              unsigned char* inside, int n_sms, cudaStream_t stream);
cudaError_t launch_peer_barrier(unsigned long long* const* peer_flags, int rank, int world,
                                unsigned long long seq, unsigned long long wait_seq, cudaStream_t stream);
cudaError_t launch_frame_u8(const float* rgb, int rgb_stride, const float* depth, int depth_stride, long long N,
                            float d_lo, float d_hi, const unsigned char* lut, unsigned char* rgb8,
                            unsigned char* depth8, int n_sms, cudaStream_t stream);
cudaError_t launch_app_products(const FieldDev& F, const float* xyz, long long M, float* out,
                                cudaStream_t stream);
cudaError_t launch_density_backward(const FieldDev& F, float* const* d_plane, float* const* d_line,
                                    const float* xyz, const float* gout, long long M, float* dxyz,
                                    cudaStream_t stream);
cudaError_t launch_app_products_backward(const FieldDev& F, float* const* d_plane,
                                         float* const* d_line, const float* xyz, const float* gout,
                                         long long M, float* dxyz, cudaStream_t stream);
size_t backward_prepared_bytes();
size_t backward_scratch_bytes(long long n_rays, int S);
cudaError_t launch_prepare_backward(const float* basis, const float* w1, const float* b1,
                                    const float* w2, const float* b2, const float* w3,
                                    const float* b3, float* bp, cudaStream_t stream);
cudaError_t launch_render_backward_abi(const FieldDev& F, const float* rays, long long n_rays,
                                       int white_bg, const float* g_rgb, const float* g_depth,
                                       const float* bp, const LrfGradients& G, void* scratch,
                                       int n_sms, cudaStream_t stream);
}  // namespace lrf

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return LRF_ERR_CUDA;
}

struct DevInfo {
  int n_sms = 0;
  int max_smem = 0;
  bool ok = false;
  unsigned long long* sched = nullptr;   // ring of per-launch ray counters (device memory)
  unsigned int next = 0;
};
constexpr unsigned int SCHED_RING = 256;

// per-device launch resources, process-global (one ring of launch counters per device, allocated on first
// use by whichever host thread gets there first; the autograd engine's worker threads share it)
int device_info(DevInfo& d) {
  static DevInfo cache[64];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(LRF_ERR_INVALID, "device index out of range");
  if (!cache[dev].ok) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    cache[dev].n_sms = p.multiProcessorCount;
    cache[dev].max_smem = (int)p.sharedMemPerBlockOptin;
    e = cudaMalloc(&cache[dev].sched, 2 * SCHED_RING * sizeof(unsigned long long));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(ray counters)");
    cache[dev].ok = true;
  }
  cache[dev].next = (cache[dev].next + 1) % SCHED_RING;
  d = cache[dev];
  return LRF_OK;
}

// validates the parts of a field every entry point needs and fills the device-side view
int make_field(const LrfField* f, bool need_mlp, bool need_table, const void* prepared,
               lrf::FieldDev& F, bool allow_pe = false, bool allow_bf16 = false) {
  if (!f) return fail(LRF_ERR_INVALID, "field is NULL");
  if (f->grid_dtype != LRF_GRID_F32 && f->grid_dtype != LRF_GRID_BF16)
    return fail(LRF_ERR_INVALID, "grid_dtype must be LRF_GRID_F32 or LRF_GRID_BF16");
  if (f->grid_dtype == LRF_GRID_BF16 && !allow_bf16)
    return fail(LRF_ERR_UNSUPPORTED, "bf16 grid storage is built for lrf_render (pe = 0), lrf_density_feature and "
                                     "lrf_app_feature only (inference); pass the fp32 parameters here");
  F.grid16 = f->grid_dtype == LRF_GRID_BF16 ? 1 : 0;
  if (f->n_dcomp != lrf::CD || f->n_acomp != lrf::CA)
    return fail(LRF_ERR_UNSUPPORTED, "only density_n_comp=8 / appearance_n_comp=24 per plane are built");
  for (int a = 0; a < 3; ++a) {
    if (f->grid[a] < 2) return fail(LRF_ERR_INVALID, "gridSize must be >= 2 on every axis");
    if (!(f->aabb[3 + a] > f->aabb[a])) return fail(LRF_ERR_INVALID, "aabb max must exceed aabb min");
    F.g[a] = f->grid[a];
    F.amin[a] = f->aabb[a];
    F.ainv[a] = 2.0f / (f->aabb[3 + a] - f->aabb[a]);
  }
  for (int i = 0; i < 3; ++i) {
    if (!f->dplane[i] || !f->dline[i] || !f->aplane[i] || !f->aline[i])
      return fail(LRF_ERR_INVALID, "plane/line pointer is NULL");
    if (((uintptr_t)f->dplane[i] | (uintptr_t)f->dline[i] | (uintptr_t)f->aplane[i] |
         (uintptr_t)f->aline[i]) & 15)
      return fail(LRF_ERR_INVALID, "plane/line pointers must be 16-byte aligned");
    F.dplane[i] = f->dplane[i]; F.dline[i] = f->dline[i];
    F.aplane[i] = f->aplane[i]; F.aline[i] = f->aline[i];
  }
  F.prep = static_cast<const float*>(prepared);
  F.alpha_vol = f->alpha_vol;
  if (f->alpha_vol) {
    for (int a = 0; a < 3; ++a) {
      if (f->alpha_dims[a] < 1) return fail(LRF_ERR_INVALID, "alpha mask dims must be >= 1");
      F.ad[a] = f->alpha_dims[a];
      F.aamin[a] = f->alpha_aabb[a];
      F.aainv[a] = 1.0f / (f->alpha_aabb[3 + a] - f->alpha_aabb[a]) * 2.0f;
    }
  } else {
    for (int a = 0; a < 3; ++a) { F.ad[a] = 1; F.aamin[a] = 0.f; F.aainv[a] = 0.f; }
  }
  F.density_shift = f->density_shift;
  F.distance_scale = f->distance_scale;
  F.weight_thres = f->weight_thres;
  if (f->act != 0 && f->act != 1) return fail(LRF_ERR_INVALID, "act must be 0 (softplus) or 1 (relu)");
  F.act = f->act;
  if (need_mlp) {
    if (f->app_dim != lrf::APP_DIM || f->featureC != lrf::FC)
      return fail(LRF_ERR_UNSUPPORTED, "only app_dim=27 / featureC=128 are built");
    if (f->fea_pe < 0 || f->view_pe < 0) return fail(LRF_ERR_INVALID, "fea_pe / view_pe must be >= 0");
    if ((f->fea_pe != 0 || f->view_pe != 0) && !allow_pe)
      return fail(LRF_ERR_UNSUPPORTED, "positional encodings (fea_pe/view_pe > 0) are built for lrf_render only");
    if (f->fea_pe > lrf::PE_MAX_FEA || f->view_pe > lrf::PE_MAX_VIEW)
      return fail(LRF_ERR_UNSUPPORTED, "fea_pe / view_pe above 8 are not built");
  }
  F.fea_pe = need_mlp ? f->fea_pe : 0;
  F.view_pe = need_mlp ? f->view_pe : 0;
  F.w3 = f->w3;
  F.z = f->z_vals;
  F.S = f->n_samples;
  if (need_table) {
    if (!f->z_vals) return fail(LRF_ERR_INVALID, "z_vals is NULL");
    if (f->n_samples < 2 || f->n_samples > 65535)
      return fail(LRF_ERR_INVALID, "n_samples must be in [2, 65535]");
  }
  return LRF_OK;
}

}  // namespace

extern "C" {

int lrf_version(void) { return 3; }

size_t lrf_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(LrfField);
    case 1: return sizeof(LrfBatch);
    case 2: return sizeof(LrfOutputs);
    case 3: return sizeof(LrfGradients);
    default: return 0;
  }
}

const char* lrf_last_error(void) { return g_err; }

size_t lrf_prepared_bytes(void) { return (size_t)lrf::PREP_BYTES; }

size_t lrf_prepared_bytes_for(const LrfField* f) {
  if (!f || f->fea_pe < 0 || f->view_pe < 0 || f->fea_pe > lrf::PE_MAX_FEA || f->view_pe > lrf::PE_MAX_VIEW) return 0;
  if (f->fea_pe == 0 && f->view_pe == 0) return (size_t)lrf::PREP_BYTES;
  return (size_t)lrf::pe_prepared_bytes(f->fea_pe);
}

int lrf_field_prepare(const LrfField* f, void* prepared, lrf_stream_t stream) {
  if (!f || !prepared) return fail(LRF_ERR_INVALID, "field or prepared is NULL");
  if ((uintptr_t)prepared & 15) return fail(LRF_ERR_INVALID, "prepared must be 16-byte aligned");
  if (f->app_dim != lrf::APP_DIM || f->featureC != lrf::FC || f->n_acomp != lrf::CA)
    return fail(LRF_ERR_UNSUPPORTED, "only app_dim=27 / featureC=128 / appearance_n_comp=24 are built");
  if (f->fea_pe < 0 || f->view_pe < 0 || f->fea_pe > lrf::PE_MAX_FEA || f->view_pe > lrf::PE_MAX_VIEW)
    return fail(LRF_ERR_UNSUPPORTED, "fea_pe / view_pe must be in [0, 8]");
  if (!f->basis || !f->w1 || !f->b1 || !f->w2 || !f->b2 || !f->w3 || !f->b3)
    return fail(LRF_ERR_INVALID, "MLP / basis pointer is NULL");
  if (f->fea_pe != 0 || f->view_pe != 0) {
    if ((uintptr_t)prepared & 1023) return fail(LRF_ERR_INVALID, "the prepared block of a field with positional encodings must be 1024-byte aligned");
    cudaError_t e = lrf::launch_prepare_pe(f->basis, f->w1, f->b1, f->w2, f->b2, f->w3, f->b3, f->fea_pe, f->view_pe,
                                           static_cast<unsigned char*>(prepared), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "prepare_pe_kernel");
    return LRF_OK;
  }
  cudaError_t e = lrf::launch_prepare(f->basis, f->w1, f->b1, f->w2, f->b2, f->w3, f->b3,
                                      static_cast<unsigned char*>(prepared), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "prepare_kernel");
  return LRF_OK;
}

int lrf_render(const LrfField* f, const void* prepared, const LrfBatch* b, const LrfOutputs* o,
               lrf_stream_t stream) {
  if (!b || !o) return fail(LRF_ERR_INVALID, "batch or outputs is NULL");
  if (!prepared) return fail(LRF_ERR_INVALID, "prepared is NULL (call lrf_field_prepare first)");
  lrf::FieldDev F;
  int rc = make_field(f, true, true, prepared, F, /*allow_pe=*/true, /*allow_bf16=*/true);
  if (rc != LRF_OK) return rc;
  if ((F.fea_pe || F.view_pe) && (!f->w3 || ((uintptr_t)prepared & 1023)))
    return fail(LRF_ERR_INVALID, "positional encodings need w3 and a 1024-byte aligned prepared block");
  if ((F.fea_pe || F.view_pe) && F.grid16)
    return fail(LRF_ERR_UNSUPPORTED, "bf16 grid storage is built for fields without positional encodings");
  if (b->n_rays < 0) return fail(LRF_ERR_INVALID, "n_rays < 0");
  if (b->n_rays == 0) {
    // nothing to render; a rank with an empty shard must still publish its step to the peers
    if (o->signal_seq && o->n_peers >= 1 && o->rank >= 0 && o->rank < o->n_peers) {
      for (int p = 0; p < o->n_peers; ++p)
        if (!o->peer_flags[p]) return fail(LRF_ERR_INVALID, "peer_flags pointer is NULL");
      cudaError_t e = lrf::launch_peer_barrier(o->peer_flags, o->rank, o->n_peers, o->signal_seq, 0ull,
                                               (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_fail(e, "peer_barrier_kernel");
    }
    return LRF_OK;
  }
  if (!o->pix && (!o->rgb || !o->depth)) return fail(LRF_ERR_INVALID, "rgb/depth output is NULL");
  lrf::BatchDev B;
  memset(&B, 0, sizeof(B));
  B.n_rays = b->n_rays;
  B.rays = b->rays;
  B.rays_per_view = 0;
  const bool per_view = b->blend || b->exposure || !b->rays;
  if (per_view) {
    if (b->n_views < 1 || b->n_rays % b->n_views != 0)
      return fail(LRF_ERR_INVALID, "n_rays must be a positive multiple of n_views");
    B.rays_per_view = b->n_rays / b->n_views;
  }
  if (!b->rays) {
    if (!b->ray_ids || !b->cam2world) return fail(LRF_ERR_INVALID, "ray_ids/cam2world is NULL");
    if (b->W < 1 || b->H < 1) return fail(LRF_ERR_INVALID, "W/H must be positive");
    if (!b->fov360 && !b->intrinsics && !(b->focal != 0.0f))
      return fail(LRF_ERR_INVALID, "focal must be non-zero");
  }
  B.ray_ids = reinterpret_cast<const long long*>(b->ray_ids);
  B.W = b->W; B.H = b->H; B.fov360 = b->fov360;
  B.focal = b->focal; B.cx = b->cx; B.cy = b->cy;
  B.intrinsics = b->intrinsics;
  B.c2w = b->cam2world;
  B.w2rf = b->world2rf;
  B.blend = b->blend;
  B.blend_stride = b->blend_stride;
  B.exposure = b->exposure;
  B.accumulate = b->accumulate; B.finalize = b->finalize; B.white_bg = b->white_bg;
  B.floater_thresh = b->floater_thresh;
  B.refine = b->refine;
  if (o->pix) {
    B.rgb = o->pix; B.depth = o->pix + 3; B.rgb_stride = 4; B.depth_stride = 4;
  } else {
    B.rgb = o->rgb; B.depth = o->depth; B.rgb_stride = 3; B.depth_stride = 1;
  }
  B.weights = o->weights;
  B.dirs = b->rays ? nullptr : o->directions;
  B.ij = b->rays ? nullptr : reinterpret_cast<long long*>(o->ij);
  B.stats = o->stats;
  if (o->n_peers < 0 || o->n_peers > LRF_MAX_PEERS) return fail(LRF_ERR_INVALID, "n_peers out of range");
  B.n_peers = o->n_peers;
  B.mc_pix = nullptr;
  if (o->n_peers > 0) {
    if (!o->pix) return fail(LRF_ERR_INVALID, "the fused pixel exchange needs the interleaved `pix` output");
    if (o->mc_pix) {
      if ((uintptr_t)o->mc_pix & 15) return fail(LRF_ERR_INVALID, "mc_pix must be 16-byte aligned");
      B.mc_pix = o->mc_pix;
    } else {
      for (int p = 0; p < o->n_peers; ++p) {
        if (!o->peer_pix[p] || ((uintptr_t)o->peer_pix[p] & 15))
          return fail(LRF_ERR_INVALID, "peer_pix pointers must be non-NULL and 16-byte aligned");
        B.peer_pix[p] = o->peer_pix[p];
      }
    }
  }
  B.signal_seq = 0; B.wait_seq = 0; B.rank = 0;
  if (o->signal_seq) {
    if (o->n_peers < 1 || o->rank < 0 || o->rank >= o->n_peers || o->wait_seq > o->signal_seq)
      return fail(LRF_ERR_INVALID, "in-kernel signalling needs n_peers >= 1, 0 <= rank < n_peers, wait_seq <= signal_seq");
    for (int p = 0; p < o->n_peers; ++p) {
      if (!o->peer_flags[p] || ((uintptr_t)o->peer_flags[p] & 7))
        return fail(LRF_ERR_INVALID, "peer_flags pointers must be non-NULL and 8-byte aligned");
      B.peer_flags[p] = o->peer_flags[p];
    }
    B.signal_seq = o->signal_seq; B.wait_seq = o->wait_seq; B.rank = o->rank;
  }
  DevInfo d;
  rc = device_info(d);
  if (rc != LRF_OK) return rc;
  const size_t smem = lrf::render_smem_bytes(F.S, B.floater_thresh > 0.0f, d.max_smem, F.fea_pe > 0 || F.view_pe > 0);
  if ((long long)smem > d.max_smem)
    return fail(LRF_ERR_UNSUPPORTED, "sample table too long for the shared-memory budget");
  B.sched = d.sched + 2 * d.next;                                   // [ray counter, finished-CTA counter]
  B.done_ctr = reinterpret_cast<unsigned int*>(d.sched + 2 * d.next + 1);
  cudaError_t e = lrf::launch_render(F, B, d.n_sms, d.max_smem, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "render_kernel");
  return LRF_OK;
}

int lrf_mlp_forward(const void* prepared, const float* feats, const float* viewdirs, int64_t M,
                    float* rgb, lrf_stream_t stream) {
  if (!prepared) return fail(LRF_ERR_INVALID, "prepared is NULL (call lrf_field_prepare first)");
  if (M < 0 || (M > 0 && (!feats || !viewdirs || !rgb))) return fail(LRF_ERR_INVALID, "bad feats/viewdirs/rgb/M");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_mlp(static_cast<const float*>(prepared), feats, viewdirs, M, rgb,
                                  d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "mlp_kernel");
  return LRF_OK;
}

int lrf_density_feature(const LrfField* f, const float* xyz, int64_t M, float* out,
                        lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F, false, /*allow_bf16=*/true);
  if (rc != LRF_OK) return rc;
  if (M < 0 || (M > 0 && (!xyz || !out))) return fail(LRF_ERR_INVALID, "bad xyz/out/M");
  cudaError_t e = lrf::launch_density_feature(F, xyz, M, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "density_feature_kernel");
  return LRF_OK;
}

int lrf_app_feature(const LrfField* f, const float* xyz, int64_t M, float* out,
                    lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F, false, /*allow_bf16=*/true);
  if (rc != LRF_OK) return rc;
  if (f->app_dim != lrf::APP_DIM) return fail(LRF_ERR_UNSUPPORTED, "only app_dim=27 is built");
  if (!f->basis) return fail(LRF_ERR_INVALID, "basis is NULL");
  if (M < 0 || (M > 0 && (!xyz || !out))) return fail(LRF_ERR_INVALID, "bad xyz/out/M");
  cudaError_t e = lrf::launch_app_feature(F, f->basis, xyz, M, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "app_feature_kernel");
  return LRF_OK;
}

int lrf_app_products(const LrfField* f, const float* xyz, int64_t M, float* out,
                     lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if (M < 0 || (M > 0 && (!xyz || !out))) return fail(LRF_ERR_INVALID, "bad xyz/out/M");
  cudaError_t e = lrf::launch_app_products(F, xyz, M, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "app_products_kernel");
  return LRF_OK;
}

static int check_grads(float* const d_plane[3], float* const d_line[3]) {
  if (!d_plane || !d_line) return fail(LRF_ERR_INVALID, "gradient pointer arrays are NULL");
  for (int i = 0; i < 3; ++i)
    if (!d_plane[i] || !d_line[i]) return fail(LRF_ERR_INVALID, "gradient buffer is NULL");
  return LRF_OK;
}

int lrf_density_feature_backward(const LrfField* f, const float* xyz, const float* gout, int64_t M,
                                 float* const d_plane[3], float* const d_line[3], float* d_xyz,
                                 lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if ((rc = check_grads(d_plane, d_line)) != LRF_OK) return rc;
  if (M < 0 || (M > 0 && (!xyz || !gout))) return fail(LRF_ERR_INVALID, "bad xyz/grad_out/M");
  cudaError_t e = lrf::launch_density_backward(F, d_plane, d_line, xyz, gout, M, d_xyz,
                                               (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "density_backward_kernel");
  return LRF_OK;
}

int lrf_app_products_backward(const LrfField* f, const float* xyz, const float* gout, int64_t M,
                              float* const d_plane[3], float* const d_line[3], float* d_xyz,
                              lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if ((rc = check_grads(d_plane, d_line)) != LRF_OK) return rc;
  if (M < 0 || (M > 0 && (!xyz || !gout))) return fail(LRF_ERR_INVALID, "bad xyz/grad_out/M");
  cudaError_t e = lrf::launch_app_products_backward(F, d_plane, d_line, xyz, gout, M, d_xyz,
                                                    (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "app_products_backward_kernel");
  return LRF_OK;
}

size_t lrf_prepared_backward_bytes(void) { return lrf::backward_prepared_bytes(); }

size_t lrf_backward_scratch_bytes(int64_t n_rays, int32_t n_samples) {
  if (n_rays < 0 || n_samples < 2) return 0;
  return lrf::backward_scratch_bytes(n_rays, n_samples);
}

int lrf_field_prepare_backward(const LrfField* f, void* prepared_bwd, lrf_stream_t stream) {
  if (!f || !prepared_bwd) return fail(LRF_ERR_INVALID, "field or prepared_bwd is NULL");
  if ((uintptr_t)prepared_bwd & 15) return fail(LRF_ERR_INVALID, "prepared_bwd must be 16-byte aligned");
  if (f->app_dim != lrf::APP_DIM || f->featureC != lrf::FC || f->n_acomp != lrf::CA)
    return fail(LRF_ERR_UNSUPPORTED, "only app_dim=27 / featureC=128 / appearance_n_comp=24 are built");
  if (f->fea_pe != 0 || f->view_pe != 0)
    return fail(LRF_ERR_UNSUPPORTED, "positional encodings (fea_pe/view_pe > 0) are not built yet");
  if (!f->basis || !f->w1 || !f->b1 || !f->w2 || !f->b2 || !f->w3 || !f->b3)
    return fail(LRF_ERR_INVALID, "MLP / basis pointer is NULL");
  cudaError_t e = lrf::launch_prepare_backward(f->basis, f->w1, f->b1, f->w2, f->b2, f->w3, f->b3,
                                               static_cast<float*>(prepared_bwd), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "prepare_backward_kernel");
  return LRF_OK;
}

int lrf_render_backward(const LrfField* f, const void* prepared_bwd, const float* rays,
                        int64_t n_rays, int32_t white_bg, const float* grad_rgb,
                        const float* grad_depth, const LrfGradients* g, void* scratch,
                        size_t scratch_bytes, lrf_stream_t stream) {
  if (!prepared_bwd) return fail(LRF_ERR_INVALID, "prepared_bwd is NULL (call lrf_field_prepare_backward first)");
  lrf::FieldDev F;
  int rc = make_field(f, true, true, nullptr, F);
  if (rc != LRF_OK) return rc;
  if (n_rays < 0) return fail(LRF_ERR_INVALID, "n_rays < 0");
  if (n_rays == 0) return LRF_OK;
  if (n_rays > 2147483647LL) return fail(LRF_ERR_INVALID, "n_rays must fit 31 bits");
  if (!rays || !grad_rgb || !grad_depth || !g) return fail(LRF_ERR_INVALID, "rays / grad_rgb / grad_depth / grads is NULL");
  if (!g->d_rays || !g->d_w1b || !g->d_b1 || !g->d_w2 || !g->d_b2 || !g->d_w3 || !g->d_b3)
    return fail(LRF_ERR_INVALID, "gradient buffer is NULL");
  for (int i = 0; i < 3; ++i) {
    if (!g->d_dplane[i] || !g->d_dline[i] || !g->d_aplane[i] || !g->d_aline[i])
      return fail(LRF_ERR_INVALID, "gradient buffer is NULL");
    if (((uintptr_t)g->d_dplane[i] | (uintptr_t)g->d_dline[i] | (uintptr_t)g->d_aplane[i] |
         (uintptr_t)g->d_aline[i]) & 15)
      return fail(LRF_ERR_INVALID, "plane/line gradient buffers must be 16-byte aligned");
  }
  if (!scratch || ((uintptr_t)scratch & 15)) return fail(LRF_ERR_INVALID, "scratch is NULL or not 16-byte aligned");
  if (scratch_bytes < lrf::backward_scratch_bytes(n_rays, F.S))
    return fail(LRF_ERR_INVALID, "scratch is smaller than lrf_backward_scratch_bytes(n_rays, n_samples)");
  DevInfo d;
  rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_render_backward_abi(F, rays, n_rays, white_bg, grad_rgb, grad_depth,
                                                  static_cast<const float*>(prepared_bwd), *g, scratch,
                                                  d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "render_backward");
  return LRF_OK;
}

int lrf_alpha_mask_build(const LrfField* f, const int32_t dims[3], float length, float thres,
                         float* alpha_scratch, float* mask, unsigned long long* kept, lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if (!dims || dims[0] < 1 || dims[1] < 1 || dims[2] < 1) return fail(LRF_ERR_INVALID, "lattice dims must be >= 1");
  if (!alpha_scratch) return fail(LRF_ERR_INVALID, "alpha_scratch is NULL");
  if (mask && !kept) return fail(LRF_ERR_INVALID, "kept is NULL");
  DevInfo d;
  if ((rc = device_info(d)) != LRF_OK) return rc;
  const int dd[3] = {dims[0], dims[1], dims[2]};
  cudaError_t e = lrf::launch_alpha_mask_build(F, f->aabb + 3, dd, length, thres, alpha_scratch, mask, kept,
                                               d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "alpha_mask_build");
  return LRF_OK;
}

int lrf_upsample(const float* src, int32_t H, int32_t W, float* dst, int32_t H2, int32_t W2, int32_t C,
                 lrf_stream_t stream) {
  if (!src || !dst || H < 1 || W < 1 || H2 < 1 || W2 < 1 || C < 4 || (C & 3))
    return fail(LRF_ERR_INVALID, "bad upsample arguments (C must be a positive multiple of 4)");
  if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(LRF_ERR_INVALID, "upsample buffers must be 16-byte aligned");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_upsample(src, H, W, dst, H2, W2, C, d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "upsample_kernel");
  return LRF_OK;
}

int lrf_density_l1(const LrfField* f, double* sum, lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if (!sum) return fail(LRF_ERR_INVALID, "sum is NULL");
  DevInfo d;
  if ((rc = device_info(d)) != LRF_OK) return rc;
  cudaError_t e = lrf::launch_density_l1(F, sum, d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "density_l1_fwd_kernel");
  return LRF_OK;
}

int lrf_density_l1_backward(const LrfField* f, const float* grad_out, float* const d_plane[3],
                            float* const d_line[3], lrf_stream_t stream) {
  lrf::FieldDev F;
  int rc = make_field(f, false, false, nullptr, F);
  if (rc != LRF_OK) return rc;
  if (!grad_out) return fail(LRF_ERR_INVALID, "grad_out is NULL");
  if (!d_plane || !d_line) return fail(LRF_ERR_INVALID, "gradient pointer arrays are NULL");
  for (int i = 0; i < 3; ++i)
    if (!d_plane[i] || !d_line[i]) return fail(LRF_ERR_INVALID, "gradient buffer is NULL");
  cudaError_t e = lrf::launch_density_l1_backward(F, grad_out, d_plane, d_line, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "density_l1_bwd_kernel");
  return LRF_OK;
}

int lrf_tv_sums(const float* x, int32_t H, int32_t W, int32_t C, double* sums, lrf_stream_t stream) {
  if (!x || !sums || H < 1 || W < 1 || C < 1) return fail(LRF_ERR_INVALID, "bad tv arguments");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_tv(x, H, W, C, sums, d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "tv_fwd_kernel");
  return LRF_OK;
}

int lrf_tv_sums_backward(const float* x, int32_t H, int32_t W, int32_t C, const float* grad_out, float kh,
                         float kw, float* dx, lrf_stream_t stream) {
  if (!x || !dx || !grad_out || H < 1 || W < 1 || C < 1) return fail(LRF_ERR_INVALID, "bad tv arguments");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_tv_backward(x, H, W, C, grad_out, kh, kw, dx, d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "tv_bwd_kernel");
  return LRF_OK;
}

int lrf_sample_ray(const float* rays, const float* jitter, int64_t N, int32_t S, const float aabb[6],
                   float near, float far, float step, float* pts, float* z, unsigned char* inside,
                   lrf_stream_t stream) {
  if (N < 0 || S < 1 || !aabb) return fail(LRF_ERR_INVALID, "bad sample_ray sizes");
  if (N > 0 && (!rays || !pts || !z || !inside)) return fail(LRF_ERR_INVALID, "sample_ray buffer is NULL");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_sample_ray(rays, jitter, N, S, aabb, near, far, step, pts, z, inside, d.n_sms,
                                         (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "sample_ray_kernel");
  return LRF_OK;
}

int lrf_frame_to_u8(const float* rgb, int32_t rgb_stride, const float* depth, int32_t depth_stride, int64_t N,
                    float d_lo, float d_hi, const unsigned char* lut, unsigned char* rgb8, unsigned char* depth8,
                    lrf_stream_t stream) {
  if (N < 0) return fail(LRF_ERR_INVALID, "N < 0");
  if (rgb8 && (!rgb || rgb_stride < 3)) return fail(LRF_ERR_INVALID, "rgb8 requested without rgb / stride >= 3");
  if (depth8 && (!depth || depth_stride < 1 || !lut)) return fail(LRF_ERR_INVALID, "depth8 requested without depth / lut");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_frame_u8(rgb, rgb_stride, depth, depth_stride, N, d_lo, d_hi, lut, rgb8, depth8,
                                       d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "frame_u8_kernel");
  return LRF_OK;
}

int lrf_peer_barrier(unsigned long long* const* peer_flags, int32_t rank, int32_t world,
                     unsigned long long seq, lrf_stream_t stream) {
  return lrf_peer_signal_wait(peer_flags, rank, world, seq, seq, stream);
}

int lrf_peer_signal_wait(unsigned long long* const* peer_flags, int32_t rank, int32_t world,
                         unsigned long long seq, unsigned long long wait_seq, lrf_stream_t stream) {
  if (wait_seq > seq) return fail(LRF_ERR_INVALID, "wait_seq must not exceed seq (a rank cannot wait for a step it has not signalled)");
  if (!peer_flags || world < 1 || world > LRF_MAX_PEERS || rank < 0 || rank >= world)
    return fail(LRF_ERR_INVALID, "bad peer_flags / rank / world");
  for (int p = 0; p < world; ++p)
    if (!peer_flags[p] || ((uintptr_t)peer_flags[p] & 7))
      return fail(LRF_ERR_INVALID, "peer flag arrays must be non-NULL and 8-byte aligned");
  cudaError_t e = lrf::launch_peer_barrier(peer_flags, rank, world, seq, wait_seq, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "peer_barrier_kernel");
  return LRF_OK;
}

int lrf_pack_bf16(const float* src, void* dst, int64_t n, lrf_stream_t stream) {
  if (n < 0 || (n > 0 && (!src || !dst))) return fail(LRF_ERR_INVALID, "bad src/dst/n");
  if ((uintptr_t)dst & 1) return fail(LRF_ERR_INVALID, "dst must be 2-byte aligned");
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  cudaError_t e = lrf::launch_pack_bf16(src, dst, n, d.n_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack_bf16_kernel");
  return LRF_OK;
}

int lrf_repack_nchw_to_nhwc(const float* src, float* dst, int32_t C, int32_t H, int32_t W,
                            lrf_stream_t stream) {
  if (!src || !dst || C < 1 || H < 1 || W < 1) return fail(LRF_ERR_INVALID, "bad repack arguments");
  cudaError_t e = lrf::launch_repack(src, dst, C, (long long)H * W, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "repack_kernel");
  return LRF_OK;
}

int lrf_launch_info(int32_t* n_sms, int32_t* threads_per_cta, int32_t* smem_bytes_per_cta) {
  DevInfo d;
  int rc = device_info(d);
  if (rc != LRF_OK) return rc;
  if (n_sms) *n_sms = d.n_sms;
  if (threads_per_cta) *threads_per_cta = lrf::render_threads();
  if (smem_bytes_per_cta) *smem_bytes_per_cta = (int32_t)lrf::render_smem_bytes(344, false, d.max_smem, false);
  return LRF_OK;
}

}  // extern "C"
