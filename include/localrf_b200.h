/*
 * localrf_b200.h -- C ABI of the B200-native render path of localrf.
 *
 * This is the drop-in boundary for the per-ray-batch volume-rendering hot path of
 * facebookresearch/localrf (reference at localTensoRF/, commit 3905e39):
 *
 *     LocalTensorfs.forward      local_tensorfs.py:382-499   (ray generation, per-field render,
 *                                                             blend, exposure, clamp)
 *     TensorBase.forward         models/tensorBase.py:567-636 (sample, contract, density,
 *                                                             alpha/T/weights, floater filter,
 *                                                             appearance, MLP, composite)
 *     compute_densityfeature     models/tensoRF.py:112-151
 *     compute_appfeature         models/tensoRF.py:153-196
 *
 * The reference has no FFI for this path: its "plugin API" is the two nn.Module classes.  The
 * Python mirror of those classes (localrf_b200/tensorf.py, localrf_b200/local_tensorfs.py) binds
 * these entry points through ctypes; INTEGRATION.md shows the stub a maintainer of the reference
 * would add.  Plain pointers and sizes only -- no torch types.
 *
 * All pointers are DEVICE pointers on the current CUDA device unless stated otherwise; all
 * floating-point data is fp32 (the plane / line grids optionally bfloat16, LrfField.grid_dtype); work is
 * enqueued on `stream` and never synchronises the host.
 * Every function returns LRF_OK (0) or a negative error code; lrf_last_error() gives the text.
 *
 * HBM layout of a field ("channel-last", one texel's components contiguous):
 *     plane i : [H_i][W_i][C]     W_i = grid[matMode[i][0]], H_i = grid[matMode[i][1]]
 *     line  i : [L_i][C]          L_i = grid[vecMode[i]]
 *     matMode = {{0,1},{0,2},{1,2}}, vecMode = {2,1,0}          (models/tensorBase.py:274-275)
 * This is exactly the memory of the reference's [1,C,H,W] / [1,C,L,1] parameters when they are
 * allocated with torch.channels_last; lrf_repack_nchw_to_nhwc() converts a contiguous NCHW tensor.
 * C = 8 (density) and 24 (appearance) per plane are the reference defaults (opt.py:117-119) and
 * the only component counts built so far; app_dim = 27, featureC = 128; positional encodings
 * fea_pe / view_pe 0..8 (0, 0 is the reference default and the fastest path; lrf_render only).
 */
#ifndef LOCALRF_B200_H
#define LOCALRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRF_OK 0
#define LRF_ERR_INVALID -1      /* bad argument (null pointer, size, unsupported configuration) */
#define LRF_ERR_UNSUPPORTED -2  /* valid reference configuration that this build does not cover */
#define LRF_ERR_CUDA -3         /* a CUDA runtime call failed */

typedef struct CUstream_st *lrf_stream_t; /* == cudaStream_t */

/* One radiance field (TensorVMSplit): the parameters TensorBase.forward reads.
 * replaces: TensorVMSplit attributes, models/tensorBase.py:232-287, models/tensoRF.py:18-50 */
typedef struct LrfField {
  int32_t grid[3];            /* gridSize (x,y,z)                       tensorBase.py:322 */
  float aabb[6];              /* aabb[0] = min xyz, aabb[1] = max xyz   tensorBase.py:259 */
  int32_t n_dcomp;            /* density components per plane (8)       tensorBase.py:256 */
  int32_t n_acomp;            /* appearance components per plane (24)   tensorBase.py:257 */
  const float *dplane[3];     /* [H_i][W_i][n_dcomp] */
  const float *dline[3];      /* [L_i][n_dcomp] */
  const float *aplane[3];     /* [H_i][W_i][n_acomp] */
  const float *aline[3];      /* [L_i][n_acomp] */
  int32_t app_dim;            /* 27 */
  const float *basis;         /* basis_mat.weight [app_dim][3*n_acomp]  tensoRF.py:25-27 */
  int32_t featureC;           /* 128 */
  int32_t fea_pe, view_pe;    /* positional encodings of the MLP inputs (tensorBase.py:14-21,115-125); 0..8 */
  const float *w1, *b1;       /* renderModule.mlp[0]      [featureC][app_dim (1 + 2 fea_pe)], [featureC] */
  const float *w2, *b2;       /* renderModule.mlp[2]      [featureC][featureC], [featureC] */
  const float *w3, *b3;       /* renderModule.mlp_view[0] [3][featureC + 3 (1 + 2 view_pe)], [3] */
  const float *alpha_vol;     /* AlphaGridMask.alpha_volume [D][H][W], or NULL  tensorBase.py:38-62 */
  int32_t alpha_dims[3];      /* D, H, W */
  float alpha_aabb[6];
  float density_shift;        /* tensorBase.py:263 */
  float distance_scale;       /* tensorBase.py:265 */
  float weight_thres;         /* rayMarch_weight_thres, tensorBase.py:266 */
  int32_t act;                /* 0 = softplus, 1 = relu                 tensorBase.py:495-499 */
  const float *z_vals;        /* [n_samples] per-batch distance table (tensorBase.py:419-437);
                                 the host builds it (it owns the RNG of the train-mode jitter) */
  int32_t n_samples;          /* S = 2*(nSamples/6) */
  int32_t grid_dtype;         /* storage type of the twelve plane / line tensors: LRF_GRID_F32 (0, default) or
                                 LRF_GRID_BF16: the dplane/dline/aplane/aline pointers then address bfloat16
                                 data in the SAME [H][W][C] / [L][C] layout (a density texel = one 16-byte
                                 load, an appearance texel three); arithmetic stays fp32 and the conversion
                                 is exact, so a field whose parameters are bf16-representable renders the same
                                 values from either storage.  Inference only: accepted by lrf_render (pe = 0),
                                 lrf_density_feature and lrf_app_feature; every other entry point returns
                                 LRF_ERR_UNSUPPORTED.  lrf_pack_bf16() makes the 16-bit copy. */
} LrfField;
#define LRF_GRID_F32 0
#define LRF_GRID_BF16 1

/* One ray batch.  Either explicit rays (TensorBase.forward's rays_chunk) or ray ids + cameras
 * (LocalTensorfs.forward; the kernel then generates origin/direction itself).
 * replaces: local_tensorfs.py:397-456 (ids2pixel, get_ray_directions_*, cam2rf, get_rays_lean) */
typedef struct LrfBatch {
  int64_t n_rays;
  const float *rays;          /* [n_rays][6] (o, d) or NULL -> generate from the fields below */
  const int64_t *ray_ids;     /* [n_rays]  col = id % W, row = (id / W) % H   local_tensorfs.py:23-29 */
  int32_t W, H;
  int32_t fov360;             /* 1: get_ray_directions_360, 0: get_ray_directions_lean */
  float focal, cx, cy;        /* LocalTensorfs.focal(W), .center(W,H)   local_tensorfs.py:377-380 */
  const float *intrinsics;    /* optional DEVICE [3] = {focal, cx, cy}; overrides the three values
                                 above so learnable intrinsics never round-trip through the host */
  const float *cam2world;     /* [n_views][3][4] */
  int64_t n_views;            /* ray r uses view r / (n_rays / n_views)  local_tensorfs.py:437 */
  const float *world2rf;      /* DEVICE [3] added to the camera translation (NULL = 0)  local_tensorfs.py:428 */
  const float *blend;         /* per-view blending weight of THIS field: blend[v*blend_stride], or
                                 NULL for weight 1                       local_tensorfs.py:438,452 */
  int64_t blend_stride;
  const float *exposure;      /* [n_views][3][3] or NULL; applied when `finalize` local_tensorfs.py:481-496 */
  int32_t accumulate;         /* 0: rgb/depth = w*result; 1: rgb/depth += w*result  (:467-474) */
  int32_t finalize;           /* 1: apply exposure (if any) and clamp(0,1) to rgb   (:481-497) */
  int32_t white_bg;           /* tensorBase.py:633-634 (the caller resolves the train-mode coin) */
  float floater_thresh;       /* tensorBase.py:617-620 */
  int32_t refine;             /* MLPRender_Fea_late_view's `refine` (tensorBase.py:118-124; LocalTensorfs passes
                                 is_refining, local_tensorfs.py:463): 0 = zeros instead of the feature encoding.
                                 Only read when fea_pe > 0. */
} LrfBatch;

typedef struct LrfOutputs {
  float *rgb;                 /* [n_rays][3] */
  float *depth;               /* [n_rays] */
  float *weights;             /* optional [n_rays][n_samples]: final per-sample weights of this field */
  float *directions;          /* optional [n_rays][3]: camera-space directions (ray-generation mode) */
  int64_t *ij;                /* optional [n_rays][2]: pixel (col, row) of every ray id (ray-generation mode) */
  float *pix;                 /* optional [n_rays][4]: when set, rgb and depth are written interleaved
                                 here (r,g,b,depth) instead of to rgb/depth -- the layout the multi-GPU
                                 path all-gathers in one collective */
  unsigned long long *stats;  /* optional [2]: += {density samples marched, appearance samples shaded} */
  /* ---- fused pixel exchange (multi-GPU): when n_peers > 0 the thread that finishes a
   * ray on the `finalize` launch ALSO stores its (r,g,b,depth) as one 16-byte word into every
   * peer_pix[p] + 4*ray -- peer-mapped device memory of the other GPUs (NVLink P2P / symmetric
   * memory), each pointer already offset to where THIS rank's rays live in that peer's gathered
   * [total_rays][4] buffer -- so the all-gather of rendered pixels happens inside the render kernel,
   * store by store, instead of as a collective after it.  `pix` must be set (local copy).  If
   * mc_pix is set (NVLS multicast mapping of the same buffer) ONE multimem store reaches all peers
   * and peer_pix is ignored.  lrf_peer_barrier() then closes the step. */
  int32_t n_peers;            /* 0 = no exchange; <= LRF_MAX_PEERS */
  float *peer_pix[16];
  float *mc_pix;
  /* In-kernel step signalling (optional, signal_seq > 0): peer_flags[p] = peer p's [n_peers] uint64 flag
   * array (peer-mapped; peer_flags[rank] is this GPU's own).  Before its first peer store the launch waits
   * until all local flags have reached wait_seq (the peers are done with the buffer about to be overwritten;
   * 0 = no wait); the LAST CTA to finish release-stores signal_seq into slot `rank` of every peer's array,
   * after all pixel stores of the launch.  With wait_seq = signal_seq - lag (lag >= 1) this replaces
   * lrf_peer_signal_wait altogether: a step costs no extra launch. */
  unsigned long long *peer_flags[16];
  int32_t rank;
  unsigned long long signal_seq, wait_seq;
} LrfOutputs;
#define LRF_MAX_PEERS 16

int lrf_version(void);
const char *lrf_last_error(void);
/* sizeof() of the ABI structs as this library was compiled: 0 LrfField, 1 LrfBatch, 2 LrfOutputs,
 * 3 LrfGradients (0 for any other index).  Bindings compare these with their own mirrors at load
 * time, so a stale library or a drifted mirror fails loudly instead of mis-reading arguments. */
size_t lrf_sizeof(int32_t which);

/* Device bytes of the per-field "prepared" block (folded / re-laid-out MLP weights) of a field without
 * positional encodings; lrf_prepared_bytes_for() covers every field (fea_pe / view_pe in [0, 8]; 0 if not
 * built).  A field with encodings needs the block 1024-byte aligned (its layer-1 operand is streamed by TMA). */
size_t lrf_prepared_bytes(void);
size_t lrf_prepared_bytes_for(const LrfField *field);
/* Builds the prepared block from the field's current MLP / basis weights.  Call again whenever
 * those parameters change (every optimiser step while training). */
int lrf_field_prepare(const LrfField *field, void *prepared, lrf_stream_t stream);

/* The hot path: one fused launch per (field, ray batch).
 * replaces: TensorBase.forward (tensorBase.py:567-636) + the per-field body of
 * LocalTensorfs.forward's chunk loop (local_tensorfs.py:451-474) + :481-497 when finalize. */
int lrf_render(const LrfField *field, const void *prepared, const LrfBatch *batch,
               const LrfOutputs *out, lrf_stream_t stream);

/* basis_mat + MLPRender_Fea_late_view.forward (tensoRF.py:196 + tensorBase.py:115-135, pe = 0) on
 * explicit inputs: feats [M][72] = the plane x line products (tensoRF.py:192-194 order: plane-major,
 * channel-minor), viewdirs [M][3] normalised -> rgb [M][3].  Same tensor-core code as lrf_render. */
int lrf_mlp_forward(const void *prepared, const float *feats, const float *viewdirs, int64_t M,
                    float *rgb, lrf_stream_t stream);

/* compute_densityfeature (tensoRF.py:112-151): xyz_norm [M][3] in [-1,1]^3 -> out [M] */
int lrf_density_feature(const LrfField *field, const float *xyz_norm, int64_t M, float *out,
                        lrf_stream_t stream);
/* compute_appfeature (tensoRF.py:153-196): xyz_norm [M][3] -> out [M][app_dim] */
int lrf_app_feature(const LrfField *field, const float *xyz_norm, int64_t M, float *out,
                    lrf_stream_t stream);

/* ---- differentiable lookups (the composed training path) -------------------
 * The 72 plane x line products of compute_appfeature BEFORE basis_mat (tensoRF.py:174-194 order):
 * xyz_norm [M][3] -> out [M][3*n_acomp]. */
int lrf_app_products(const LrfField *field, const float *xyz_norm, int64_t M, float *out,
                     lrf_stream_t stream);
/* Backward of lrf_density_feature (autograd of F.grid_sample x6 + product-sum, tensoRF.py:112-151):
 * grad_out [M]; d_plane[i] / d_line[i] have the layout of the parameters and are ACCUMULATED into
 * (atomic adds; zero them first); d_xyz [M][3] is overwritten (may be NULL). */
int lrf_density_feature_backward(const LrfField *field, const float *xyz_norm,
                                 const float *grad_out, int64_t M, float *const d_plane[3],
                                 float *const d_line[3], float *d_xyz, lrf_stream_t stream);
/* Backward of lrf_app_products: grad_out [M][3*n_acomp]; same conventions. */
int lrf_app_products_backward(const LrfField *field, const float *xyz_norm, const float *grad_out,
                              int64_t M, float *const d_plane[3], float *const d_line[3],
                              float *d_xyz, lrf_stream_t stream);

/* ---- backward of the whole render -----------------------------------------
 * What torch autograd computes through TensorBase.forward (tensorBase.py:567-636, the training path
 * local_tensorfs.py:417-452 takes per field) with floater_thresh = 0: given dL/d(rgb_map) [n][3] and
 * dL/d(depth_map) [n] for a batch of explicit rays, ACCUMULATES dL/d(parameter) into buffers laid
 * out like the parameters (zero them first) and overwrites d_rays [n][6].
 *   d_w1b is dL/d(mlp[0].weight @ basis_mat.weight) [featureC][3*n_acomp]; the caller finishes
 *   dL/d(mlp[0].weight) = d_w1b @ basis^T and dL/d(basis_mat.weight) = mlp[0].weight^T @ d_w1b.
 * `prepared_bwd`: lrf_prepared_backward_bytes() bytes filled by lrf_field_prepare_backward (redo after
 * a parameter update).  `scratch`: lrf_backward_scratch_bytes(n_rays, n_samples) bytes of device
 * memory, contents irrelevant.  white_bg as in LrfBatch.  PE > 0 is LRF_ERR_UNSUPPORTED. */
typedef struct LrfGradients {
  float *d_rays;                                /* [n][6]                                  */
  float *d_dplane[3], *d_dline[3];              /* [H][W][n_dcomp], [L][n_dcomp]           */
  float *d_aplane[3], *d_aline[3];              /* [H][W][n_acomp], [L][n_acomp]           */
  float *d_w1b;                                 /* [featureC][3*n_acomp]                   */
  float *d_b1, *d_w2, *d_b2, *d_w3, *d_b3;      /* shapes of mlp[0].bias .. mlp_view[0].bias */
} LrfGradients;
size_t lrf_prepared_backward_bytes(void);
size_t lrf_backward_scratch_bytes(int64_t n_rays, int32_t n_samples);
int lrf_field_prepare_backward(const LrfField *field, void *prepared_bwd, lrf_stream_t stream);
int lrf_render_backward(const LrfField *field, const void *prepared_bwd, const float *rays,
                        int64_t n_rays, int32_t white_bg, const float *grad_rgb,
                        const float *grad_depth, const LrfGradients *grads, void *scratch,
                        size_t scratch_bytes, lrf_stream_t stream);

/* Closes a fused pixel exchange: enqueues on `stream` a one-CTA kernel that (1) makes this GPU's
 * earlier peer stores visible system-wide, (2) writes `seq` into slot `rank` of every peer's flag
 * array (peer_flags[p] = peer p's [world] uint64 array, peer-mapped; peer_flags[rank] is this GPU's
 * own), (3) waits until all `world` slots of the local array have reached `seq`.  When the kernel
 * ends, every rank's pixels of step `seq` are in this GPU's gathered buffer.  `seq` must increase by
 * one per step, starting at 1 (flags zero-initialised).  replaces: the ncclAllGather of SURVEY.md 8e. */
int lrf_peer_barrier(unsigned long long *const *peer_flags, int32_t rank, int32_t world,
                     unsigned long long seq, lrf_stream_t stream);
/* The same with the two halves decoupled: signals `seq`, waits until every peer has signalled `wait_seq`
 * (<= seq).  wait_seq = seq - 1 is the double-buffered exchange: step i's pixels are pushed while the
 * consumer works on step i-1's, so a step never waits for the slowest peer of the SAME step. */
int lrf_peer_signal_wait(unsigned long long *const *peer_flags, int32_t rank, int32_t world,
                         unsigned long long seq, unsigned long long wait_seq, lrf_stream_t stream);

/* ---- schedule-time operators on the field's tensors ---------------------------
 * Occupancy-mask rebuild.  replaces: getDenseAlpha + updateAlphaMask (models/tensorBase.py:501-536), which
 * move the model to the CPU and loop over slabs.  dims = (gx,gy,gz) of the lattice (the caller passes
 * gridSize/2, local_tensorfs.py:264-266); `length` = stepSize.  alpha_scratch [gx][gy][gz] receives
 * getDenseAlpha()'s alpha (an existing mask in `field` culls as compute_alpha does, :538-558); mask
 * [gz][gy][gx] receives {0,1} = max_pool3d(3, pad 1)(clamp(alpha,0,1)) >= thres; *kept += #ones.
 * mask may be NULL (dense alpha only). */
int lrf_alpha_mask_build(const LrfField *field, const int32_t dims[3], float length, float thres,
                         float *alpha_scratch, float *mask, unsigned long long *kept,
                         lrf_stream_t stream);
/* Grid upsampling.  replaces: F.interpolate(bilinear, align_corners=True) of up_sampling_VM
 * (models/tensoRF.py:198-221) on one channel-last tensor: src [H][W][C] -> dst [H2][W2][C] (lines: W = W2 = 1).
 * C must be a multiple of 4. */
int lrf_upsample(const float *src, int32_t H, int32_t W, float *dst, int32_t H2, int32_t W2, int32_t C,
                 lrf_stream_t stream);
/* density_L1 (models/tensoRF.py:83-92) without the 8*G^3 intermediate: *sum += sum_n sqrt(clamp(
 * feature2density(f[n]), 1e-5)) over the G^3 flat indices (the caller divides by G^3).  The backward
 * ACCUMULATES d(mean)/d(plane_i, line_i) * (*grad_out) into buffers laid out like the parameters. */
int lrf_density_l1(const LrfField *field, double *sum, lrf_stream_t stream);
int lrf_density_l1_backward(const LrfField *field, const float *grad_out, float *const d_plane[3],
                            float *const d_line[3], lrf_stream_t stream);
/* TVLoss (utils/utils.py:293-312) pieces on one channel-last [H][W][C] tensor: sums[0] += sum of squared
 * differences along H, sums[1] += along W.  Backward: dx += (*grad_out) * (kh * d sums[0]/dx + kw * d sums[1]/dx). */
int lrf_tv_sums(const float *x, int32_t H, int32_t W, int32_t C, double *sums, lrf_stream_t stream);
int lrf_tv_sums_backward(const float *x, int32_t H, int32_t W, int32_t C, const float *grad_out, float kh,
                         float kw, float *dx, lrf_stream_t stream);
/* sample_ray (models/tensorBase.py:396-417): ray-AABB entry distance clamped to [near, far], S uniform steps
 * of `step` from there (+ jitter[r] steps when jitter != NULL: the train-mode draw, one per ray), points and
 * the inside-the-box mask.  rays [N][6]; aabb = {min xyz, max xyz}; pts [N][S][3], z [N][S], inside [N][S]. */
int lrf_sample_ray(const float *rays, const float *jitter, int64_t N, int32_t S, const float aabb[6],
                   float near, float far, float step, float *pts, float *z, unsigned char *inside,
                   lrf_stream_t stream);

/* Frame post-processing on the device.  replaces: the host-side conversions of renderer.py:126-131,173-176
 * (rgb_map.cpu() ... cv2.imwrite(255 * rgb[..., ::-1]); visualize_depth -> cv2.applyColorMap, utils/utils.py:
 * 179-197).  rgb [N] x 3 floats `rgb_stride` apart, depth [N] floats `depth_stride` apart (3 / 1, or 4 / 4 for the
 * interleaved pix layout) -> rgb8 [N][3] 8-bit BGR and depth8 [N][3] = lut[(uint8)(255 * clip((d - d_lo) /
 * (d_hi - d_lo + 1e-8), 0, 1))] with lut [256][3] (the caller's colour map, e.g. cv2.COLORMAP_JET).  Either
 * output may be NULL.  Outputs may live in pinned host memory (zero-copy). */
int lrf_frame_to_u8(const float *rgb, int32_t rgb_stride, const float *depth, int32_t depth_stride, int64_t N,
                    float d_lo, float d_hi, const unsigned char *lut, unsigned char *rgb8, unsigned char *depth8,
                    lrf_stream_t stream);

/* fp32 -> bfloat16 (round to nearest even), n elements: the 16-bit copy of a channel-last plane / line
 * tensor for LrfField.grid_dtype = LRF_GRID_BF16.  dst must be 16-byte aligned when used as a grid. */
int lrf_pack_bf16(const float *src, void *dst, int64_t n, lrf_stream_t stream);

/* [C][H][W] (contiguous NCHW parameter of the reference) -> [H][W][C] */
int lrf_repack_nchw_to_nhwc(const float *src, float *dst, int32_t C, int32_t H, int32_t W,
                            lrf_stream_t stream);

/* Launch configuration used by lrf_render on this device (for the bench's roofline record). */
int lrf_launch_info(int32_t *n_sms, int32_t *threads_per_cta, int32_t *smem_bytes_per_cta);

#ifdef __cplusplus
}
#endif
#endif /* LOCALRF_B200_H */
