/*
 * lrf_oracle.h -- CPU restatement of localrf's per-ray-batch volume-rendering path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; localrf_b200/ never does.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py) against golden
 * vectors produced by running the unmodified reference (facebookresearch/localrf @ 3905e39,
 * /root/reference/localTensoRF) on CPU in the build container; see tests/golden/make_golden.py.
 *
 * All tensors use the REFERENCE layouts (host pointers, fp32):
 *   plane_i : [C][H_i][W_i]  with W_i = grid[matMode[i][0]], H_i = grid[matMode[i][1]]
 *   line_i  : [C][L_i]       with L_i = grid[vecMode[i]]
 *   matMode = {{0,1},{0,2},{1,2}}, vecMode = {2,1,0}     (models/tensorBase.py:274-275,
 *                                                        models/tensoRF.py:29-50)
 */
#ifndef LRF_ORACLE_H
#define LRF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcField {
  int32_t grid[3];           /* gridSize (x, y, z)                        tensorBase.py:322 */
  float aabb[6];             /* aabb[0] (min xyz), aabb[1] (max xyz)      tensorBase.py:259 */
  int32_t n_dcomp[3];        /* density_n_comp                            tensorBase.py:256 */
  int32_t n_acomp[3];        /* appearance_n_comp                         tensorBase.py:257 */
  const float *dplane[3], *dline[3], *aplane[3], *aline[3];
  int32_t app_dim;           /* basis_mat: Linear(sum n_acomp -> app_dim, bias=False) tensoRF.py:25 */
  const float *basis;        /* [app_dim][sum n_acomp] (nn.Linear.weight)               */
  int32_t featureC, fea_pe, view_pe;
  const float *w1, *b1;      /* mlp[0]      : [featureC][app_dim*(1+2*fea_pe)], [featureC] */
  const float *w2, *b2;      /* mlp[2]      : [featureC][featureC], [featureC]             */
  const float *w3, *b3;      /* mlp_view[0] : [3][featureC+3*(1+2*view_pe)], [3]           */
  const float *alpha_vol;    /* AlphaGridMask.alpha_volume [D][H][W] or NULL  tensorBase.py:38-62 */
  int32_t alpha_dims[3];     /* D, H, W */
  float alpha_aabb[6];
  float density_shift, distance_scale, weight_thres;
  int32_t act;               /* 0 = softplus, 1 = relu                   tensorBase.py:495-499 */
} OrcField;

/* ray_utils.py:9-12 -- in place on xyz[n][3] */
void orc_contract(float *xyz, int64_t n);

/* tensorBase.py:419-437 -- z[2*N], N = nSamples/6; j1/j2 = the two successive rand([1,N]) draws
 * (NULL for eval).  Returns S = 2*N. */
int32_t orc_sample_table(int32_t nSamples, const float *j1, const float *j2, float *z);

/* tensoRF.py:112-151 -- xyz_norm[M][3] in [-1,1]^3 -> out[M] */
void orc_density_feature(const OrcField *f, const float *xyz_norm, int64_t M, float *out);
/* tensoRF.py:153-196 -- xyz_norm[M][3] -> out[M][app_dim] */
void orc_app_feature(const OrcField *f, const float *xyz_norm, int64_t M, float *out);
/* tensorBase.py:97-135 (+14-21) -- feat[M][app_dim], viewdirs[M][3] -> rgb[M][3] */
void orc_mlp_late_view(const OrcField *f, const float *feat, const float *viewdirs, int64_t M,
                       int refine, float *rgb);
/* tensorBase.py:51-58 -- xyz[M][3] (un-normalised, contracted space) -> alpha values [M] */
void orc_alpha_mask_sample(const OrcField *f, const float *xyz, int64_t M, float *out);

/* TensorBase.forward, tensorBase.py:567-636.  rays[N][6]; z[S]; outputs rgb[N][3], depth[N];
 * optional (may be NULL) weights[N][S] (final weights, after the floater filter), acc[N],
 * n_app[N] (number of samples with weight > thres).  white_bg: 0/1 (the caller resolves the
 * train-mode coin flip of :633).  n_threads <= 0: all cores. */
void orc_field_forward(const OrcField *f, const float *rays, int64_t N, const float *z, int32_t S,
                       int white_bg, float floater_thresh, int refine, float *rgb, float *depth,
                       float *weights, float *acc, int32_t *n_app, int n_threads);

/* ray generation helpers (ray_utils.py:14-53, local_tensorfs.py:23-29, utils/utils.py:381-388) */
void orc_sixD_to_mtx(const float *r6 /*[V][3][2]*/, int64_t V, float *R /*[V][3][3]*/);
void orc_ray_directions(const int64_t *ray_ids, int64_t N, int32_t W, int32_t H, int fov360,
                        float focal, float cx, float cy, float *dirs /*[N][3]*/,
                        int64_t *ij /*[N][2] or NULL*/);

/* LocalTensorfs.forward, local_tensorfs.py:382-499 (eval or train; blending rows, cam2world and
 * exposure are resolved by the caller exactly as :403-416 and :481-495 do).
 *   fields[n_fields], zs[n_fields] (per-field sample tables), Ss[n_fields]
 *   cam2world [V][3][4]; world2rf [n_fields][3]; blend [V][n_fields]; exposure [V][3][3] or NULL
 *   rays are grouped by view: ray r belongs to view r / (N / V)        (:437-438)
 * Fields with an all-zero blend column are skipped (:418). */
void orc_local_forward(const OrcField *fields, int32_t n_fields, const float *const *zs,
                       const int32_t *Ss, const int64_t *ray_ids, int64_t N, int32_t W, int32_t H,
                       int fov360, float focal, float cx, float cy, const float *cam2world,
                       int64_t V, const float *world2rf, const float *blend, const float *exposure,
                       int white_bg, float floater_thresh, int refine, float *rgb, float *depth,
                       float *dirs, int n_threads);
/* same, plus margin [N] (or NULL): each ray's smallest |w - rayMarch_weight_thres| over the samples of
 * all active fields (test bookkeeping for exact ties on the hard switch of tensorBase.py:622) */
void orc_local_forward_m(const OrcField *fields, int32_t n_fields, const float *const *zs,
                         const int32_t *Ss, const int64_t *ray_ids, int64_t N, int32_t W, int32_t H,
                         int fov360, float focal, float cx, float cy, const float *cam2world,
                         int64_t V, const float *world2rf, const float *blend, const float *exposure,
                         int white_bg, float floater_thresh, int refine, float *rgb, float *depth,
                         float *dirs, float *margin, int n_threads);

/* Gradients of TensorBase.forward (what torch autograd computes through tensorBase.py:567-636 with
 * floater_thresh = 0): given dL/d(rgb_map) [N][3] and dL/d(depth_map) [N], accumulates dL/d(parameter)
 * into the OrcGrads buffers (reference layouts, caller zero-initialises) and writes dL/d(rays) [N][6].
 * Pinned against the reference's own autograd (tests/golden/grads_field.npz).  The checker for the
 * fused backward kernel (SURVEY.md par. 8f rank 1). */
typedef struct OrcGrads {
  float *dplane[3], *dline[3], *aplane[3], *aline[3];
  float *basis, *w1, *b1, *w2, *b2, *w3, *b3;
} OrcGrads;

void orc_field_backward(const OrcField *f, const float *rays, int64_t N, const float *z, int32_t S,
                        int white_bg, const float *g_rgb, const float *g_depth, OrcGrads *grads,
                        float *d_rays);

#ifdef __cplusplus
}
#endif
#endif
