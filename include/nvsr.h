/*
 * nvsr.h — C-ABI of libnvsr_b200.so: the B200 (sm_100a) ray-rendering hot path of
 * Neural-Volume-Super-Resolution.
 *
 * The reference has no FFI / plugin registry: its seams are module-level Python functions
 * (SURVEY.md §8b).  Every entry point below therefore names the reference *Python* interface it
 * replaces (file:line under the reference checkout) — the Python host side in
 * `neural-volume-super-resolution_b200/` binds these with ctypes and mirrors the reference signatures.
 *
 * Conventions
 *   - every function returns int32 status: 0 = OK, <0 = NVSR_ERR_*, >0 = cudaError_t.
 *   - no exceptions, no allocation, no global state: every buffer is a caller-owned DEVICE pointer
 *     (except where a parameter is documented "host"), every call takes the caller's cudaStream_t
 *     (as void*) and is asynchronous on that stream.
 *   - all floating-point tensors are fp32 and row-major unless stated; index tensors are int64.
 *   - "tile image" = the bf16 feature layout the tcgen05 decoder consumes with one bulk copy per
 *     128-row tile:  [tile][K/8][128 rows][8 bf16]   (UMMA K-major, no-swizzle canonical layout).
 */
#ifndef NVSR_H_
#define NVSR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVSR_ABI_VERSION 4

#define NVSR_OK 0
#define NVSR_ERR_INVALID_ARG (-1)
#define NVSR_ERR_UNSUPPORTED (-2)
#define NVSR_ERR_ALIGNMENT (-3)
#define NVSR_ERR_RESOURCE (-4)

#define NVSR_F32 0
#define NVSR_BF16 1
#define NVSR_F16 2  /* IEEE half: same tensor-core rate as bf16, 8x finer rounding, max 65504 (saturating) */

#define NVSR_TILE_ROWS 128

/* Row order of a (n_rays x S samples) point set inside feature tile images and the planar raw buffer.
 *   RAY_MAJOR : row = ray*S + s                                (the reference's flatten order)
 *   BLOCKED   : a 128-row tile holds NVSR_BLK_RAYS consecutive rays x NVSR_BLK_SAMPLES consecutive
 *               samples:  tile = (ray/8)*ceil(S/16) + s/16,  row = tile*128 + (s%16)*8 + ray%8.
 *               Consecutive rows are adjacent pixels at the same depth, so the texel reads of a warp
 *               coalesce in L1, and the 8 rays of a block own one contiguous span of the raw buffer.
 *               Rows of padding rays/samples exist in the buffers and are ignored.
 * Buffers hold nvsr_rows_padded() rows. */
#define NVSR_ROWS_RAY_MAJOR 0
#define NVSR_ROWS_BLOCKED 1
#define NVSR_BLK_RAYS 8
#define NVSR_BLK_SAMPLES 16
int64_t nvsr_rows_padded(int64_t n_rays, int32_t n_samples, int32_t row_order);
#define NVSR_MAX_LAYERS 8
#define NVSR_MAX_SAMPLES 1024 /* samples per ray handled by the warp-per-ray kernels */

int32_t nvsr_abi_version(void);
const char* nvsr_status_string(int32_t status);

/* ------------------------------------------------------------------------------------------------
 * a1  get_ray_bundle            nerf_helpers.py:507-549 (+ meshgrid_xy :396-406, get_focal :432-437)
 * Rays of image rows [row_begin,row_end) of the (H+2p)x(W+2p) pixel grid, written as
 * ro, rd : [(row_end-row_begin), W+2p, 3].  Ray index = r*W + c (row-major) — the ordering to keep.
 * focal_x divides the x term, focal_y the y term (the reference's get_focal quirk is resolved by
 * the Python caller).  c2w_host: 16 floats, row-major 4x4, HOST memory (copied by value).
 */
int32_t nvsr_ray_bundle(int32_t height, int32_t width, float focal_x, float focal_y,
                        const float* c2w_host, int32_t padding, float offset, int32_t row_begin,
                        int32_t row_end, float* ro, float* rd, void* stream);
/* Same, with the pose in DEVICE memory (row-major 4x4 fp32, read by the kernel): no device->host copy of a pose
 * the caller keeps on the GPU (the reference does, train_nerf.py:659), hence no synchronisation per frame. */
int32_t nvsr_ray_bundle_dev(int32_t height, int32_t width, float focal_x, float focal_y,
                            const float* c2w_device, int32_t padding, float offset, int32_t row_begin,
                            int32_t row_end, float* ro, float* rd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a2/a3  ray preparation of run_one_iter_of_nerf      train_utils.py:210-226, ndc_rays
 * nerf_helpers.py:578-605.  viewdirs = rd/|rd| (computed from the pre-NDC directions); when
 * use_ndc != 0, (ro,rd) are mapped to NDC with near plane `ndc_near` (the call site passes 1.0).
 * Any output may alias nothing; viewdirs may be NULL.
 */
int32_t nvsr_prepare_rays(const float* ro_in, const float* rd_in, int64_t n_rays, int32_t use_ndc,
                          int32_t height, int32_t width, double focal, double ndc_near, float* ro_out,
                          float* rd_out, float* viewdirs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * plane re-pack (once per scene / per SR inference): reference planes are NCHW fp32 [1,C,Rh,Rw]
 * (models.py:436-439).  dst_dtype NVSR_F32: channels-last [Rh][Rw][C] fp32 (parity-mode gather, view
 * plane).  NVSR_BF16 | NVSR_F16: "x-pair records" [Rh][C/8][Rw][2][8] 16-bit, C % 8 == 0 — the 32-byte
 * record (y*(C/8) + c)*Rw + x holds the 8-channel chunk c of texel (y,x) followed by the same chunk of
 * texel (y, min(x+1, Rw-1)): both x corners of a bilinear footprint row arrive with one 256-bit load and
 * neighbouring rays share 128-byte lines (2x the plane bytes: 7.7 MB per 200^2 plane, 123 MB per 800^2).
 * Values are saturated to the 16-bit format's finite range.  dst must be 32-byte aligned for these.
 */
int32_t nvsr_pack_plane(const float* src_nchw, int32_t channels, int32_t rh, int32_t rw, void* dst,
                        int32_t dst_dtype, void* stream);

/* nn.Linear weight [n_out, k] (row stride ldw, fp32) -> 16-bit (NVSR_BF16|NVSR_F16) UMMA image
 * [k_pad/8][n_out][8], zero-padded for k <= kk < k_pad.  k_pad % 16 == 0. */
int32_t nvsr_pack_weight16(const float* w, int32_t n_out, int32_t k, int32_t ldw, int32_t k_pad,
                           void* dst, int32_t dst_dtype, void* stream);
/* `count` weights in one call and one launch per 8 (same arguments as nvsr_pack_weight16, as arrays).  absmax (device
 * fp32, may be NULL): max(*absmax, max |w| of everything packed) is left there (atomic on the bit pattern: start it at
 * 0; a NaN weight leaves a NaN) — the fp16 range check of a step that re-packs its weights, without a host read. */
int32_t nvsr_pack_weights16(int32_t count, const float* const* w, const int32_t* n_out, const int32_t* k, const int32_t* ldw,
                            const int32_t* k_pad, void* const* dst, int32_t dst_dtype, float* absmax, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a4 + a5  stratified sampler (train_utils.py:95-111) fused with the tri-plane bilinear gather of
 * TwoDimPlanesModel.forward (models.py:261-268 normalize_coords, :495-497 CoordProjector,
 * :289-310 project_xyz = F.grid_sample(bilinear, align_corners=True, padding_mode='border'),
 * :355-361 combine_pos_planes('avg')).
 */
typedef struct nvsr_planes {
  const void* plane[3];   /* nvsr_pack_plane images: fp32 [rh][rw][channels] | 16-bit [rh][channels/8][rw][2][8] */
  int32_t rh[3], rw[3];
  int32_t channels;       /* multiple of 8, <= 64 */
  int32_t dtype;          /* NVSR_F32 | NVSR_BF16 | NVSR_F16 */
  float box_lo[3];        /* fp32(box_coords[scene][0,:3]) */
  float box_rng[3];       /* fp32(box[1,:3] - box[0,:3]) with the difference taken in fp64 */
  float proj[3][6];       /* rot_mats[d][:,1:] row-major [3][2]: grid = n_xyz @ proj[d] */
  int32_t combine;        /* combine_pos_planes (models.py:355-361) for featM: 0 = 'avg' (sum / 3), 1 = 'sum' */
} nvsr_planes_t;

typedef struct nvsr_sampler {
  int64_t n_rays;
  int32_t n_samples;      /* S */
  const float* ro;        /* [n,3] */
  const float* rd;        /* [n,3] */
  float near_, far_;
  int32_t lindisp;
  const float* t_vals;    /* [S]   coarse pass: torch.linspace(0,1,S) made by the caller */
  const float* t_rand;    /* [n,S] caller-supplied uniforms (perturb) or NULL */
  const float* z_in;      /* [n,S] fine pass: merged depths; when non-NULL t_vals/t_rand unused */
} nvsr_sampler_t;

#define NVSR_FEAT_ROWMAJOR_F32 0 /* featP [rows,3C] fp32, featM [rows,C] fp32; rows RAY_MAJOR       */
#define NVSR_FEAT_TILE_BF16 1    /* featP [tiles][3C/8][128][8] bf16, featM [tiles][C/8][128][8]; BLOCKED */
#define NVSR_FEAT_TILE_F16 2     /* same tile image with fp16 elements (planes must be NVSR_F16)   */

/* feat_p may be NULL for the tile-image layouts (density features only: the sparse colour path gathers the
 * 3-plane features of the contributing rows afterwards, nvsr_sample_gather_rows).
 * Row-major output: rows = n_rays*n_samples, row = ray*S + s.  Tile-image outputs use the BLOCKED
 * row order and must be sized for nvsr_rows_padded(n,S,BLOCKED)/128 tiles; padding rows are written
 * as zeros.  z_out [n,S] (always ray-major) may be NULL. */
int32_t nvsr_sample_gather(const nvsr_sampler_t* sampler, const nvsr_planes_t* planes,
                           int32_t feat_layout, void* feat_p, void* feat_m, float* z_out,
                           void* stream);

/* a5 view-direction half: cart2az_el (nerf_helpers.py:492-496) + normalize + project_viewdir
 * (models.py:312-326) — hoisted per ray (the reference recomputes it per sample).
 * vplane: channels-last fp32 [rh][rw][channels].  vfeat: [n,channels] fp32.
 * (lo, rng) = fp32(box[0,3:5]), fp32(box[1,3:5]-box[0,3:5]) as in nvsr_planes_t. */
int32_t nvsr_viewdir_gather(const float* viewdirs, int64_t n_rays, const float* vplane, int32_t rh,
                            int32_t rw, int32_t channels, float az_lo, float az_rng, float el_lo,
                            float el_rng, float* vfeat, void* stream);

/* per-ray bias of a decoder layer whose input is concat(per-sample, per-ray) features:
 * out[ray,n] = b[n] + sum_k w[n*ldw + k] * vin[ray,k]     (the per-ray columns of rgb_dec[0],
 * models.py:186 / layers_dir[0], models.py:68) */
int32_t nvsr_row_bias(const float* vin, int64_t n_rays, int32_t k, const float* w, int32_t ldw,
                      const float* b, int32_t n_out, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse colour path (exact): a sample with sigma (+ noise) <= 0 has alpha = 1 - exp(-relu(sigma) * dist) = 0, so
 * its weight alpha * T is exactly 0 and its colour cannot reach rgb_map / depth / acc
 * (volume_rendering_utils.py:29-44).  After the density chain, nvsr_keep_rows lists the BLOCKED row ids of the
 * samples that can contribute (sigma + noise > 0, or NaN) — `count` must be zeroed by the caller, the order of the
 * list is unspecified — and nvsr_sample_gather_rows writes the 3-plane features (featP) of exactly those rows into
 * densely packed tile images in list order (entry i -> tile i/128, row i%128; the last tile is zero-padded),
 * ready for nvsr_mlp_chain with row_ids/row_count.  sigma: channel 3 of the planar raw buffer (BLOCKED order).
 * z_in of the sampler is required ([n,S] depths of the pass).  max_rows bounds *count (buffer capacity).
 */
int32_t nvsr_keep_rows(const float* sigma, const float* noise, int64_t n_rays, int32_t n_samples,
                       int32_t* keep_rows, int32_t* count, void* stream);
int32_t nvsr_sample_gather_rows(const nvsr_sampler_t* sampler, const nvsr_planes_t* planes, int32_t feat_layout,
                                const int32_t* keep_rows, const int32_t* count, int64_t max_rows, void* feat_p,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * a6 / a6'  decoder MLP as a chain of dense layers   models.py:168-197,393-421 (planes decoder),
 * models.py:14-108 (FlexibleNeRFModel).  One call evaluates one chain over `rows` rows and
 * writes its heads into planar raw[ch][raw_stride].
 */
typedef struct nvsr_layer {
  const void* w;         /* F32: [n_out,k] row-major;  BF16/F16: UMMA image [k/8][n_out][8] */
  const float* bias;     /* [n_out]; ignored when row_bias != NULL */
  const float* row_bias; /* [n_rays,n_out] per-ray bias (already includes bias) or NULL */
  const float* head_w;   /* [head_n,n_out] fp32 head tapped on this layer's output, or NULL */
  const float* head_b;   /* [head_n] */
  int32_t k, n_out, relu;
  int32_t head_n, head_ch; /* head writes raw channels head_ch .. head_ch+head_n-1 */
} nvsr_layer_t;

typedef struct nvsr_mlp {
  int32_t precision;     /* NVSR_F32: SIMT fp32 kernel, input row-major fp32 [rows,k0];
                            NVSR_BF16 | NVSR_F16: tcgen05 kernel, input = 16-bit tile image */
  int32_t n_layers;
  nvsr_layer_t layer[NVSR_MAX_LAYERS];
  const void* in;
  int64_t rows;          /* rows to evaluate (RAY_MAJOR: n_rays*S; BLOCKED: nvsr_rows_padded) */
  int32_t samples_per_ray; /* S: with row_order gives row -> ray (for row_bias) */
  int64_t n_rays;
  float* raw;            /* planar [4][raw_stride], same row order as the input */
  int64_t raw_stride;
  int32_t row_order;     /* NVSR_ROWS_* of the input rows (tcgen05 path; the fp32 path is RAY_MAJOR) */
  /* sparse evaluation (tcgen05 path, NVSR_ROWS_BLOCKED): input row i is the BLOCKED row row_ids[i] of the chunk —
   * the per-ray bias and the raw output position follow row_ids; *row_count (device) rows are evaluated, `rows`
   * is then the capacity of the input buffer.  NULL/NULL: every row is evaluated in place. */
  const int32_t* row_ids;
  const int32_t* row_count;
} nvsr_mlp_t;

int32_t nvsr_mlp_chain(const nvsr_mlp_t* mlp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a7 + a8  volume_render_radiance_field (volume_rendering_utils.py:6-51, cumprod_exclusive
 * nerf_helpers.py:409-430) and, on the coarse pass, sample_pdf_2 (nerf_helpers.py:668-702) with
 * the call-site prep and sort-merge of train_utils.py:144-156.  One warp per ray.
 */
typedef struct nvsr_composite {
  int64_t n_rays;
  int32_t n_samples;       /* S: radiance samples per ray */
  const float* raw;        /* planar [4][raw_stride]: r,g,b,sigma of row(ray,s) in `row_order` */
  int64_t raw_stride;
  int32_t row_order;       /* NVSR_ROWS_RAY_MAJOR | NVSR_ROWS_BLOCKED */
  const float* z;          /* [n,S]  (mip: [n,S+1] interval edges) */
  const float* rd;         /* [n,3] */
  const float* noise;      /* [n,S] noise already scaled by radiance_field_noise_std, or NULL */
  int32_t white_bkgd;
  int32_t mip;
  float* rgb;              /* [n,3] */
  float* disp;             /* [n] */
  float* acc;              /* [n] */
  float* depth;            /* [n] */
  float* weights;          /* [n,S] or NULL */
  /* hierarchical resampling; n_fine == 0 disables (fine pass) */
  int32_t n_fine;          /* samples to draw (mip: caller passes num_fine+1) */
  const float* u;          /* [n_fine] (u_per_ray==0, sorted/deterministic) or [n,n_fine] */
  int32_t u_per_ray;
  int64_t* inds;           /* [n,n_fine] searchsorted(right) indices or NULL */
  float* z_samples;        /* [n,n_fine] or NULL */
  float* z_merged;         /* [n, S(+1 if mip) + n_fine] sorted */
} nvsr_composite_t;

int32_t nvsr_composite(const nvsr_composite_t* args, void* stream);

/* a8 stand-alone: sample_pdf(bins[n,B], weights[n,B-1], num_samples, det) nerf_helpers.py:668.
 * cdf_in [n,B] (optional): skip the pdf/cdf construction and search this cdf instead (stage test:
 * identical cdf,u => bit-exact inds). */
int32_t nvsr_sample_pdf(const float* bins, const float* weights, const float* cdf_in, int64_t n_rays,
                        int32_t n_bins, const float* u, int32_t u_per_ray, int32_t n_samples,
                        int64_t* inds, float* samples, float* cdf_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a9  mip path: cast_rays + conical_frustum_to_gaussian + lift_gaussian (mip.py:9-43) fused with
 * IntegratedPositionalEncoding (mip.py:154-199).  z: [n,S+1] interval edges; out: [n*S, 6*n_freqs]
 * (row-major fp32, out_layout 0) or a 16-bit tile image padded to k_pad columns (out_layout 1|2);
 * rows are RAY_MAJOR in both.
 */
int32_t nvsr_ipe(const float* z, const float* ro, const float* rd, int64_t n_rays, int32_t n_intervals,
                 float radius, int32_t n_freqs, int32_t out_layout, int32_t k_pad, void* out,
                 void* stream);

/* a9 stage-level, with the reference's own tensors at the boundary (same-signature drop-ins):
 * cast_rays(t_vals, origins, directions, radii, ray_shape) mip.py:9-18 (+ conical_frustum_to_gaussian :21-29,
 * lift_gaussian :32-43): z [n,S+1] interval edges -> means, covs [n,S,3] (diagonal covariances).
 * radii: [n] per-ray base radius (device) or NULL, then the scalar `radius` (train_utils.py:21-24 builds a
 * constant column). */
int32_t nvsr_cast_rays(const float* z, const float* ro, const float* rd, const float* radii, float radius,
                       int64_t n_rays, int32_t n_intervals, float* means, float* covs, void* stream);
/* IntegratedPositionalEncoding.forward((means, covs)) mip.py:164-191: means, covs [rows,3] ->
 * out [rows, 6*n_freqs] = exp(-0.5*[y_var,y_var]) * sin([y, y+pi/2]), y = means x 2^i, i < n_freqs = multires-1. */
int32_t nvsr_ipe_encode(const float* means, const float* covs, int64_t rows, int32_t n_freqs, float* out,
                        void* stream);

/* positional_encoding(viewdirs, n_freqs, include_input) nerf_helpers.py:552-575, per ray. */
int32_t nvsr_dir_encoding(const float* dirs, int64_t n_rays, int32_t n_freqs, int32_t include_input,
                          float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convenience: the whole coarse -> fine pipeline of predict_and_render_radiance (train_utils.py:71-182) for one batch
 * of PREPARED rays (nvsr_prepare_rays) of a tri-plane scene, with one host call and no allocation: the stage entry
 * points above, in the order the reference runs them, out of a caller-owned workspace of nvsr_workspace_bytes() bytes
 * (256-byte aligned).  Every sample goes through both decoders (the sparse colour path is a host-side optimisation of
 * the Python shim).  Results are bit-identical to issuing the stage calls one by one.
 */
typedef struct nvsr_decoder {
  nvsr_layer_t density[NVSR_MAX_LAYERS];  /* density_dec + fc_alpha head on the last layer (models.py:393-404) */
  int32_t n_density;
  nvsr_layer_t rgb[NVSR_MAX_LAYERS];      /* rgb_dec over the 3C position columns + fc_rgb head; rgb[0].row_bias is set by the call */
  int32_t n_rgb;
  const float* view_w;                    /* the per-ray (view feature) columns of rgb_dec[0].weight: [n_out, C], row stride view_ldw */
  int32_t view_ldw;
  const float* view_b;                    /* rgb_dec[0].bias [n_out] */
} nvsr_decoder_t;

typedef struct nvsr_render {
  int32_t precision;                      /* NVSR_F32 | NVSR_BF16 | NVSR_F16: selects gather layout and decoder kernel */
  int64_t n_rays;
  int32_t n_coarse, n_fine;               /* n_fine == 0: coarse pass only */
  const float *ro, *rd, *viewdirs;        /* [n,3] each, as nvsr_prepare_rays returns them */
  float near_, far_;
  int32_t lindisp, white_bkgd;
  const float* t_vals;                    /* [n_coarse] torch.linspace(0,1,n_coarse) */
  const float* t_rand;                    /* [n,n_coarse] stratified jitter or NULL */
  const float* u;                         /* [n_fine] (u_per_ray == 0) or [n,n_fine] */
  int32_t u_per_ray;
  const float *noise_c, *noise_f;         /* [n,n_coarse] / [n,n_coarse+n_fine], already scaled, or NULL */
  const nvsr_planes_t* planes_coarse;
  const nvsr_planes_t* planes_fine;       /* may equal planes_coarse; same channel count */
  const float* vplane_coarse;             /* channels-last fp32 view plane [vrh][vrw][C] */
  const float* vplane_fine;               /* == vplane_coarse when the models share it */
  int32_t vrh, vrw;
  float az_lo, az_rng, el_lo, el_rng;
  const nvsr_decoder_t* dec_coarse;
  const nvsr_decoder_t* dec_fine;
  float *rgb_c, *disp_c, *acc_c, *depth_c; /* [n,3], [n], [n], [n] */
  float *rgb_f, *disp_f, *acc_f, *depth_f; /* fine maps (n_fine > 0) */
  void* workspace;
  int64_t workspace_bytes;
} nvsr_render_t;

/* bytes of device workspace nvsr_render_rays needs for this request (-1: invalid request) */
int64_t nvsr_workspace_bytes(const nvsr_render_t* request);
int32_t nvsr_render_rays(const nvsr_render_t* request, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BACKWARD of the memory-bound stages (the reference differentiates
 * run_one_iter_of_nerf with autograd, train_nerf.py:860-916 — mse on rgb_coarse / rgb_fine, loss.backward(),
 * PlanesOptimizer.step()).  The decoder between them stays with the caller's autograd in this version
 * (`nvsr_b200.autograd`); z_samples are detached in the reference (train_utils.py:153), so sample_pdf has no
 * backward.  All tensors fp32, rows RAY_MAJOR (row = ray*S + s), as in the reference.
 *
 * a5 backward — scatter-add of the feature gradients through the bilinear footprints of
 * F.grid_sample(bilinear, align_corners=True, padding_mode='border') (models.py:289-310) and the 'avg'
 * combination (models.py:355-361):  d plane_d[y,x,c] += w * (d_feat_p[row, d*C+c] + d_feat_m[row, c] / 3).
 * sampler->z_in ([n,S], the depths the forward used) is required; planes->plane[] is not read (geometry only:
 * rh, rw, channels, box, proj).  d_plane: HOST array of 3 device pointers to channels-last fp32 accumulators
 * [rh][rw][channels] that the caller zeroed (or wants accumulated into); the reference's NCHW parameter gradient is
 * their permutation.  Either of d_feat_p [rows,3C] / d_feat_m [rows,C] may be NULL.  channels % 4 == 0 and 16-byte
 * aligned accumulators (NVSR_ERR_ALIGNMENT otherwise): the updates are 128-bit vector reductions.
 */
int32_t nvsr_sample_gather_bwd(const nvsr_sampler_t* sampler, const nvsr_planes_t* planes, const float* d_feat_p,
                               const float* d_feat_m, float* const d_plane[3], void* stream);

/* a5 backward, view-direction half (cart2az_el nerf_helpers.py:492-496 + project_viewdir models.py:312-326):
 * d_vplane[y,x,c] += w * d_vfeat[ray,c]; d_vplane channels-last fp32 [rh][rw][channels]; arguments as in
 * nvsr_viewdir_gather. */
int32_t nvsr_viewdir_gather_bwd(const float* viewdirs, int64_t n_rays, int32_t rh, int32_t rw, int32_t channels,
                                float az_lo, float az_rng, float el_lo, float el_rng, const float* d_vfeat,
                                float* d_vplane, void* stream);

/* a7 backward — volume_render_radiance_field (volume_rendering_utils.py:15-51, cumprod_exclusive
 * nerf_helpers.py:409-430).  radiance_field / d_radiance_field: [n,S,4] interleaved (the reference's tensor);
 * z: [n,S] depths ([n,S+1] interval edges when mip != 0); noise: [n,S] already scaled by radiance_field_noise_std, or
 * NULL.  Upstream gradients: d_rgb [n,3] (required), d_acc [n], d_depth [n], d_weights [n,S] (each may be NULL).
 * disp_map is not differentiated (the reference's losses do not use it).  d_radiance_field is fully overwritten. */
int32_t nvsr_composite_bwd(const float* radiance_field, const float* z, const float* rd, const float* noise,
                           int64_t n_rays, int32_t n_samples, int32_t white_bkgd, int32_t mip, const float* d_rgb,
                           const float* d_acc, const float* d_depth, const float* d_weights,
                           float* d_radiance_field, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder chain on the tensor cores at fp32-grade accuracy ('fp16-split' precision mode): one tri-plane chain
 * (k0 -> 128 x4 -> head_n, ReLU) with every operand split into two fp16 terms and three tcgen05.mma passes per layer
 * (hi.hi + lo.hi + hi.lo, fp32 accumulation; csrc/mlp_split.cu).  feat: fp32 row-major features [n_rays * n_samples][k0]
 * in RAY-MAJOR rows (the fp32 gather; nvsr_sample_gather with NVSR_FEAT_ROWMAJOR_F32 and feat_p = NULL writes only the
 * combined features); w_hi[l] / w_lo[l]: nvsr_pack_weight16(NVSR_F16) images of W and of W - fp16(W); bias[l]: fp32 [128].
 * Heads are written into planar raw[head_ch ..][raw_stride] in the BLOCKED row order, like nvsr_mlp_chain. */
int32_t nvsr_mlp_chain_split(const float* feat, int32_t k0, const void* const* w_hi, const void* const* w_lo,
                             const float* const* bias, const float* head_w, const float* head_b, int32_t head_n,
                             int32_t head_ch, int64_t n_rays, int32_t n_samples, float* raw, int64_t raw_stride,
                             void* stream);
/* The faster producer / consumer pair of that mode.  nvsr_sample_gather_hilo: the 16-bit tile gather (nvsr_sample_gather,
 * NVSR_FEAT_TILE_F16) reading every plane as TWO fp16 x-pair images — planes->plane[d] = pack(fp16(p)), lo_plane[d] =
 * pack(fp16(p - fp16(p))) — in one pass: feat_p (may be NULL) is the colour chain's fp16 tile image, interpolated from the
 * high halves exactly as nvsr_sample_gather does; feat_m32 is the combined feature interpolated from hi + lo in fp32, as
 * an fp32 tile image [tiles][channels/4][128 rows][4] in BLOCKED rows (padding rows zero); feat_m16 (may be NULL): the
 * same combined feature rounded to fp16, as the plain gather's featM tile image.
 * nvsr_mlp_chain_split_tiled: nvsr_mlp_chain_split reading that image (k0 = channels). */
int32_t nvsr_sample_gather_hilo(const nvsr_sampler_t* sampler, const nvsr_planes_t* planes, const void* const lo_plane[3],
                                void* feat_p, float* feat_m32, void* feat_m16, float* z_out, void* stream);
int32_t nvsr_mlp_chain_split_tiled(const float* feat_tiles, int32_t k0, const void* const* w_hi, const void* const* w_lo,
                                   const float* const* bias, const float* head_w, const float* head_b, int32_t head_n,
                                   int32_t head_ch, int64_t n_rays, int32_t n_samples, float* raw, int64_t raw_stride,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder training path on the tensor cores: forward that keeps what the backward needs, data
 * gradient chain, weight gradients — the reference's loss.backward() through models.py:393-421 (train_nerf.py:860-916).
 * fp16 operands, fp32 accumulation; every delta carries the caller's power-of-two loss `scale`.
 * All images are 16-bit "tile images" [tile][channels/8][128 rows][8] in the BLOCKED row order of the gather.
 *
 * nvsr_mlp_chain_train: nvsr_mlp_chain for one of the two tri-plane chains (4 layers x 128, NVSR_F16, NVSR_ROWS_BLOCKED) that
 *   also stores act_out[l] = image of layer l's post-ReLU 16-bit output x_{l+1} (128 channels), l = 0..3.
 * nvsr_mlp_dgrad: g[3] = (d_raw . head_w) * [x_4 > 0];  g[l-1] = (g[l] . W_l) * [x_l > 0];  d_x0 = g[0] . W_0 (fp32, unscaled,
 *   ray-major rows [n_rays * n_samples][k0]); w[l] are the FORWARD weight images (read MN-major: no transposed copies);
 *   dout_img receives the scaled head gradient as a 16-column image.  d_raw of padding rows must be 0.
 * nvsr_mlp_wgrad: dw[m][n] += inv_scale * sum_rows a_img[r][m] * b_img[r][n] (a: 128 channels, b: n_b channels, n_b % 16 == 0),
 *   db[m] += inv_scale * sum_rows a_img[r][m] (db may be NULL).  dw / db are ACCUMULATED into (fp32 atomics): zero them first.
 *   Layer l: a = g[l], b = x_l (l = 0: the feature image);  head: a = x_4, b = dout_img, n_b = 16 (dw = head weight grad^T).
 * nvsr_ray_sum: out[ray][128] = inv_scale * sum over the ray's samples of a 128-channel image (per-ray bias gradients). */
typedef struct nvsr_dgrad {
  const void* w[4];      /* forward weight images of layers 0..3 (nvsr_pack_weight16, NVSR_F16) */
  int32_t k0;            /* input width of layer 0 (multiple of 16, <= 192) */
  const float* head_w;   /* [head_n][128] fp32 */
  int32_t head_n, head_ch;
  const float* d_raw;    /* planar [4][raw_stride], BLOCKED rows */
  int64_t raw_stride;
  float scale;
  const void* act[4];    /* x_1..x_4 images written by nvsr_mlp_chain_train */
  void* g[4];            /* out: delta images g_0..g_3 */
  void* dout_img;        /* out: [tiles][2][128][8] */
  float* d_x0;           /* out: [n_rays * n_samples][k0] fp32 */
  int64_t n_rays;
  int32_t n_samples;
  const int32_t* row_count; /* NULL, or row-list mode (below): device count of listed rows; g / dout_img / d_x0 come out
                             * LIST-ordered (d_x0 [tiles*128][k0]) and only the tiles the list fills are visited */
  const int32_t* row_ids;   /* row-list mode: NULL = act / d_raw are LIST-ordered already (nvsr_compact_rows); else they
                             * are the forward's BLOCKED buffers, read through the list, and the kernel writes ... */
  void* act_list[4];        /* ... the listed rows of x_1..x_4 here in LIST order (out; the weight gradients' operands) */
  const void* x0_img;       /* optional (both or neither): the forward's k0-channel feature image and ... */
  void* x0_list;            /* ... its listed rows in LIST order (out) */
  int32_t acts_listed;      /* with row_ids: 1 = act is LIST-ordered already (nvsr_mlp_chain_train over the same list);
                             * only d_raw is read through the list and act_list / x0_list are not written */
} nvsr_dgrad_t;

/* mlp->row_ids / row_count (the rgb chain only): the sparse colour path of the forward — input row i stands for BLOCKED
 * row row_ids[i], the heads go to that row of raw, and act_out comes out in LIST order. */
int32_t nvsr_mlp_chain_train(const nvsr_mlp_t* mlp, void* const* act_out, void* stream);
int32_t nvsr_mlp_dgrad(const nvsr_dgrad_t* args, void* stream);
int32_t nvsr_mlp_wgrad(const void* a_img, const void* b_img, int32_t n_b, int64_t n_tiles, float inv_scale, float* dw,
                       int64_t ldw, float* db, void* stream);
int32_t nvsr_ray_sum(const void* img, int64_t n_rays, int32_t n_samples, float inv_scale, float* out, void* stream);
/* all five weight gradients of one chain in one call (the host side of a training step is call-bound): layers 0..3
 * (a = g[l], b = x0_img for l = 0 with k0 channels, else act[l-1]; db[l] bias gradients) and the head (a = act[3],
 * b = dout_img, dw_head [128][16]).  Same accumulation contract as nvsr_mlp_wgrad. */
int32_t nvsr_mlp_wgrad_chain(const void* const* g, const void* x0_img, int32_t k0, const void* const* act,
                             const void* dout_img, int64_t n_tiles, float inv_scale, float* const* dw, const int64_t* ldw,
                             float* const* db, float* dw_head, void* stream);

/* Row-list ("sparse") backward.  A sample whose raw gradient is identically zero — alpha = 0 because sigma + noise <= 0
 * (volume_rendering_utils.py:29-44), or transmittance 0 — adds nothing to any weight or plane gradient, so the
 * backward chains need only the other rows; on a trained scene that is 10-20 % of the samples.  Results equal the
 * dense backward's up to the order of the fp32 sums.
 * nvsr_nonzero_rows: row_ids[0..count) = BLOCKED row ids r < n_rows with d_raw[h][r] != 0 for some h (NaN counts as
 *   non-zero); order unspecified; count (device int32) is reset first.  row_ids needs n_rows entries.
 * nvsr_compact_rows: copies the listed rows of n_img tile images (channels[k] each, multiple of 8) and of the four d_raw
 *   planes into dense tiles in LIST order; the tail of the last tile is zero-filled (it then contributes nothing to
 *   nvsr_mlp_wgrad_chain_rows).  max_tiles: capacity of the outputs in 128-row tiles; out_stride >= max_tiles * 128.
 *   d_raw / d_raw_out may both be NULL (images only).
 * nvsr_mlp_dgrad with row_count + row_ids set gathers the listed rows itself (no separate compaction pass) and emits the
 *   LIST-ordered activation / feature images the weight gradients read; the list's tail rows are zero-filled there too.
 * nvsr_mlp_dgrad with row_count set / nvsr_mlp_wgrad_chain_rows / nvsr_ray_sum_rows / nvsr_sample_gather_bwd_rows: the
 *   dense entries over the LIST-ordered images, reading the row count on the device (no host synchronisation).
 *   nvsr_ray_sum_rows ACCUMULATES into out [n_rays][128] (zero it first). */
int32_t nvsr_nonzero_rows(const float* d_raw, int64_t raw_stride, int64_t n_rows, int32_t* row_ids, int32_t* count,
                          void* stream);
int32_t nvsr_compact_rows(const void* const* src, void* const* dst, const int32_t* channels, int32_t n_img,
                          const float* d_raw, int64_t raw_stride, float* d_raw_out, int64_t out_stride,
                          const int32_t* row_ids, const int32_t* count, int64_t max_tiles, void* stream);
int32_t nvsr_mlp_wgrad_chain_rows(const void* const* g, const void* x0_img, int32_t k0, const void* const* act,
                                  const void* dout_img, int64_t n_tiles, const int32_t* row_count, float inv_scale,
                                  float* const* dw, const int64_t* ldw, float* const* db, float* dw_head, void* stream);
int32_t nvsr_ray_sum_rows(const void* img, const int32_t* row_ids, const int32_t* count, int64_t max_rows,
                          int32_t n_samples, float inv_scale, float* out, void* stream);
int32_t nvsr_sample_gather_bwd_rows(const nvsr_sampler_t* sampler, const nvsr_planes_t* planes, const float* d_feat_p,
                                    const float* d_feat_m, const int32_t* row_ids, const int32_t* count, int64_t max_rows,
                                    float* const d_plane[3], void* stream);

/* out[ray][0 .. sa+sb) = sort(cat(a[ray][0..sa), b[ray][0..sb))) ascending — the merge of the coarse depths with the
 * inverse-CDF samples (train_utils.py:144-156) when those are not sorted (training with perturbation).  One warp per
 * ray, bitonic network in registers; sa + sb <= 512 (NVSR_ERR_UNSUPPORTED beyond).  Values are moved, never changed;
 * NaNs sort last (torch.sort). */
int32_t nvsr_sort_cat(const float* a, int32_t sa, const float* b, int32_t sb, int64_t n_rays, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Plane super-resolution, last step: PlanesSR.forward (models.py:884-926) ends with
 *   out = inner_model(pad(LR))[..., crop:-crop, crop:-crop] + F.interpolate(LR, scale_factor, 'bilinear', align_corners)
 * then caches `out` on the CPU (:925) and re-uploads it on every call (:893).  This entry fuses everything after the
 * conv chain and writes the plane DIRECTLY in the layout the gather reads, device-resident:
 *   diff        : conv-chain output, channels-last — element (y, x, c) at diff[y*row_stride + x*px_stride + c]
 *                 (strides in elements; dtype NVSR_F32 | NVSR_BF16 | NVSR_F16); `crop` = PlanesSR.HR_overpadding
 *   lr_nchw     : the LR plane [channels][rh][rw] fp32 as stored in planes_ (models.py:436-439)
 *   packed      : nvsr_pack_plane image of the [rh*scale][rw*scale] SR plane (packed_dtype NVSR_F32: channels-last fp32;
 *                 NVSR_BF16 | NVSR_F16: x-pair records, 32-byte aligned), or NULL
 *   nchw_out    : optional fp32 [channels][rh*scale][rw*scale] copy for the reference's own consumers, or NULL
 * Bilinear weights follow ATen's upsample_bilinear2d (area_pixel_compute_source_index). */
int32_t nvsr_sr_finalize(const void* diff, int32_t diff_dtype, int64_t diff_row_stride, int64_t diff_px_stride,
                         int32_t crop, const float* lr_nchw, int32_t channels, int32_t rh, int32_t rw, int32_t scale,
                         int32_t align_corners, void* packed, int32_t packed_dtype, float* nchw_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Frame sink: write_image's conversion (train_nerf.py:270,273:
 * np.array(255*torch.clamp(im,0,1).cpu()).astype(np.uint8)) done on the device, so the device->host copy carries one
 * byte per channel.  rgb: n_elems fp32 values (a frame [H,W,3] flat, or any map); out: n_elems bytes, 4-byte aligned
 * (NVSR_ERR_ALIGNMENT otherwise).  fp32 product, truncating cast, NaN -> 0 (what the x86 cast of the reference gives). */
int32_t nvsr_frame_to_u8(const float* rgb, int64_t n_elems, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVSR_H_ */
