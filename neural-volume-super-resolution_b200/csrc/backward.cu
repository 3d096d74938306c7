// Backward of the gather and compositing stages (first, correctness-first version).
//
// The reference differentiates run_one_iter_of_nerf with autograd (train_nerf.py:860-916: mse on rgb_coarse and
// rgb_fine, loss.backward(), PlanesOptimizer.step()).  These kernels are the hand-written backward of the two
// memory-bound stages either side of the decoder:
//   composite_bwd_kernel     d rgb_map (+ d acc / d depth / d weights) -> d radiance_field        (a7 backward)
//   gather_bwd_kernel        d features -> scatter-add into the tri-plane gradients               (a5 backward)
//   viewdir_gather_bwd_kernel  d per-ray view features -> scatter-add into the view plane gradient
// z_samples are detached in the reference (train_utils.py:153), so nothing flows back through sample_pdf.
// The per-element arithmetic lives in backward_bodies.h, which also compiles for the host
// (tests/hostcheck): what runs here is what the CPU test checked against autograd.
#include "backward_bodies.h"
#include "common.cuh"

namespace nvsr {

struct GatherBwdArgs {
  bwd::PlaneGeom g;
  int64_t n_rays;
  int S;
  const float* ro;
  const float* rd;
  const float* z;   // [n,S]
  const float* d_feat_p;
  const float* d_feat_m;
  float* d_plane[3];
};

// one thread per (row, 4-channel chunk): consecutive threads hold consecutive channel chunks of one row, so the
// four corner updates of a warp are runs of adjacent 128-bit vector reductions (RED.E.ADD.F32x4) into channels-last
// accumulators
__global__ void __launch_bounds__(256) gather_bwd_kernel(GatherBwdArgs a) {
  const int chunks = a.g.C / 4;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_rays * a.S * chunks) return;
  int64_t row = idx / chunks;
  int ch = (int)(idx % chunks) * 4;
  int64_t ray = row / a.S;
  bwd::gather_bwd_row(a.g, a.ro, a.rd, a.z[row], ray, row, ch, a.d_feat_p, a.d_feat_m, a.d_plane);
}

__global__ void __launch_bounds__(256)
viewdir_gather_bwd_kernel(const float* __restrict__ viewdirs, int64_t n, int rh, int rw, int C, float az_lo, float az_rng,
                          float el_lo, float el_rng, const float* __restrict__ d_vfeat, float* d_vplane) {
  const int chunks = C / 4;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * chunks) return;
  bwd::viewdir_gather_bwd_ray(viewdirs, idx / chunks, (int)(idx % chunks) * 4, rh, rw, C, az_lo, az_rng, el_lo, el_rng,
                              d_vfeat, d_vplane);
}

// one thread per ray (a training batch is 4 096 rays x <= 192 samples, config/TrainModels.yml:8: launch-bound)
__global__ void __launch_bounds__(128)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rd,
                     const float* __restrict__ noise, int64_t n, int S, int white, int mip,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_acc, const float* __restrict__ g_depth,
                     const float* __restrict__ g_w, float* __restrict__ d_raw) {
  int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n) return;
  float dx = rd[ray * 3 + 0], dy = rd[ray * 3 + 1], dz = rd[ray * 3 + 2];
  float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const int Z = S + (mip ? 1 : 0);
  bwd::composite_bwd_ray(raw + ray * S * 4, z + ray * Z, nrm, noise ? noise + ray * S : nullptr, S, white, mip,
                         g_rgb + ray * 3, g_acc ? g_acc + ray : nullptr, g_depth ? g_depth + ray : nullptr,
                         g_w ? g_w + ray * S : nullptr, d_raw + ray * S * 4);
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_sample_gather_bwd(const nvsr_sampler_t* s, const nvsr_planes_t* pl, const float* d_feat_p,
                                          const float* d_feat_m, float* const d_plane[3], void* stream) {
  NVSR_CHECK_ARG(s && pl && d_plane && (d_feat_p || d_feat_m));
  NVSR_CHECK_ARG(s->n_rays >= 0 && s->n_samples > 0 && s->ro && s->rd && s->z_in);
  NVSR_CHECK_ARG(pl->channels > 0 && pl->channels % 4 == 0);
  GatherBwdArgs a;
  for (int d = 0; d < 3; ++d) {
    NVSR_CHECK_ARG(d_plane[d] && pl->rh[d] > 0 && pl->rw[d] > 0);
    if (!aligned16(d_plane[d])) return NVSR_ERR_ALIGNMENT;   // 128-bit vector reductions
    a.g.rh[d] = pl->rh[d], a.g.rw[d] = pl->rw[d];
    a.g.lo[d] = pl->box_lo[d], a.g.rng[d] = pl->box_rng[d];
    for (int k = 0; k < 6; ++k) a.g.proj[d][k] = pl->proj[d][k];
    a.d_plane[d] = d_plane[d];
  }
  a.g.C = pl->channels;
  a.n_rays = s->n_rays, a.S = s->n_samples;
  a.ro = s->ro, a.rd = s->rd, a.z = s->z_in;
  a.d_feat_p = d_feat_p, a.d_feat_m = d_feat_m;
  int64_t total = a.n_rays * a.S * (a.g.C / 4);
  if (total == 0) return NVSR_OK;
  gather_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_viewdir_gather_bwd(const float* viewdirs, int64_t n_rays, int32_t rh, int32_t rw,
                                           int32_t channels, float az_lo, float az_rng, float el_lo, float el_rng,
                                           const float* d_vfeat, float* d_vplane, void* stream) {
  NVSR_CHECK_ARG(viewdirs && d_vfeat && d_vplane && n_rays >= 0 && rh > 0 && rw > 0);
  NVSR_CHECK_ARG(channels > 0 && channels % 4 == 0);
  if (!aligned16(d_vplane)) return NVSR_ERR_ALIGNMENT;
  int64_t total = n_rays * (channels / 4);
  if (total == 0) return NVSR_OK;
  viewdir_gather_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      viewdirs, n_rays, rh, rw, channels, az_lo, az_rng, el_lo, el_rng, d_vfeat, d_vplane);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_composite_bwd(const float* radiance_field, const float* z, const float* rd, const float* noise,
                                      int64_t n_rays, int32_t n_samples, int32_t white_bkgd, int32_t mip,
                                      const float* d_rgb, const float* d_acc, const float* d_depth,
                                      const float* d_weights, float* d_radiance_field, void* stream) {
  NVSR_CHECK_ARG(radiance_field && z && rd && d_rgb && d_radiance_field && n_rays >= 0 && n_samples > 0);
  if (n_rays == 0) return NVSR_OK;
  composite_bwd_kernel<<<(unsigned)ceil_div64(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      radiance_field, z, rd, noise, n_rays, n_samples, white_bkgd, mip, d_rgb, d_acc, d_depth, d_weights,
      d_radiance_field);
  NVSR_RETURN_LAST_ERROR();
}
