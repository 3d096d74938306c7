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
  const int32_t* ids;     // row-list mode: d_feat rows are in LIST order, ids[i] = BLOCKED row id of list row i
  const int32_t* count;
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

// row-list variant (the sparse training backward): list row i carries the gradient of BLOCKED row ids[i]
__global__ void __launch_bounds__(256) gather_bwd_rows_kernel(GatherBwdArgs a) {
  const int chunks = a.g.C / 4;
  const int TS = tiles_per_block(a.S);
  // grid-stride over the listed rows (device-side count): no blocks that only exit
  const int64_t total = (int64_t)__ldg(a.count) * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / chunks;
    const int ch = (int)(idx % chunks) * 4;
    const int32_t rid = __ldg(a.ids + i);
    int64_t ray;
    int s;
    blocked_decode(rid / kTileRows, rid % kTileRows, TS, &ray, &s);
    if (ray >= a.n_rays || s >= a.S) continue;
    bwd::gather_bwd_row(a.g, a.ro, a.rd, a.z[ray * a.S + s], ray, i, ch, a.d_feat_p, a.d_feat_m, a.d_plane);
  }
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


// Warp-per-ray variant for S <= 32 * kCbM samples (a training batch is 4 096 rays: one thread per ray leaves most SMs
// idle and walks 2 x S dependent steps).  Lane l owns samples [l * m, (l + 1) * m): the transmittance is a lane-local
// running product on top of an exclusive warp scan of the lanes' products, the suffix sum of dL/dw_k * w_k a lane-local
// running sum on top of a reverse exclusive scan — the same formulas as composite_bwd_ray (backward_bodies.h), another
// association order of the fp32 products / sums (the GPU tests compare both with autograd of the oracle).
constexpr int kCbM = 8;

__global__ void __launch_bounds__(128)
composite_bwd_warp_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rd,
                          const float* __restrict__ noise, int64_t n, int S, int white, int mip,
                          const float* __restrict__ g_rgb, const float* __restrict__ g_acc,
                          const float* __restrict__ g_depth, const float* __restrict__ g_w, float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n) return;
  const float dx = rd[ray * 3 + 0], dy = rd[ray * 3 + 1], dz = rd[ray * 3 + 2];
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const int Z = S + (mip ? 1 : 0);
  const float* zr = z + ray * Z;
  const float4* rr = reinterpret_cast<const float4*>(raw) + ray * S;
  float4* out = reinterpret_cast<float4*>(d_raw) + ray * S;
  const int m = (S + 31) >> 5;
  const int i0 = lane * m;
  const float gr = g_rgb[ray * 3 + 0], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
  const float gsum = white ? gr + gg + gb : 0.f;
  const float ga = g_acc ? g_acc[ray] : 0.f, gd = g_depth ? g_depth[ray] : 0.f;

  float q[kCbM], a_[kCbM], e_[kCbM], dist_[kCbM], dwv[kCbM];
  float4 rw[kCbM];
  bool pos[kCbM];
  float prod = 1.f;
#pragma unroll
  for (int j = 0; j < kCbM; ++j) {
    const int i = i0 + j;
    q[j] = 1.f, a_[j] = 0.f, e_[j] = 1.f, dist_[j] = 0.f, dwv[j] = 0.f, pos[j] = false;
    rw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < m && i < S) {
      rw[j] = __ldg(rr + i);
      float dist = (i + 1 < S || mip) ? __fsub_rn(zr[i + 1], zr[i]) : 1e10f;
      dist = __fmul_rn(dist, nrm);
      const float pre = rw[j].w + (noise ? noise[ray * S + i] : 0.f);
      const float sg = pre > 0.f ? pre : 0.f;
      const float e = expf(-__fmul_rn(sg, dist));
      const float a = __fsub_rn(1.f, e);
      pos[j] = pre > 0.f, dist_[j] = dist, e_[j] = e, a_[j] = a;
      q[j] = __fadd_rn(__fsub_rn(1.f, a), 1e-10f);
      prod *= q[j];
    }
  }
  // exclusive scan of the lanes' products -> transmittance at the lane's first sample
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= up;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  // forward sweep: T_i, w_i, dL/dw_i, the colour gradients; lane total of dL/dw * w
  float Ti[kCbM], wv[kCbM];
  float lane_sum = 0.f;
#pragma unroll
  for (int j = 0; j < kCbM; ++j) {
    const int i = i0 + j;
    Ti[j] = T, wv[j] = 0.f;
    if (j < m && i < S) {
      const float w = a_[j] * T;
      wv[j] = w;
      const float t = mip ? 0.5f * (zr[i] + zr[i + 1]) : zr[i];
      float dw = ga + gd * t + (g_w ? g_w[ray * S + i] : 0.f) - gsum;
      const float c0 = 1.f / (1.f + expf(-rw[j].x)), c1 = 1.f / (1.f + expf(-rw[j].y)), c2 = 1.f / (1.f + expf(-rw[j].z));
      dw += gr * c0 + gg * c1 + gb * c2;
      dwv[j] = dw;
      rw[j].x = w * gr * c0 * (1.f - c0), rw[j].y = w * gg * c1 * (1.f - c1), rw[j].z = w * gb * c2 * (1.f - c2);
      lane_sum += dw * w;
      T *= q[j];
    }
  }
  // reverse exclusive scan of the lane totals -> sum over every later lane's samples
  float rincl = lane_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float dn = __shfl_down_sync(0xffffffffu, rincl, o);
    if (lane + o < 32) rincl += dn;
  }
  float suffix = __shfl_down_sync(0xffffffffu, rincl, 1);
  if (lane == 31) suffix = 0.f;
  // backward sweep within the lane
#pragma unroll
  for (int j = kCbM - 1; j >= 0; --j) {
    const int i = i0 + j;
    if (j < m && i < S) {
      const float da = dwv[j] * Ti[j] - suffix / q[j];
      rw[j].w = pos[j] ? da * dist_[j] * e_[j] : 0.f;
      suffix += dwv[j] * wv[j];
      out[i] = rw[j];
    }
  }
}

}  // namespace nvsr

using namespace nvsr;

static int32_t gather_bwd_launch(const nvsr_sampler_t* s, const nvsr_planes_t* pl, const float* d_feat_p,
                                 const float* d_feat_m, const int32_t* row_ids, const int32_t* count, int64_t max_rows,
                                 float* const d_plane[3], void* stream) {
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
  a.ids = row_ids, a.count = count;
  int64_t total = (row_ids ? max_rows : a.n_rays * a.S) * (a.g.C / 4);
  if (total == 0) return NVSR_OK;
  if (row_ids) {
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    gather_bwd_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  }
  else
    gather_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_sample_gather_bwd(const nvsr_sampler_t* s, const nvsr_planes_t* pl, const float* d_feat_p,
                                          const float* d_feat_m, float* const d_plane[3], void* stream) {
  return gather_bwd_launch(s, pl, d_feat_p, d_feat_m, nullptr, nullptr, 0, d_plane, stream);
}

extern "C" int32_t nvsr_sample_gather_bwd_rows(const nvsr_sampler_t* s, const nvsr_planes_t* pl, const float* d_feat_p,
                                               const float* d_feat_m, const int32_t* row_ids, const int32_t* count,
                                               int64_t max_rows, float* const d_plane[3], void* stream) {
  NVSR_CHECK_ARG(row_ids && count && max_rows >= 0);
  return gather_bwd_launch(s, pl, d_feat_p, d_feat_m, row_ids, count, max_rows, d_plane, stream);
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
  if (n_samples <= 32 * kCbM && aligned16(radiance_field) && aligned16(d_radiance_field))
    composite_bwd_warp_kernel<<<(unsigned)ceil_div64(n_rays, 4), 128, 0, (cudaStream_t)stream>>>(
        radiance_field, z, rd, noise, n_rays, n_samples, white_bkgd, mip, d_rgb, d_acc, d_depth, d_weights,
        d_radiance_field);
  else
    composite_bwd_kernel<<<(unsigned)ceil_div64(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        radiance_field, z, rd, noise, n_rays, n_samples, white_bkgd, mip, d_rgb, d_acc, d_depth, d_weights,
        d_radiance_field);
  NVSR_RETURN_LAST_ERROR();
}
