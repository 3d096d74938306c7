// a7+a8: alpha compositing with cumulative transmittance, inverse-CDF resampling and sort-merge.
// One warp per ray; samples are striped over the lanes (sample i -> lane i%32) so every global
// access is coalesced, and the two scans (cumprod of 1-alpha, cumsum of the pdf) are warp-shuffle
// scans with a running carry.  Scans run in fp64 and are rounded to fp32 per prefix: that is what
// the reference's CPU path does (ATen accumulates float cumsum/cumprod in double), so cdf edges —
// and with them the searchsorted bin indices — track the oracle.
#include "common.cuh"

namespace nvsr {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_up_d(double v, int off) { return __shfl_up_sync(kFull, v, off); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// inclusive scans over the 32 lanes
__device__ __forceinline__ double warp_scan_mul(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_add(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// torch.sigmoid on CPU: 1/(1+exp(-x))
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// total order used by the slow-path merge: NaN sorts last (torch.sort), ties by index
__device__ __forceinline__ bool key_less(float a, int ia, float b, int ib) {
  bool na = isnan(a), nb = isnan(b);
  if (na || nb) return (!na && nb) || (na && nb && ia < ib);
  return a < b || (a == b && ia < ib);
}

// searchsorted(cdf[0..n), u, side='right'): first index with cdf[idx] > u, n if none
__device__ __forceinline__ int upper_bound_f(const float* a, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] > u) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound_f(const float* a, int n, float x) {  // #(a < x)
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// nerf_helpers.py:686-700 given cdf/bins (B entries) in shared memory
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u, int* ind_out) {
  int ind = upper_bound_f(cdf, B, u);
  int below = max(0, ind - 1);
  int above = min(B - 1, ind);
  float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  float bb = bins[below], ba = bins[above];
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// torch.sum(x, -1) of ATen's CPU kernel for a contiguous inner reduction of n fp32 values
// (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum -> row_sum -> multi_row_sum): 8-lane
// vectors, 4 interleaved vector accumulators, a 4-level cascade every 16 steps, then the leftover
// vectors, the scalar tail and the 8 lane partials added sequentially.  Floating-point sums are
// order dependent; reproducing THIS order makes `total` — and with it every cdf edge and every
// searchsorted index — bit-identical to the reference's CPU path for identical weights (probed:
// 100% agreement with torch.sum for n in 14..2050).  The 32 (accumulator, lane) pairs map onto the
// 32 lanes of the warp.  x(i) = w[i] + 1e-5 (nerf_helpers.py:672).
__device__ __forceinline__ float aten_sum_w(const float* w, int n, int lane) {
  auto X = [&](int i) { return __fadd_rn(w[i], 1e-5f); };
  if (n < 8) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, X(i));
    return s;
  }
  const int vec_size = n >> 3, size_ilp = vec_size >> 2;
  const int k = lane >> 3, L = lane & 7;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int i = 0;
  while (i + 16 <= size_ilp) {
    for (int j = 0; j < 16; ++j, ++i) acc[0] = __fadd_rn(acc[0], X(((i << 2) + k) * 8 + L));
    for (int j = 1; j < 4; ++j) {
      acc[j] = __fadd_rn(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (15 << (j * 4))) != 0) break;
    }
  }
  for (; i < size_ilp; ++i) acc[0] = __fadd_rn(acc[0], X(((i << 2) + k) * 8 + L));
  for (int j = 1; j < 4; ++j) acc[0] = __fadd_rn(acc[0], acc[j]);
  if (k == 0)
    for (int v = size_ilp << 2; v < vec_size; ++v) acc[0] = __fadd_rn(acc[0], X(v * 8 + L));
  float p = acc[0];
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 8));
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 16));
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 24));
  float fin = 0.f;
  for (int idx = vec_size << 3; idx < n; ++idx) fin = __fadd_rn(fin, X(idx));
  for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(kFull, p, l));
  return fin;
}

// pdf/cdf of sample_pdf_2 (nerf_helpers.py:673-676) from weights w[0..nw) in shared memory:
// cdf[0]=0, cdf[i] = sum_{m<=i-1} (w[m]+1e-5)/total, i = 1..nw  -> nw+1 entries.
// cumsum: ATen accumulates fp32 cumsum in double on the CPU and rounds each prefix to fp32.
__device__ __forceinline__ void build_cdf(const float* w, int nw, float* cdf, int lane) {
  const float total = aten_sum_w(w, nw, lane);
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nw; base += 32) {
    int i = base + lane;
    double pdf = (i < nw) ? (double)__fdiv_rn(__fadd_rn(w[i], 1e-5f), total) : 0.0;
    double inc = warp_scan_add(pdf, lane);
    if (i < nw) cdf[i + 1] = (float)(carry + inc);
    carry += shfl_d(inc, 31);
  }
}

struct CompositeArgs {
  int64_t n_rays;
  int S;
  const float* raw;
  int64_t raw_stride;
  const float* z;
  const float* rd;
  const float* noise;
  int white_bkgd, mip;
  float *rgb, *disp, *acc, *depth, *weights;
  int n_fine;
  const float* u;
  int u_per_ray;
  int64_t* inds;
  float* z_samples;
  float* z_merged;
  int row_order;
};

// BLOCKED raw staging: the 8 rays of a block own one contiguous span of TS*128 floats per channel.
// The CTA copies it (coalesced) into shared memory as [ch][ray-in-block][pitch]; pitch % 32 == 4 makes
// both the transposing store (8 rays x 4 samples per warp) and the per-ray reads conflict-free.
__host__ __device__ inline int stage_pitch(int S) { return ((S - 4 + 31) / 32) * 32 + 4; }

// per-warp shared floats: zv[S+1] | w[S] | cdf[S] | bins[S] | zs[n_fine]
__host__ __device__ inline int composite_smem_floats(int S, int n_fine) {
  return (S + 1) + (n_fine > 0 ? 3 * S + n_fine : 0) + 3;
}

__global__ void __launch_bounds__(256)
composite_kernel(CompositeArgs a, int warps_per_cta) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int S = a.S;
  const int S1 = S + (a.mip ? 1 : 0);  // depth entries per ray
  float* zv = sm + (size_t)wid * composite_smem_floats(S, a.n_fine);
  float* wbuf = zv + (S + 1);
  float* cdf = wbuf + S;
  float* bins = cdf + S;
  float* zs = bins + S;
  const bool blocked = a.row_order == NVSR_ROWS_BLOCKED;  // then warps_per_cta == kBlkRays
  float* sraw = sm + (size_t)warps_per_cta * composite_smem_floats(S, a.n_fine);
  const int P = stage_pitch(S);
  const int TS = tiles_per_block(S);
  const int64_t n_groups = ceil_div64(a.n_rays, warps_per_cta);

  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t ray = grp * warps_per_cta + wid;
    if (blocked) {
      __syncthreads();  // previous group's readers are done with sraw
      const int span = TS * kTileRows;
      for (int ch = 0; ch < 4; ++ch) {
        const float* src = a.raw + ch * a.raw_stride + grp * span;
        for (int e = threadIdx.x; e < span; e += blockDim.x) {
          int r = e & (kTileRows - 1);
          int sidx = (e >> 7) * kBlkSamples + (r >> 3);
          if (sidx < S) sraw[(ch * kBlkRays + (r & 7)) * P + sidx] = __ldg(src + e);
        }
      }
      __syncthreads();
    }
    if (ray >= a.n_rays) continue;
    __syncwarp();
    for (int i = lane; i < S1; i += 32) zv[i] = __ldg(a.z + ray * S1 + i);
    float dx = __ldg(a.rd + ray * 3), dy = __ldg(a.rd + ray * 3 + 1), dz = __ldg(a.rd + ray * 3 + 2);
    float dnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    __syncwarp();

    const float* raw_r = blocked ? sraw + wid * P : a.raw + ray * S;
    const int64_t chs = blocked ? (int64_t)kBlkRays * P : a.raw_stride;
    double carry = 1.0;  // product of (1-alpha+1e-10) over previous chunks
    double sr = 0.0, sg = 0.0, sb = 0.0, sd = 0.0, sa = 0.0;
    for (int base = 0; base < S; base += 32) {
      int i = base + lane;
      float alpha = 0.f, r = 0.f, g = 0.f, b = 0.f, zc = 0.f;
      double t = 1.0;
      if (i < S) {
        float dist;
        if (a.mip) {
          dist = __fsub_rn(zv[i + 1], zv[i]);
          zc = __fmul_rn(0.5f, __fadd_rn(zv[i], zv[i + 1]));
        } else {
          dist = (i + 1 < S) ? __fsub_rn(zv[i + 1], zv[i]) : 1e10f;
          zc = zv[i];
        }
        dist = __fmul_rn(dist, dnorm);
        r = sigmoid_ref(raw_r[i]);
        g = sigmoid_ref(raw_r[chs + i]);
        b = sigmoid_ref(raw_r[2 * chs + i]);
        float sig = raw_r[3 * chs + i];
        if (a.noise) sig = __fadd_rn(sig, __ldg(a.noise + ray * S + i));
        sig = fmaxf(sig, 0.f);
        alpha = __fsub_rn(1.f, expf(__fmul_rn(-sig, dist)));
        t = (double)__fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
      }
      double inc = warp_scan_mul(t, lane);
      double prev = shfl_up_d(inc, 1);
      double excl = carry * (lane ? prev : 1.0);
      carry *= shfl_d(inc, 31);
      if (i < S) {
        float T = (float)excl;
        float w = __fmul_rn(alpha, T);
        if (a.weights) a.weights[ray * S + i] = w;
        if (a.n_fine > 0) wbuf[i] = w;
        sr += (double)__fmul_rn(w, r);
        sg += (double)__fmul_rn(w, g);
        sb += (double)__fmul_rn(w, b);
        sd += (double)__fmul_rn(w, zc);
        sa += (double)w;
      }
    }
    sr = warp_sum_d(sr), sg = warp_sum_d(sg), sb = warp_sum_d(sb), sd = warp_sum_d(sd), sa = warp_sum_d(sa);
    if (lane == 0) {
      float accv = (float)sa, depthv = (float)sd;
      float cr = (float)sr, cg = (float)sg, cb = (float)sb;
      // 1/max(1e-10, depth/acc): torch.max propagates NaN (acc == 0 -> 0/0)
      float q = __fdiv_rn(depthv, accv);
      float m = isnan(q) ? q : fmaxf(1e-10f, q);
      float dispv = __fdiv_rn(1.f, m);
      if (a.white_bkgd) {
        float bg = __fsub_rn(1.f, accv);
        cr = __fadd_rn(cr, bg), cg = __fadd_rn(cg, bg), cb = __fadd_rn(cb, bg);
      }
      a.rgb[ray * 3] = cr, a.rgb[ray * 3 + 1] = cg, a.rgb[ray * 3 + 2] = cb;
      a.disp[ray] = dispv, a.acc[ray] = accv, a.depth[ray] = depthv;
    }
    if (a.n_fine <= 0) continue;

    // ---- hierarchical resampling: train_utils.py:144-156 + nerf_helpers.py:668-702 ----
    __syncwarp();
    const int B = S - 1;  // bins (z_mid; mip: mids of mids)
    for (int i = lane; i < B; i += 32) {
      float m0 = __fmul_rn(0.5f, __fadd_rn(zv[i + 1], zv[i]));
      if (a.mip) {
        float m1 = __fmul_rn(0.5f, __fadd_rn(zv[i + 2], zv[i + 1]));
        m0 = __fmul_rn(0.5f, __fadd_rn(m1, m0));
      }
      bins[i] = m0;
    }
    build_cdf(wbuf + 1, S - 2, cdf, lane);  // weights[...,1:-1]
    __syncwarp();
    const int nf = a.n_fine;
    for (int j = lane; j < nf; j += 32) {
      float u = a.u_per_ray ? __ldg(a.u + ray * nf + j) : __ldg(a.u + j);
      int ind;
      float smp = invert_cdf(cdf, bins, B, u, &ind);
      zs[j] = smp;
      if (a.inds) a.inds[ray * nf + j] = ind;
      if (a.z_samples) a.z_samples[ray * nf + j] = smp;
    }
    __syncwarp();
    // ---- sort(cat(z_vals, z_samples)): merge by rank ----
    bool sorted = true;
    for (int i = lane; i + 1 < S1; i += 32) sorted &= (zv[i] <= zv[i + 1]);
    for (int j = lane; j + 1 < nf; j += 32) sorted &= (zs[j] <= zs[j + 1]);
    sorted = __all_sync(kFull, sorted);
    float* out = a.z_merged + ray * (int64_t)(S1 + nf);
    if (sorted) {
      for (int i = lane; i < S1; i += 32) out[i + lower_bound_f(zs, nf, zv[i])] = zv[i];
      for (int j = lane; j < nf; j += 32) out[j + upper_bound_f(zv, S1, zs[j])] = zs[j];
    } else {
      // rare (rounding-induced inversions, random u, NaNs): all-pairs ranking, still exact
      const int M = S1 + nf;
      for (int e = lane; e < M; e += 32) {
        float x = e < S1 ? zv[e] : zs[e - S1];
        int rank = 0;
        for (int k = 0; k < M; ++k) {
          float y = k < S1 ? zv[k] : zs[k - S1];
          rank += key_less(y, k, x, e) ? 1 : 0;
        }
        out[rank] = x;
      }
    }
  }
}

// stand-alone sample_pdf: bins [n,B], weights [n,B-1] (or cdf_in [n,B])
__global__ void __launch_bounds__(256)
sample_pdf_kernel(const float* __restrict__ bins_g, const float* __restrict__ w_g, const float* __restrict__ cdf_in,
                  int64_t n_rays, int B, const float* __restrict__ u_g, int u_per_ray, int nf,
                  int64_t* __restrict__ inds, float* __restrict__ samples, float* __restrict__ cdf_out,
                  int warps_per_cta) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* bins = sm + (size_t)wid * (3 * B + 2);
  float* cdf = bins + B;
  float* w = cdf + B + 1;
  for (int64_t ray = (int64_t)blockIdx.x * warps_per_cta + wid; ray < n_rays;
       ray += (int64_t)gridDim.x * warps_per_cta) {
    __syncwarp();
    for (int i = lane; i < B; i += 32) bins[i] = __ldg(bins_g + ray * B + i);
    if (cdf_in) {
      for (int i = lane; i < B; i += 32) cdf[i] = __ldg(cdf_in + ray * B + i);
    } else {
      for (int i = lane; i < B - 1; i += 32) w[i] = __ldg(w_g + ray * (B - 1) + i);
      __syncwarp();
      build_cdf(w, B - 1, cdf, lane);
    }
    __syncwarp();
    if (cdf_out)
      for (int i = lane; i < B; i += 32) cdf_out[ray * B + i] = cdf[i];
    for (int j = lane; j < nf; j += 32) {
      float u = u_per_ray ? __ldg(u_g + ray * nf + j) : __ldg(u_g + j);
      int ind;
      float smp = invert_cdf(cdf, bins, B, u, &ind);
      if (inds) inds[ray * nf + j] = ind;
      if (samples) samples[ray * nf + j] = smp;
    }
  }
}

template <typename K>
static int32_t pick_warps(K kernel, size_t floats_per_warp, int* warps_out, size_t* smem_out) {
  size_t per_warp = floats_per_warp * sizeof(float);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  if (per_warp * warps > 200 * 1024) return NVSR_ERR_RESOURCE;
  size_t smem = per_warp * warps;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int32_t)e;
  }
  *warps_out = warps;
  *smem_out = smem;
  return NVSR_OK;
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_composite(const nvsr_composite_t* c, void* stream) {
  NVSR_CHECK_ARG(c && c->n_rays >= 0 && c->n_samples > 0 && c->n_samples <= NVSR_MAX_SAMPLES);
  NVSR_CHECK_ARG(c->raw && c->z && c->rd && c->rgb && c->disp && c->acc && c->depth);
  if (c->n_fine > 0) {
    NVSR_CHECK_ARG(c->u && c->z_merged && c->n_samples >= 3 && c->n_fine <= NVSR_MAX_SAMPLES);
  }
  if (c->n_rays == 0) return NVSR_OK;
  NVSR_CHECK_ARG(c->row_order == NVSR_ROWS_RAY_MAJOR || c->row_order == NVSR_ROWS_BLOCKED);
  NVSR_CHECK_ARG(c->raw_stride >= rows_padded(c->n_rays, c->n_samples, c->row_order));
  CompositeArgs a{c->n_rays, c->n_samples, c->raw, c->raw_stride, c->z, c->rd, c->noise, c->white_bkgd, c->mip,
                  c->rgb, c->disp, c->acc, c->depth, c->weights, c->n_fine > 0 ? c->n_fine : 0, c->u, c->u_per_ray,
                  c->inds, c->z_samples, c->z_merged, c->row_order};
  int warps;
  size_t smem;
  if (a.row_order == NVSR_ROWS_BLOCKED) {
    // one CTA = the 8 rays of a block, staged through shared memory
    warps = kBlkRays;
    smem = ((size_t)warps * composite_smem_floats(a.S, a.n_fine) + (size_t)4 * kBlkRays * stage_pitch(a.S)) * sizeof(float);
    if (smem > 200 * 1024) return NVSR_ERR_RESOURCE;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int32_t)e;
    }
  } else {
    int32_t st = pick_warps(composite_kernel, composite_smem_floats(a.S, a.n_fine), &warps, &smem);
    if (st != NVSR_OK) return st;
  }
  int64_t blocks = ceil_div64(c->n_rays, warps);
  int64_t max_blocks = (int64_t)kNumSMs * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  composite_kernel<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(a, warps);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_sample_pdf(const float* bins, const float* weights, const float* cdf_in, int64_t n_rays,
                                   int32_t n_bins, const float* u, int32_t u_per_ray, int32_t n_samples,
                                   int64_t* inds, float* samples, float* cdf_out, void* stream) {
  NVSR_CHECK_ARG(bins && (weights || cdf_in) && u && n_rays >= 0 && n_bins >= 2 && n_bins <= NVSR_MAX_SAMPLES);
  NVSR_CHECK_ARG(n_samples > 0);
  if (n_rays == 0) return NVSR_OK;
  int warps;
  size_t smem;
  int32_t st = pick_warps(sample_pdf_kernel, (size_t)3 * n_bins + 2, &warps, &smem);
  if (st != NVSR_OK) return st;
  int64_t blocks = ceil_div64(n_rays, warps);
  int64_t max_blocks = (int64_t)kNumSMs * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  sample_pdf_kernel<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, cdf_in, n_rays, n_bins, u, u_per_ray, n_samples, inds, samples, cdf_out, warps);
  NVSR_RETURN_LAST_ERROR();
}
