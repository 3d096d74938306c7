// a9: mip-NeRF conical-frustum Gaussians + integrated positional encoding, fused (mip.py:9-43,
// 154-199), and the per-ray direction encoding (nerf_helpers.py:552-575).  The reference
// materialises means/covs/[N*S,63] in HBM; here one thread turns (t0,t1,o,d) straight into the
// encoded row.
#include "common.cuh"

namespace nvsr {

constexpr int kIpeMaxFreqs = 16;

struct IpeRow {
  float mean[3], cov[3];
};

__device__ __forceinline__ IpeRow ipe_gaussian(float t0, float t1, const float o[3], const float d[3], float radius) {
  // conical_frustum_to_gaussian (mip.py:21-29); python-double scalars are cast to fp32 by torch
  // every op separately rounded (no FMA contraction): sin(2^i * mean) amplifies last-ulp differences
  float mu = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  float hw = __fmul_rn(__fsub_rn(t1, t0), 0.5f);
  float mu2 = __fmul_rn(mu, mu), hw2 = __fmul_rn(hw, hw), hw4 = __fmul_rn(hw2, hw2);  // hw**4 == (hw^2)^2 in torch.pow
  float den = __fadd_rn(__fmul_rn(3.f, mu2), hw2);
  float t_mean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, mu), hw2), den));
  float t_var = __fsub_rn(__fdiv_rn(hw2, 3.f),
                          __fmul_rn((float)(4.0 / 15.0),
                                    __fdiv_rn(__fmul_rn(hw4, __fsub_rn(__fmul_rn(12.f, mu2), hw2)), __fmul_rn(den, den))));
  float r_var = __fmul_rn(__fmul_rn(radius, radius),
                          __fsub_rn(__fadd_rn(__fdiv_rn(mu2, 4.f), __fmul_rn((float)(5.0 / 12.0), hw2)),
                                    __fdiv_rn(__fmul_rn((float)(4.0 / 15.0), hw4), den)));
  // lift_gaussian (mip.py:32-43)
  float dmag = fmaxf(1e-10f, __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  IpeRow g;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float dd = __fmul_rn(d[c], d[c]);
    g.mean[c] = __fadd_rn(__fmul_rn(d[c], t_mean), o[c]);
    g.cov[c] = __fadd_rn(__fmul_rn(t_var, dd), __fmul_rn(r_var, __fsub_rn(1.f, __fdiv_rn(dd, dmag))));
  }
  return g;
}

// out index: [sin block: i*3+c | shifted block: 3*nf + i*3+c]
__device__ __forceinline__ float ipe_value(const IpeRow& g, int nf, int j) {
  int blk = j >= 3 * nf;
  int jj = j - blk * 3 * nf;
  int i = jj / 3, c = jj - i * 3;
  float sc = (float)(1 << i);
  float y = __fmul_rn(g.mean[c], sc);
  float yv = __fmul_rn(g.cov[c], sc * sc);
  if (blk) y = __fadd_rn(y, (float)(0.5 * 3.14159265358979323846));
  return __fmul_rn(expf(__fmul_rn(-0.5f, yv)), sinf(y));
}

__global__ void ipe_rowmajor_kernel(const float* __restrict__ z, const float* __restrict__ ro, const float* __restrict__ rd,
                                    int64_t n_rays, int S, float radius, int nf, float* __restrict__ out) {
  const int D = 6 * nf;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * S * D) return;
  int64_t row = idx / D;
  int j = (int)(idx - row * D);
  int64_t ray = row / S;
  int s = (int)(row - ray * S);
  float o[3] = {__ldg(ro + ray * 3), __ldg(ro + ray * 3 + 1), __ldg(ro + ray * 3 + 2)};
  float d[3] = {__ldg(rd + ray * 3), __ldg(rd + ray * 3 + 1), __ldg(rd + ray * 3 + 2)};
  IpeRow g = ipe_gaussian(__ldg(z + ray * (S + 1) + s), __ldg(z + ray * (S + 1) + s + 1), o, d, radius);
  out[idx] = ipe_value(g, nf, j);
}

// 16-bit tile image [tile][k_pad/8][128][8]: one thread per (row, 8-column chunk)
template <bool F16>
__global__ void ipe_tile_kernel(const float* __restrict__ z, const float* __restrict__ ro, const float* __restrict__ rd,
                                int64_t n_rays, int S, float radius, int nf, int k_pad, uint4* __restrict__ out,
                                int64_t n_tiles) {
  const int chunks = k_pad / 8;
  const int D = 6 * nf;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tiles * chunks * kTileRows) return;
  int r = (int)(idx % kTileRows);
  int c = (int)((idx / kTileRows) % chunks);
  int64_t tile = idx / ((int64_t)kTileRows * chunks);
  int64_t row = tile * kTileRows + r;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (row < n_rays * S) {
    int64_t ray = row / S;
    int s = (int)(row - ray * S);
    float o[3] = {__ldg(ro + ray * 3), __ldg(ro + ray * 3 + 1), __ldg(ro + ray * 3 + 2)};
    float d[3] = {__ldg(rd + ray * 3), __ldg(rd + ray * 3 + 1), __ldg(rd + ray * 3 + 2)};
    IpeRow g = ipe_gaussian(__ldg(z + ray * (S + 1) + s), __ldg(z + ray * (S + 1) + s + 1), o, d, radius);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int j = c * 8 + e;
      if (j < D) f[e] = ipe_value(g, nf, j);
    }
  }
  uint4 o4;
  o4.x = pack16x2<F16>(f[0], f[1]), o4.y = pack16x2<F16>(f[2], f[3]);
  o4.z = pack16x2<F16>(f[4], f[5]), o4.w = pack16x2<F16>(f[6], f[7]);
  out[idx] = o4;  // idx == (tile*chunks + c)*128 + r
}

// 16-bit tile image, the six-frequency encoding of the reference's mip baseline (multires = 7 -> 36 columns,
// k_pad = 48): ONE thread per row.  The Gaussian is computed once per row (the per-chunk kernel above repeats
// its eight IEEE divisions for every 8-column chunk), the exponential once per (frequency, axis) for both the
// sin and the phase-shifted block, and sin / cos of the base frequency once per axis with the precise
// sincosf; the higher frequencies follow by the double-angle recurrence (the abs error doubles per octave:
// < 4e-6 after five, against the 16-bit rounding of 5e-4 / 4e-3 of the value).  The reference evaluates
// sin(fl(y + pi/2)) for the shifted block, which differs from cos(y) by the rounding of the sum (< 8e-6 at
// |y| ~ 200): also below the 16-bit rounding.  Stores: a warp's 32 rows of one chunk are 512 contiguous bytes.
template <bool F16>
__global__ void __launch_bounds__(256)
ipe_tile6_kernel(const float* __restrict__ z, const float* __restrict__ ro, const float* __restrict__ rd,
                 int64_t n_rays, int S, float radius, uint4* __restrict__ out, int64_t n_tiles) {
  constexpr int NF = 6, CHUNKS = 6;  // 36 values + 12 zero columns = 48 = 6 chunks of 8
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tiles * kTileRows) return;
  const int r = (int)(idx % kTileRows);
  const int64_t tile = idx / kTileRows;
  const int64_t row = idx;
  float f[48];
#pragma unroll
  for (int j = 0; j < 48; ++j) f[j] = 0.f;
  if (row < n_rays * S) {
    int64_t ray = row / S;
    int s = (int)(row - ray * S);
    float o[3] = {__ldg(ro + ray * 3), __ldg(ro + ray * 3 + 1), __ldg(ro + ray * 3 + 2)};
    float d[3] = {__ldg(rd + ray * 3), __ldg(rd + ray * 3 + 1), __ldg(rd + ray * 3 + 2)};
    IpeRow g = ipe_gaussian(__ldg(z + ray * (S + 1) + s), __ldg(z + ray * (S + 1) + s + 1), o, d, radius);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(g.mean[c], &sn, &cs);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const float sc2 = (float)(1 << (2 * i));
        const float e = expf(__fmul_rn(-0.5f, __fmul_rn(g.cov[c], sc2)));
        f[i * 3 + c] = e * sn;            // sin block
        f[3 * NF + i * 3 + c] = e * cs;   // shifted block: sin(y + pi/2)
        const float s2 = 2.f * sn * cs, c2 = fmaf(-2.f * sn, sn, 1.f);
        sn = s2, cs = c2;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    uint4 o4;
    o4.x = pack16x2<F16>(f[c * 8 + 0], f[c * 8 + 1]), o4.y = pack16x2<F16>(f[c * 8 + 2], f[c * 8 + 3]);
    o4.z = pack16x2<F16>(f[c * 8 + 4], f[c * 8 + 5]), o4.w = pack16x2<F16>(f[c * 8 + 6], f[c * 8 + 7]);
    out[(tile * CHUNKS + c) * kTileRows + r] = o4;
  }
}

// Stand-alone stages with the reference's own tensors at the boundary (SURVEY.md §8b lists both as same-signature
// drop-ins): cast_rays (mip.py:9-18) -> means, covs [n,S,3]; IntegratedPositionalEncoding.forward((means, covs))
// (mip.py:164-191) -> [rows, 6*nf].  Same per-element arithmetic as the fused kernels above.
__global__ void cast_rays_kernel(const float* __restrict__ z, const float* __restrict__ ro, const float* __restrict__ rd,
                                 const float* __restrict__ radii, float radius, int64_t n_rays, int S,
                                 float* __restrict__ means, float* __restrict__ covs) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rays * S) return;
  int64_t ray = row / S;
  int s = (int)(row - ray * S);
  float o[3] = {__ldg(ro + ray * 3), __ldg(ro + ray * 3 + 1), __ldg(ro + ray * 3 + 2)};
  float d[3] = {__ldg(rd + ray * 3), __ldg(rd + ray * 3 + 1), __ldg(rd + ray * 3 + 2)};
  IpeRow g = ipe_gaussian(__ldg(z + ray * (S + 1) + s), __ldg(z + ray * (S + 1) + s + 1), o, d,
                          radii ? __ldg(radii + ray) : radius);
#pragma unroll
  for (int c = 0; c < 3; ++c) means[row * 3 + c] = g.mean[c], covs[row * 3 + c] = g.cov[c];
}

__global__ void ipe_encode_kernel(const float* __restrict__ means, const float* __restrict__ covs, int64_t rows, int nf,
                                  float* __restrict__ out) {
  const int D = 6 * nf;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * D) return;
  int64_t row = idx / D;
  IpeRow g;
#pragma unroll
  for (int c = 0; c < 3; ++c) g.mean[c] = __ldg(means + row * 3 + c), g.cov[c] = __ldg(covs + row * 3 + c);
  out[idx] = ipe_value(g, nf, (int)(idx - row * D));
}

// positional_encoding(d, nf, include_input): [d, sin(2^0 d), cos(2^0 d), sin(2^1 d), ...]
__global__ void dir_encoding_kernel(const float* __restrict__ dirs, int64_t n, int nf, int include_input,
                                    float* __restrict__ out) {
  const int D = (include_input ? 3 : 0) + 6 * nf;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * D) return;
  int64_t ray = idx / D;
  int j = (int)(idx - ray * D);
  float v;
  if (include_input && j < 3) {
    v = __ldg(dirs + ray * 3 + j);
  } else {
    int jj = j - (include_input ? 3 : 0);
    int i = jj / 6, rem = jj - i * 6;
    int c = rem % 3;
    float x = (float)(1 << i) * __ldg(dirs + ray * 3 + c);
    v = rem < 3 ? sinf(x) : cosf(x);
  }
  out[idx] = v;
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_ipe(const float* z, const float* ro, const float* rd, int64_t n_rays, int32_t n_intervals,
                            float radius, int32_t n_freqs, int32_t out_layout, int32_t k_pad, void* out,
                            void* stream) {
  NVSR_CHECK_ARG(z && ro && rd && out && n_rays >= 0 && n_intervals > 0 && n_freqs > 0 && n_freqs <= kIpeMaxFreqs);
  if (n_rays == 0) return NVSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_layout == NVSR_FEAT_ROWMAJOR_F32) {
    int64_t total = n_rays * n_intervals * 6 * n_freqs;
    int64_t blocks = ceil_div64(total, 256);
    NVSR_CHECK_ARG(blocks < ((int64_t)1 << 31));
    ipe_rowmajor_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ro, rd, n_rays, n_intervals, radius, n_freqs, (float*)out);
    NVSR_RETURN_LAST_ERROR();
  }
  if (out_layout == NVSR_FEAT_TILE_BF16 || out_layout == NVSR_FEAT_TILE_F16) {
    NVSR_CHECK_ARG(k_pad >= 6 * n_freqs && k_pad % 16 == 0);
    if (!aligned16(out)) return NVSR_ERR_ALIGNMENT;
    int64_t n_tiles = ceil_div64(n_rays * n_intervals, kTileRows);
    if (n_freqs == 6 && k_pad == 48) {  // the reference's mip baseline: one thread per row
      int64_t blocks6 = ceil_div64(n_tiles * kTileRows, 256);
      NVSR_CHECK_ARG(blocks6 < ((int64_t)1 << 31));
      if (out_layout == NVSR_FEAT_TILE_F16)
        ipe_tile6_kernel<true><<<(unsigned)blocks6, 256, 0, st>>>(z, ro, rd, n_rays, n_intervals, radius, (uint4*)out, n_tiles);
      else
        ipe_tile6_kernel<false><<<(unsigned)blocks6, 256, 0, st>>>(z, ro, rd, n_rays, n_intervals, radius, (uint4*)out, n_tiles);
      NVSR_RETURN_LAST_ERROR();
    }
    int64_t total = n_tiles * (k_pad / 8) * kTileRows;
    int64_t blocks = ceil_div64(total, 256);
    NVSR_CHECK_ARG(blocks < ((int64_t)1 << 31));
    if (out_layout == NVSR_FEAT_TILE_F16)
      ipe_tile_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(z, ro, rd, n_rays, n_intervals, radius, n_freqs, k_pad,
                                                             (uint4*)out, n_tiles);
    else
      ipe_tile_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(z, ro, rd, n_rays, n_intervals, radius, n_freqs, k_pad,
                                                              (uint4*)out, n_tiles);
    NVSR_RETURN_LAST_ERROR();
  }
  return NVSR_ERR_UNSUPPORTED;
}

extern "C" int32_t nvsr_cast_rays(const float* z, const float* ro, const float* rd, const float* radii, float radius,
                                  int64_t n_rays, int32_t n_intervals, float* means, float* covs, void* stream) {
  NVSR_CHECK_ARG(z && ro && rd && means && covs && n_rays >= 0 && n_intervals > 0);
  if (n_rays == 0) return NVSR_OK;
  int64_t blocks = ceil_div64(n_rays * n_intervals, 256);
  NVSR_CHECK_ARG(blocks < ((int64_t)1 << 31));
  cast_rays_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, ro, rd, radii, radius, n_rays, n_intervals,
                                                                       means, covs);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_ipe_encode(const float* means, const float* covs, int64_t rows, int32_t n_freqs, float* out,
                                   void* stream) {
  NVSR_CHECK_ARG(means && covs && out && rows >= 0 && n_freqs > 0 && n_freqs <= kIpeMaxFreqs);
  if (rows == 0) return NVSR_OK;
  int64_t blocks = ceil_div64(rows * 6 * n_freqs, 256);
  NVSR_CHECK_ARG(blocks < ((int64_t)1 << 31));
  ipe_encode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(means, covs, rows, n_freqs, out);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_dir_encoding(const float* dirs, int64_t n_rays, int32_t n_freqs, int32_t include_input,
                                     float* out, void* stream) {
  NVSR_CHECK_ARG(dirs && out && n_rays >= 0 && n_freqs >= 0 && n_freqs <= kIpeMaxFreqs);
  if (n_rays == 0) return NVSR_OK;
  int D = (include_input ? 3 : 0) + 6 * n_freqs;
  int64_t total = n_rays * D;
  dir_encoding_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(dirs, n_rays, n_freqs,
                                                                                         include_input, out);
  NVSR_RETURN_LAST_ERROR();
}
