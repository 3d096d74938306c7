// a4+a5: fused stratified sampler + tri-plane bilinear gather.
//
// Two kernels:
//   gather_rowmajor_f32 : fp32 planes -> fp32 row-major features (parity mode, feeds the SIMT decoder)
//   gather_tile_16      : bf16|fp16 channels-last planes -> 16-bit "tile image" features.  A CTA builds one
//                         128-row decoder tile in shared memory (coalesced 16-byte texel reads, six
//                         lanes per texel) and ships it with two bulk (TMA-engine) stores, so the
//                         tcgen05 decoder can fetch the tile with a single bulk copy.  Rows are in the
//                         BLOCKED order (8 adjacent rays x 16 samples per tile): consecutive rows are
//                         adjacent pixels at one depth, whose bilinear footprints overlap, so the
//                         texel requests of a warp collapse onto a few L1 lines.
// Nothing of pts / embedded ([N,S,3] / [N*S,6] in the reference) is ever materialised in HBM.
#include "bilinear.cuh"
#include "common.cuh"

namespace nvsr {

struct SamplerArgs {
  int64_t n_rays;
  int S;
  const float* ro;
  const float* rd;
  float near_, far_;
  int lindisp;
  const float* t_vals;
  const float* t_rand;
  const float* z_in;
};

struct PlaneArgs {
  const void* plane[3];
  int rh[3], rw[3];
  int C;
  float lo[3], rng[3];
  float proj[3][6];
};

// train_utils.py:95-109: depth of sample s on a ray (every op separately rounded)
__device__ __forceinline__ float coarse_z_at(const SamplerArgs& a, int s) {
  float t = __ldg(a.t_vals + s);
  if (!a.lindisp) return __fadd_rn(__fmul_rn(a.near_, __fsub_rn(1.f, t)), __fmul_rn(a.far_, t));
  float inv_n = __fdiv_rn(1.f, a.near_), inv_f = __fdiv_rn(1.f, a.far_);
  return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(inv_n, __fsub_rn(1.f, t)), __fmul_rn(inv_f, t)));
}

__device__ __forceinline__ float sample_depth(const SamplerArgs& a, int64_t ray, int s) {
  if (a.z_in) return __ldg(a.z_in + ray * a.S + s);
  float z = coarse_z_at(a, s);
  if (a.t_rand) {
    // mids = .5*(z[1:]+z[:-1]); upper = cat(mids, z[-1]); lower = cat(z[0], mids)
    float upper = (s + 1 < a.S) ? __fmul_rn(0.5f, __fadd_rn(coarse_z_at(a, s + 1), z)) : z;
    float lower = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(z, coarse_z_at(a, s - 1))) : z;
    float tr = __ldg(a.t_rand + ray * a.S + s);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
  }
  return z;
}

// pts = ro + rd*z (train_utils.py:111), box-normalise (models.py:264-265), project (models.py:497)
__device__ __forceinline__ void sample_corners(const SamplerArgs& a, const PlaneArgs& p, int64_t ray, float z,
                                               Bilin out[3]) {
  float n[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = __fadd_rn(__ldg(a.ro + ray * 3 + k), __fmul_rn(__ldg(a.rd + ray * 3 + k), z));
    n[k] = box_normalize(x, p.lo[k], p.rng[k]);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float gx = n[0] * p.proj[d][0] + n[1] * p.proj[d][2] + n[2] * p.proj[d][4];
    float gy = n[0] * p.proj[d][1] + n[1] * p.proj[d][3] + n[2] * p.proj[d][5];
    out[d] = bilinear_setup(gx, gy, p.rw[d], p.rh[d]);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 parity kernel: one thread per (row, 4-channel chunk)
__global__ void __launch_bounds__(256)
gather_rowmajor_f32(SamplerArgs a, PlaneArgs p, float* __restrict__ featP, float* __restrict__ featM,
                    float* __restrict__ z_out) {
  const int chunks = p.C / 4;
  int64_t rows = a.n_rays * a.S;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * chunks) return;
  int64_t row = idx / chunks;
  int ch = (int)(idx % chunks) * 4;
  int64_t ray = row / a.S;
  int s = (int)(row % a.S);
  float z = sample_depth(a, ray, s);
  if (z_out && ch == 0) z_out[row] = z;
  Bilin b[3];
  sample_corners(a, p, ray, z, b);
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float* pl = (const float*)p.plane[d];
    int rw = p.rw[d];
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(pl + ((int64_t)b[d].y0 * rw + b[d].x0) * p.C + ch));
    const float4 v01 = __ldg(reinterpret_cast<const float4*>(pl + ((int64_t)b[d].y0 * rw + b[d].x1) * p.C + ch));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(pl + ((int64_t)b[d].y1 * rw + b[d].x0) * p.C + ch));
    const float4 v11 = __ldg(reinterpret_cast<const float4*>(pl + ((int64_t)b[d].y1 * rw + b[d].x1) * p.C + ch));
    float4 o;
    // ATen order: nw*w + ne*w + sw*w + se*w, unfused
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00.x, b[d].w00), __fmul_rn(v01.x, b[d].w01)), __fmul_rn(v10.x, b[d].w10)), __fmul_rn(v11.x, b[d].w11));
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00.y, b[d].w00), __fmul_rn(v01.y, b[d].w01)), __fmul_rn(v10.y, b[d].w10)), __fmul_rn(v11.y, b[d].w11));
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00.z, b[d].w00), __fmul_rn(v01.z, b[d].w01)), __fmul_rn(v10.z, b[d].w10)), __fmul_rn(v11.z, b[d].w11));
    o.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00.w, b[d].w00), __fmul_rn(v01.w, b[d].w01)), __fmul_rn(v10.w, b[d].w10)), __fmul_rn(v11.w, b[d].w11));
    *reinterpret_cast<float4*>(featP + row * (3 * p.C) + d * p.C + ch) = o;
    m.x = __fadd_rn(m.x, o.x), m.y = __fadd_rn(m.y, o.y), m.z = __fadd_rn(m.z, o.z), m.w = __fadd_rn(m.w, o.w);
  }
  // combine_pos_planes('avg'): stack(...).mean(0) = sum / 3
  m.x = __fdiv_rn(m.x, 3.f), m.y = __fdiv_rn(m.y, 3.f), m.z = __fdiv_rn(m.z, 3.f), m.w = __fdiv_rn(m.w, 3.f);
  *reinterpret_cast<float4*>(featM + row * p.C + ch) = m;
}

// ------------------------------------------------------------------------------------------------
// bf16 tile kernel
struct __align__(16) RowCorners {
  int o00[3], o01[3], o10[3], o11[3];  // texel indices (y*rw + x)
  float w00[3], w01[3], w10[3], w11[3];
};

template <bool F16>
__device__ __forceinline__ void fma_16x8(float acc[8], const uint4& v, float w) {
  float2 a = unpack16x2<F16>(v.x), b = unpack16x2<F16>(v.y), c = unpack16x2<F16>(v.z), d = unpack16x2<F16>(v.w);
  acc[0] = fmaf(a.x, w, acc[0]), acc[1] = fmaf(a.y, w, acc[1]);
  acc[2] = fmaf(b.x, w, acc[2]), acc[3] = fmaf(b.y, w, acc[3]);
  acc[4] = fmaf(c.x, w, acc[4]), acc[5] = fmaf(c.y, w, acc[5]);
  acc[6] = fmaf(d.x, w, acc[6]), acc[7] = fmaf(d.y, w, acc[7]);
}

constexpr int kGatherThreads = 256;

// dynamic smem: [P image 3C/8 x 2048 B][M image C/8 x 2048 B][RowCorners x 128]
template <bool F16>
__global__ void __launch_bounds__(kGatherThreads)
gather_tile_16(SamplerArgs a, PlaneArgs p, uint8_t* __restrict__ featP, uint8_t* __restrict__ featM,
                 float* __restrict__ z_out, int64_t n_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int CH = p.C / 8;                 // 16-byte chunks per plane texel (6 for C=48)
  const uint32_t p_bytes = 3u * CH * 2048u;  // 128 rows * 3C * 2 B
  const uint32_t m_bytes = (uint32_t)CH * 2048u;
  uint8_t* sP = smem;
  uint8_t* sM = smem + p_bytes;
  RowCorners* sc = reinterpret_cast<RowCorners*>(smem + p_bytes + m_bytes);
  const int TS = tiles_per_block(a.S);
  const int tid = threadIdx.x;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // previous tile's bulk stores must have finished READING smem before we overwrite it
    if (tid == 0) bulk_wait_read<0>();
    __syncthreads();
    // phase 1: per-row sample position -> 3 planes' corner indices and weights
    if (tid < kTileRows) {
      int64_t ray;
      int s;
      blocked_decode(tile, tid, TS, &ray, &s);
      RowCorners rc;
      if (ray < a.n_rays && s < a.S) {
        float z = sample_depth(a, ray, s);
        if (z_out) z_out[ray * a.S + s] = z;
        Bilin b[3];
        sample_corners(a, p, ray, z, b);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          int rw = p.rw[d];
          rc.o00[d] = b[d].y0 * rw + b[d].x0, rc.o01[d] = b[d].y0 * rw + b[d].x1;
          rc.o10[d] = b[d].y1 * rw + b[d].x0, rc.o11[d] = b[d].y1 * rw + b[d].x1;
          rc.w00[d] = b[d].w00, rc.w01[d] = b[d].w01, rc.w10[d] = b[d].w10, rc.w11[d] = b[d].w11;
        }
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          rc.o00[d] = rc.o01[d] = rc.o10[d] = rc.o11[d] = 0;
          rc.w00[d] = rc.w01[d] = rc.w10[d] = rc.w11[d] = 0.f;  // padded rows -> zeros
        }
      }
      sc[tid] = rc;
    }
    __syncthreads();
    // phase 2: items (row, chunk); consecutive lanes take consecutive chunks of one texel
    for (int item = tid; item < kTileRows * CH; item += kGatherThreads) {
      int r = item / CH;
      int c = item - r * CH;
      const RowCorners& rc = sc[r];
      float mean[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) mean[e] = 0.f;
      uint4 v[3][4];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const uint4* pl = reinterpret_cast<const uint4*>(p.plane[d]);
        v[d][0] = __ldg(pl + (int64_t)rc.o00[d] * CH + c);
        v[d][1] = __ldg(pl + (int64_t)rc.o01[d] * CH + c);
        v[d][2] = __ldg(pl + (int64_t)rc.o10[d] * CH + c);
        v[d][3] = __ldg(pl + (int64_t)rc.o11[d] * CH + c);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        fma_16x8<F16>(acc, v[d][0], rc.w00[d]);
        fma_16x8<F16>(acc, v[d][1], rc.w01[d]);
        fma_16x8<F16>(acc, v[d][2], rc.w10[d]);
        fma_16x8<F16>(acc, v[d][3], rc.w11[d]);
#pragma unroll
        for (int e = 0; e < 8; ++e) mean[e] += acc[e];
        uint4 o;
        o.x = pack16x2<F16>(acc[0], acc[1]), o.y = pack16x2<F16>(acc[2], acc[3]);
        o.z = pack16x2<F16>(acc[4], acc[5]), o.w = pack16x2<F16>(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(sP + (uint32_t)(d * CH + c) * 2048u + (uint32_t)r * 16u) = o;
      }
      uint4 o;
      o.x = pack16x2<F16>(mean[0] / 3.f, mean[1] / 3.f), o.y = pack16x2<F16>(mean[2] / 3.f, mean[3] / 3.f);
      o.z = pack16x2<F16>(mean[4] / 3.f, mean[5] / 3.f), o.w = pack16x2<F16>(mean[6] / 3.f, mean[7] / 3.f);
      *reinterpret_cast<uint4*>(sM + (uint32_t)c * 2048u + (uint32_t)r * 16u) = o;
    }
    // phase 3: ship the two images with the bulk-copy engine
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(featP + tile * (int64_t)p_bytes, sP, p_bytes);
      bulk_s2g(featM + tile * (int64_t)m_bytes, sM, m_bytes);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait<0>();
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_sample_gather(const nvsr_sampler_t* s, const nvsr_planes_t* pl, int32_t feat_layout,
                                      void* feat_p, void* feat_m, float* z_out, void* stream) {
  NVSR_CHECK_ARG(s && pl && feat_p && feat_m);
  NVSR_CHECK_ARG(s->n_rays >= 0 && s->n_samples > 0 && s->ro && s->rd);
  NVSR_CHECK_ARG(s->z_in || s->t_vals);
  NVSR_CHECK_ARG(pl->channels > 0 && pl->channels % 8 == 0 && pl->channels <= 64);
  for (int d = 0; d < 3; ++d) {
    NVSR_CHECK_ARG(pl->plane[d] && pl->rh[d] > 0 && pl->rw[d] > 0);
    if (!aligned16(pl->plane[d])) return NVSR_ERR_ALIGNMENT;
  }
  if (!aligned16(feat_p) || !aligned16(feat_m)) return NVSR_ERR_ALIGNMENT;
  if (s->n_rays == 0) return NVSR_OK;

  SamplerArgs a{s->n_rays, s->n_samples, s->ro, s->rd, s->near_, s->far_, s->lindisp, s->t_vals, s->t_rand, s->z_in};
  PlaneArgs p;
  for (int d = 0; d < 3; ++d) {
    p.plane[d] = pl->plane[d], p.rh[d] = pl->rh[d], p.rw[d] = pl->rw[d];
    p.lo[d] = pl->box_lo[d], p.rng[d] = pl->box_rng[d];
    for (int j = 0; j < 6; ++j) p.proj[d][j] = pl->proj[d][j];
  }
  p.C = pl->channels;
  int64_t rows = s->n_rays * (int64_t)s->n_samples;
  cudaStream_t st = (cudaStream_t)stream;

  if (feat_layout == NVSR_FEAT_ROWMAJOR_F32) {
    if (pl->dtype != NVSR_F32) return NVSR_ERR_UNSUPPORTED;
    int64_t total = rows * (p.C / 4);
    int64_t blocks = ceil_div64(total, 256);
    NVSR_CHECK_ARG(blocks < (int64_t)1 << 31);
    gather_rowmajor_f32<<<(unsigned)blocks, 256, 0, st>>>(a, p, (float*)feat_p, (float*)feat_m, z_out);
    NVSR_RETURN_LAST_ERROR();
  }
  if (feat_layout == NVSR_FEAT_TILE_BF16 || feat_layout == NVSR_FEAT_TILE_F16) {
    const bool f16 = feat_layout == NVSR_FEAT_TILE_F16;
    if (pl->dtype != (f16 ? NVSR_F16 : NVSR_BF16)) return NVSR_ERR_UNSUPPORTED;
    auto kernel = f16 ? gather_tile_16<true> : gather_tile_16<false>;
    int CH = p.C / 8;
    size_t smem = (size_t)4 * CH * 2048 + sizeof(RowCorners) * kTileRows;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int32_t)e;
    int64_t n_tiles = rows_padded(s->n_rays, s->n_samples, NVSR_ROWS_BLOCKED) / kTileRows;
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) return NVSR_ERR_RESOURCE;
    int64_t grid = (int64_t)kNumSMs * ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kernel<<<(unsigned)grid, kGatherThreads, smem, st>>>(a, p, (uint8_t*)feat_p, (uint8_t*)feat_m, z_out,
                                                                  n_tiles);
    NVSR_RETURN_LAST_ERROR();
  }
  return NVSR_ERR_UNSUPPORTED;
}
