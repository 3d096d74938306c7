// a4+a5: fused stratified sampler + tri-plane bilinear gather.
//
// Two kernels:
//   gather_rowmajor_f32 : fp32 planes -> fp32 row-major features (parity mode, feeds the SIMT decoder)
//   gather_tile_16      : bf16|fp16 channels-last planes -> 16-bit "tile image" features, which the tcgen05
//                         decoder fetches with a single bulk copy per 128-row tile.  One thread per row,
//                         corners/weights in registers, 12 texel loads in flight per chunk, packed-fp32
//                         (FFMA2) interpolation, coalesced streaming stores; no shared memory.  Rows are in the
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
  int combine_sum;  // combine_pos_planes: 0 = 'avg', 1 = 'sum' (models.py:355-361)
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
// fp32 parity kernel: one thread per (row, CPT consecutive 4-channel chunks).  CPT = 4 (C % 16 == 0) computes the row's
// depth / point / footprints once per 16 channels instead of once per 4 — the per-element arithmetic is the same.
template <int CPT>
__global__ void __launch_bounds__(256)
gather_rowmajor_f32(SamplerArgs a, PlaneArgs p, float* __restrict__ featP, float* __restrict__ featM,
                    float* __restrict__ z_out) {
  const int chunks = p.C / (4 * CPT);
  int64_t rows = a.n_rays * a.S;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * chunks) return;
  int64_t row = idx / chunks;
  const int ch0 = (int)(idx % chunks) * 4 * CPT;
  int64_t ray = row / a.S;
  int s = (int)(row % a.S);
  float z = sample_depth(a, ray, s);
  if (z_out && ch0 == 0) z_out[row] = z;
  Bilin b[3];
  sample_corners(a, p, ray, z, b);
  float4 m[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) m[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float* pl = (const float*)p.plane[d];
    int rw = p.rw[d];
    const float* p00 = pl + ((int64_t)b[d].y0 * rw + b[d].x0) * p.C + ch0;
    const float* p01 = pl + ((int64_t)b[d].y0 * rw + b[d].x1) * p.C + ch0;
    const float* p10 = pl + ((int64_t)b[d].y1 * rw + b[d].x0) * p.C + ch0;
    const float* p11 = pl + ((int64_t)b[d].y1 * rw + b[d].x1) * p.C + ch0;
    float4 v00[CPT], v01[CPT], v10[CPT], v11[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      v00[c] = __ldg(reinterpret_cast<const float4*>(p00) + c), v01[c] = __ldg(reinterpret_cast<const float4*>(p01) + c);
      v10[c] = __ldg(reinterpret_cast<const float4*>(p10) + c), v11[c] = __ldg(reinterpret_cast<const float4*>(p11) + c);
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float4 o;
      // ATen order: nw*w + ne*w + sw*w + se*w, unfused
      o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00[c].x, b[d].w00), __fmul_rn(v01[c].x, b[d].w01)), __fmul_rn(v10[c].x, b[d].w10)), __fmul_rn(v11[c].x, b[d].w11));
      o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00[c].y, b[d].w00), __fmul_rn(v01[c].y, b[d].w01)), __fmul_rn(v10[c].y, b[d].w10)), __fmul_rn(v11[c].y, b[d].w11));
      o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00[c].z, b[d].w00), __fmul_rn(v01[c].z, b[d].w01)), __fmul_rn(v10[c].z, b[d].w10)), __fmul_rn(v11[c].z, b[d].w11));
      o.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00[c].w, b[d].w00), __fmul_rn(v01[c].w, b[d].w01)), __fmul_rn(v10[c].w, b[d].w10)), __fmul_rn(v11[c].w, b[d].w11));
      if (featP) *reinterpret_cast<float4*>(featP + row * (3 * p.C) + d * p.C + ch0 + 4 * c) = o;
      m[c].x = __fadd_rn(m[c].x, o.x), m[c].y = __fadd_rn(m[c].y, o.y), m[c].z = __fadd_rn(m[c].z, o.z), m[c].w = __fadd_rn(m[c].w, o.w);
    }
  }
  // combine_pos_planes('avg'): stack(...).mean(0) = sum / 3; 'sum': the sum itself
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (!p.combine_sum)
      m[c].x = __fdiv_rn(m[c].x, 3.f), m[c].y = __fdiv_rn(m[c].y, 3.f), m[c].z = __fdiv_rn(m[c].z, 3.f), m[c].w = __fdiv_rn(m[c].w, 3.f);
    *reinterpret_cast<float4*>(featM + row * p.C + ch0 + 4 * c) = m[c];
  }
}

// ------------------------------------------------------------------------------------------------
// 16-bit tile kernel: ONE THREAD PER ROW of the 128-row tile (lane = row, so consecutive lanes are
// adjacent rays at one depth and every shared-memory store of a warp covers 512 contiguous bytes of a
// K-chunk: conflict-free).  The row's 12 texel offsets / bilinear weights live in registers; the thread
// walks the C/8 16-byte channel chunks, issuing the 12 texel loads of a chunk back to back and
// interpolating with packed fp32 FMAs (FFMA2).
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
template <bool F16>
__device__ __forceinline__ unsigned long long unpack16x2_pair(uint32_t v) {
  float2 f = unpack16x2<F16>(v);
  return pack_f32x2(f.x, f.y);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16_pair(unsigned long long v) {
  return pack16x2<F16>(__uint_as_float((uint32_t)v), __uint_as_float((uint32_t)(v >> 32)));
}

constexpr int kGatherThreads = kTileRows;  // one thread per row
#ifndef NVSR_GATHER_MINB
#define NVSR_GATHER_MINB 5  // resident CTAs per SM the register budget is set for (A/B measured)
#endif

// 16-byte streaming store (the feature tile is written once and read by the next kernel: keep it out of L1,
// which the texel reads need)
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Per row and plane: the two plane rows (y0, y1) of the footprint at its left column x0, and the four
// weights.  16-bit planes are "x-pair records" [Rh][C/8][Rw][2][8] (rays.cu pack_plane16_kernel): the
// 32-byte record (y, c, x) holds the 8-channel chunk c of texel (y, x) followed by the same chunk of its
// right neighbour (y, min(x+1, Rw-1)), so ONE 256-bit load fetches both x corners of a footprint row —
// half the load instructions and L1 data-pipe wavefronts of separate 16-byte corner loads — and the
// footprints of the adjacent rays of a quarter warp fall into one or two 128-byte lines.
struct Foot {
  uint32_t top, bot;      // record offsets of (y0, x0), (y1, x0), chunk 0
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Foot make_foot(const Bilin& b, int rw, int CH) {
  Foot f;
  f.top = (uint32_t)((b.y0 * CH) * rw + b.x0);
  f.bot = (uint32_t)((b.y1 * CH) * rw + b.x0);
  f.w00 = b.w00, f.w01 = b.w01, f.w10 = b.w10, f.w11 = b.w11;
  return f;
}
struct alignas(32) XPair {
  uint4 l, r;  // chunk of texel x0, chunk of texel x0+1
};
// 256-bit read-only load (LDG.E.256, sm_100+)
__device__ __forceinline__ XPair ldg256(const XPair* p) {
  XPair v;
  asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v.l.x), "=r"(v.l.y), "=r"(v.l.z), "=r"(v.l.w), "=r"(v.r.x), "=r"(v.r.y), "=r"(v.r.z), "=r"(v.r.w)
      : "l"(p));
  return v;
}

// acc[e] (+)= unpack(t) * w for the 4 channel pairs of one 16-byte texel chunk
template <bool F16, bool FIRST>
__device__ __forceinline__ void texel_fma(unsigned long long acc[4], const uint4& t, float wgt) {
  const unsigned long long w = pack_f32x2(wgt, wgt);
  if constexpr (FIRST) {
    acc[0] = mul_f32x2(unpack16x2_pair<F16>(t.x), w), acc[1] = mul_f32x2(unpack16x2_pair<F16>(t.y), w);
    acc[2] = mul_f32x2(unpack16x2_pair<F16>(t.z), w), acc[3] = mul_f32x2(unpack16x2_pair<F16>(t.w), w);
  } else {
    acc[0] = fma_f32x2(unpack16x2_pair<F16>(t.x), w, acc[0]), acc[1] = fma_f32x2(unpack16x2_pair<F16>(t.y), w, acc[1]);
    acc[2] = fma_f32x2(unpack16x2_pair<F16>(t.z), w, acc[2]), acc[3] = fma_f32x2(unpack16x2_pair<F16>(t.w), w, acc[3]);
  }
}

// No shared memory: in the tile image [K/8][128 rows][16 B] a warp's 32 rows of one chunk are 512
// contiguous bytes, so the stores go straight to global, fully coalesced.  CH_T > 0 fixes the chunk
// count at compile time (6 for the reference's 48-channel planes): the chunk loop is fully unrolled
// and the loads of the next chunk are in flight while one is interpolated.
// HILO (the fp16-split precision mode): every plane comes as TWO fp16 x-pair images, hi = fp16(p) and lo = fp16(p - hi),
// so hi + lo carries ~22 bits of the fp32 plane.  featP (the colour chain's fp16 operand) is interpolated from hi alone,
// exactly as without HILO; the combined features are interpolated from hi + lo in fp32 and written as an fp32 tile image
// [tiles][C/4][128 rows][4] (featM32) — what the split density chain splits into its own hi / lo operands.
struct LoPlanes {
  const void* plane[3];
};
template <bool F16, int CH_T, bool HILO = false>
__global__ void __launch_bounds__(kGatherThreads, HILO ? 4 : NVSR_GATHER_MINB)
gather_tile_16(SamplerArgs a, PlaneArgs p, uint8_t* __restrict__ featP, uint8_t* __restrict__ featM,
               float* __restrict__ z_out, int64_t n_tiles, LoPlanes lo = LoPlanes{}, float* __restrict__ featM32 = nullptr) {
  const int CH = CH_T > 0 ? CH_T : p.C / 8;  // 16-byte chunks per plane texel (6 for C=48)
  const uint32_t p_bytes = 3u * CH * 2048u;  // 128 rows * 3C * 2 B
  const uint32_t m_bytes = (uint32_t)CH * 2048u;
  const int TS = tiles_per_block(a.S);
  const int r = threadIdx.x;
  const float comb = p.combine_sum ? 1.f : 1.f / 3.f;
  const unsigned long long third = pack_f32x2(comb, comb);
  const XPair* const pl0 = reinterpret_cast<const XPair*>(p.plane[0]);
  const XPair* const pl1 = reinterpret_cast<const XPair*>(p.plane[1]);
  const XPair* const pl2 = reinterpret_cast<const XPair*>(p.plane[2]);
  const XPair* const lo0 = reinterpret_cast<const XPair*>(lo.plane[0]);
  const XPair* const lo1 = reinterpret_cast<const XPair*>(lo.plane[1]);
  const XPair* const lo2 = reinterpret_cast<const XPair*>(lo.plane[2]);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- per-row sample position -> 3 planes' footprints (registers) ----
    int64_t ray;
    int s;
    blocked_decode(tile, r, TS, &ray, &s);
    Foot f[3];
    if (ray < a.n_rays && s < a.S) {
      float z = sample_depth(a, ray, s);
      if (z_out) z_out[ray * a.S + s] = z;
      Bilin b[3];
      sample_corners(a, p, ray, z, b);
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = make_foot(b[d], p.rw[d], CH);
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = Foot{0u, 0u, 0.f, 0.f, 0.f, 0.f};  // padded rows -> zeros
    }
    uint8_t* gP = featP + tile * (int64_t)p_bytes + (uint32_t)r * 16u;
    uint8_t* gM = featM + tile * (int64_t)m_bytes + (uint32_t)r * 16u;
    const XPair* rowT[3] = {pl0 + f[0].top, pl1 + f[1].top, pl2 + f[2].top};
    const XPair* rowB[3] = {pl0 + f[0].bot, pl1 + f[1].bot, pl2 + f[2].bot};
    const XPair* loT[3] = {lo0 + f[0].top, lo1 + f[1].top, lo2 + f[2].top};
    const XPair* loB[3] = {lo0 + f[0].bot, lo1 + f[1].bot, lo2 + f[2].bot};
    uint8_t* gM32 = HILO ? reinterpret_cast<uint8_t*>(featM32) + tile * (int64_t)(2u * m_bytes) + (uint32_t)r * 16u : nullptr;
    // ---- channel chunks ----
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      unsigned long long mean[4];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const XPair vt = ldg256(rowT[d] + c * p.rw[d]);
        const XPair vb = ldg256(rowB[d] + c * p.rw[d]);
        unsigned long long acc[4];
        // ATen order: nw*w + ne*w + sw*w + se*w
        texel_fma<F16, true>(acc, vt.l, f[d].w00);
        texel_fma<F16, false>(acc, vt.r, f[d].w01);
        texel_fma<F16, false>(acc, vb.l, f[d].w10);
        texel_fma<F16, false>(acc, vb.r, f[d].w11);
        if (featP) {  // featP == NULL: density features only
          uint4 o;
          o.x = pack16_pair<F16>(acc[0]), o.y = pack16_pair<F16>(acc[1]);
          o.z = pack16_pair<F16>(acc[2]), o.w = pack16_pair<F16>(acc[3]);
          st_stream16(gP + (uint32_t)(d * CH + c) * 2048u, o);
        }
        if constexpr (HILO) {   // the low halves of the same four texels, on top of the high halves' sum
          const XPair wt = ldg256(loT[d] + c * p.rw[d]);
          const XPair wb = ldg256(loB[d] + c * p.rw[d]);
          texel_fma<F16, false>(acc, wt.l, f[d].w00);
          texel_fma<F16, false>(acc, wt.r, f[d].w01);
          texel_fma<F16, false>(acc, wb.l, f[d].w10);
          texel_fma<F16, false>(acc, wb.r, f[d].w11);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) mean[e] = d == 0 ? acc[e] : add_f32x2(mean[e], acc[e]);
      }
      if constexpr (HILO) {
        uint4 o0, o1;   // 8 fp32 channels = two 16-byte groups of the fp32 tile image
        const unsigned long long m0 = mul_f32x2(mean[0], third), m1 = mul_f32x2(mean[1], third);
        const unsigned long long m2 = mul_f32x2(mean[2], third), m3 = mul_f32x2(mean[3], third);
        o0.x = (uint32_t)m0, o0.y = (uint32_t)(m0 >> 32), o0.z = (uint32_t)m1, o0.w = (uint32_t)(m1 >> 32);
        o1.x = (uint32_t)m2, o1.y = (uint32_t)(m2 >> 32), o1.z = (uint32_t)m3, o1.w = (uint32_t)(m3 >> 32);
        st_stream16(gM32 + (uint32_t)(2 * c) * 2048u, o0);
        st_stream16(gM32 + (uint32_t)(2 * c + 1) * 2048u, o1);
        if (featM) {   // optionally the same combined features rounded to fp16 (the training decoder's operand image)
          uint4 o;
          o.x = pack16_pair<F16>(m0), o.y = pack16_pair<F16>(m1), o.z = pack16_pair<F16>(m2), o.w = pack16_pair<F16>(m3);
          st_stream16(gM + (uint32_t)c * 2048u, o);
        }
      } else {
        uint4 o;
        o.x = pack16_pair<F16>(mul_f32x2(mean[0], third)), o.y = pack16_pair<F16>(mul_f32x2(mean[1], third));
        o.z = pack16_pair<F16>(mul_f32x2(mean[2], third)), o.w = pack16_pair<F16>(mul_f32x2(mean[3], third));
        st_stream16(gM + (uint32_t)c * 2048u, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sparse colour path.  A sample whose density (+ noise) is <= 0 has alpha = 0 and weight exactly 0: its colour
// never reaches any output (volume_rendering_utils.py:29-44), so the rgb decoder need not see it.  After the
// density chain, keep_rows_kernel lists the rows that CAN contribute; gather_rows_16 then writes the 3-plane
// features of exactly those rows, densely packed into tile images in list order, for the rgb decoder.
// Rows are identified by their BLOCKED row id (tile * 128 + row-in-tile); the order of the list is irrelevant to
// the result (rows are independent), so it is built with warp-aggregated atomics.
__global__ void __launch_bounds__(256)
keep_rows_kernel(const float* __restrict__ sigma, const float* __restrict__ noise, int64_t n_rays, int S, int64_t n_tiles,
                 int32_t* __restrict__ keep, int32_t* __restrict__ count) {
  __shared__ int warp_cnt[8];
  __shared__ int block_base;
  const int TS = tiles_per_block(S);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_rows = n_tiles * kTileRows;
  // 1024 rows per block pass (4 per thread): one atomic on the global counter per pass
  for (int64_t base = (int64_t)blockIdx.x * 1024; base < n_rows; base += (int64_t)gridDim.x * 1024) {
    unsigned m[4];
    bool k[4];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t row = base + j * 256 + threadIdx.x;
      k[j] = false;
      if (row < n_rows) {
        int64_t ray;
        int s;
        blocked_decode(row / kTileRows, (int)(row % kTileRows), TS, &ray, &s);
        if (ray < n_rays && s < S) {
          float v = __ldg(sigma + row);
          if (noise) v = __fadd_rn(v, __ldg(noise + ray * S + s));
          k[j] = !(v <= 0.f);  // NaN is kept: it must reach the maps as NaN
        }
      }
      m[j] = __ballot_sync(0xffffffffu, k[j]);
      mine += __popc(m[j]);
    }
    if (lane == 0) warp_cnt[wid] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        int c = warp_cnt[w];
        warp_cnt[w] = tot;  // exclusive prefix
        tot += c;
      }
      block_base = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    int at = block_base + warp_cnt[wid];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k[j]) keep[at + __popc(m[j] & ((1u << lane) - 1u))] = (int32_t)(base + j * 256 + threadIdx.x);
      at += __popc(m[j]);
    }
    __syncthreads();  // warp_cnt / block_base are reused by the next pass
  }
}

// featP of the listed rows, packed densely: list entry i -> tile image i / 128, row i % 128.  One thread per
// entry, same arithmetic per row as gather_tile_16.  Entries beyond *count (up to a whole tile) are zero rows.
template <bool F16, int CH_T>
__global__ void __launch_bounds__(kGatherThreads, NVSR_GATHER_MINB)
gather_rows_16(SamplerArgs a, PlaneArgs p, const int32_t* __restrict__ keep, const int32_t* __restrict__ count,
               uint8_t* __restrict__ featP) {
  const int CH = CH_T > 0 ? CH_T : p.C / 8;
  const uint32_t p_bytes = 3u * CH * 2048u;
  const int TS = tiles_per_block(a.S);
  const int r = threadIdx.x;
  const XPair* const pl0 = reinterpret_cast<const XPair*>(p.plane[0]);
  const XPair* const pl1 = reinterpret_cast<const XPair*>(p.plane[1]);
  const XPair* const pl2 = reinterpret_cast<const XPair*>(p.plane[2]);
  const int64_t n = *count;
  const int64_t n_tiles = (n + kTileRows - 1) / kTileRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * kTileRows + r;
    Foot f[3];
    if (i < n) {
      const int64_t row = keep[i];
      int64_t ray;
      int s;
      blocked_decode(row / kTileRows, (int)(row % kTileRows), TS, &ray, &s);
      const float z = sample_depth(a, ray, s);
      Bilin b[3];
      sample_corners(a, p, ray, z, b);
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = make_foot(b[d], p.rw[d], CH);
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = Foot{0u, 0u, 0.f, 0.f, 0.f, 0.f};
    }
    uint8_t* gP = featP + tile * (int64_t)p_bytes + (uint32_t)r * 16u;
    const XPair* rowT[3] = {pl0 + f[0].top, pl1 + f[1].top, pl2 + f[2].top};
    const XPair* rowB[3] = {pl0 + f[0].bot, pl1 + f[1].bot, pl2 + f[2].bot};
#pragma unroll
    for (int c = 0; c < CH; ++c) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const XPair vt = ldg256(rowT[d] + c * p.rw[d]);
        const XPair vb = ldg256(rowB[d] + c * p.rw[d]);
        unsigned long long acc[4];
        texel_fma<F16, true>(acc, vt.l, f[d].w00);
        texel_fma<F16, false>(acc, vt.r, f[d].w01);
        texel_fma<F16, false>(acc, vb.l, f[d].w10);
        texel_fma<F16, false>(acc, vb.r, f[d].w11);
        uint4 o;
        o.x = pack16_pair<F16>(acc[0]), o.y = pack16_pair<F16>(acc[1]);
        o.z = pack16_pair<F16>(acc[2]), o.w = pack16_pair<F16>(acc[3]);
        st_stream16(gP + (uint32_t)(d * CH + c) * 2048u, o);
      }
    }
  }
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_sample_gather(const nvsr_sampler_t* s, const nvsr_planes_t* pl, int32_t feat_layout,
                                      void* feat_p, void* feat_m, float* z_out, void* stream) {
  NVSR_CHECK_ARG(s && pl && feat_m);   // feat_p == NULL: only the combined features are written
  NVSR_CHECK_ARG(s->n_rays >= 0 && s->n_samples > 0 && s->ro && s->rd);
  NVSR_CHECK_ARG(s->z_in || s->t_vals);
  NVSR_CHECK_ARG(pl->channels > 0 && pl->channels % 8 == 0 && pl->channels <= 64);
  NVSR_CHECK_ARG(pl->combine == 0 || pl->combine == 1);
  for (int d = 0; d < 3; ++d) {
    NVSR_CHECK_ARG(pl->plane[d] && pl->rh[d] > 0 && pl->rw[d] > 0);
    if ((reinterpret_cast<uintptr_t>(pl->plane[d]) & (is_16bit(pl->dtype) ? 31u : 15u)) != 0) return NVSR_ERR_ALIGNMENT;
  }
  if ((feat_p && !aligned16(feat_p)) || !aligned16(feat_m)) return NVSR_ERR_ALIGNMENT;
  if (s->n_rays == 0) return NVSR_OK;

  SamplerArgs a{s->n_rays, s->n_samples, s->ro, s->rd, s->near_, s->far_, s->lindisp, s->t_vals, s->t_rand, s->z_in};
  PlaneArgs p;
  for (int d = 0; d < 3; ++d) {
    p.plane[d] = pl->plane[d], p.rh[d] = pl->rh[d], p.rw[d] = pl->rw[d];
    p.lo[d] = pl->box_lo[d], p.rng[d] = pl->box_rng[d];
    for (int j = 0; j < 6; ++j) p.proj[d][j] = pl->proj[d][j];
  }
  p.C = pl->channels;
  p.combine_sum = pl->combine == 1;
  int64_t rows = s->n_rays * (int64_t)s->n_samples;
  cudaStream_t st = (cudaStream_t)stream;

  if (feat_layout == NVSR_FEAT_ROWMAJOR_F32) {
    if (pl->dtype != NVSR_F32) return NVSR_ERR_UNSUPPORTED;
    // CPT = 4 (coordinates once per 16 channels) measured SLOWER on B200 (11.8 vs 9.4 ms per mean-only fine-pass launch:
    // 64-byte strides between the lanes of a store instruction); kept as a template parameter for the record
    const bool wide = false;
    int64_t total = rows * (p.C / (wide ? 16 : 4));
    int64_t blocks = ceil_div64(total, 256);
    NVSR_CHECK_ARG(blocks < (int64_t)1 << 31);
    if (wide) gather_rowmajor_f32<4><<<(unsigned)blocks, 256, 0, st>>>(a, p, (float*)feat_p, (float*)feat_m, z_out);
    else gather_rowmajor_f32<1><<<(unsigned)blocks, 256, 0, st>>>(a, p, (float*)feat_p, (float*)feat_m, z_out);
    NVSR_RETURN_LAST_ERROR();
  }
  if (feat_layout == NVSR_FEAT_TILE_BF16 || feat_layout == NVSR_FEAT_TILE_F16) {
    const bool f16 = feat_layout == NVSR_FEAT_TILE_F16;
    if (pl->dtype != (f16 ? NVSR_F16 : NVSR_BF16)) return NVSR_ERR_UNSUPPORTED;
    const bool c48 = pl->channels == 48;
    auto kernel = f16 ? (c48 ? gather_tile_16<true, 6> : gather_tile_16<true, 0>)
                      : (c48 ? gather_tile_16<false, 6> : gather_tile_16<false, 0>);
    int64_t n_tiles = rows_padded(s->n_rays, s->n_samples, NVSR_ROWS_BLOCKED) / kTileRows;
    int64_t grid = (int64_t)kNumSMs * NVSR_GATHER_MINB * 4;  // a few waves of the resident CTAs per SM, grid-stride beyond
    if (grid > n_tiles) grid = n_tiles;
    kernel<<<(unsigned)grid, kGatherThreads, 0, st>>>(a, p, (uint8_t*)feat_p, (uint8_t*)feat_m, z_out, n_tiles, LoPlanes{}, nullptr);
    NVSR_RETURN_LAST_ERROR();
  }
  return NVSR_ERR_UNSUPPORTED;
}

extern "C" int32_t nvsr_sample_gather_hilo(const nvsr_sampler_t* s, const nvsr_planes_t* pl, const void* const lo_plane[3],
                                           void* feat_p, float* feat_m32, void* feat_m16, float* z_out, void* stream) {
  NVSR_CHECK_ARG(s && pl && lo_plane && feat_m32);
  NVSR_CHECK_ARG(s->n_rays >= 0 && s->n_samples > 0 && s->ro && s->rd && (s->z_in || s->t_vals));
  NVSR_CHECK_ARG(pl->channels > 0 && pl->channels % 8 == 0 && pl->channels <= 64 && (pl->combine == 0 || pl->combine == 1));
  if (pl->dtype != NVSR_F16) return NVSR_ERR_UNSUPPORTED;
  LoPlanes lo;
  for (int d = 0; d < 3; ++d) {
    NVSR_CHECK_ARG(pl->plane[d] && lo_plane[d] && pl->rh[d] > 0 && pl->rw[d] > 0);
    if ((reinterpret_cast<uintptr_t>(pl->plane[d]) & 31u) != 0 || (reinterpret_cast<uintptr_t>(lo_plane[d]) & 31u) != 0)
      return NVSR_ERR_ALIGNMENT;
    lo.plane[d] = lo_plane[d];
  }
  if ((feat_p && !aligned16(feat_p)) || !aligned16(feat_m32) || (feat_m16 && !aligned16(feat_m16))) return NVSR_ERR_ALIGNMENT;
  if (s->n_rays == 0) return NVSR_OK;
  SamplerArgs a{s->n_rays, s->n_samples, s->ro, s->rd, s->near_, s->far_, s->lindisp, s->t_vals, s->t_rand, s->z_in};
  PlaneArgs p;
  for (int d = 0; d < 3; ++d) {
    p.plane[d] = pl->plane[d], p.rh[d] = pl->rh[d], p.rw[d] = pl->rw[d];
    p.lo[d] = pl->box_lo[d], p.rng[d] = pl->box_rng[d];
    for (int j = 0; j < 6; ++j) p.proj[d][j] = pl->proj[d][j];
  }
  p.C = pl->channels;
  p.combine_sum = pl->combine == 1;
  auto kernel = pl->channels == 48 ? gather_tile_16<true, 6, true> : gather_tile_16<true, 0, true>;
  int64_t n_tiles = rows_padded(s->n_rays, s->n_samples, NVSR_ROWS_BLOCKED) / kTileRows;
  int64_t grid = (int64_t)kNumSMs * 4 * 4;
  if (grid > n_tiles) grid = n_tiles;
  kernel<<<(unsigned)grid, kGatherThreads, 0, (cudaStream_t)stream>>>(a, p, (uint8_t*)feat_p, (uint8_t*)feat_m16, z_out, n_tiles, lo, feat_m32);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_keep_rows(const float* sigma, const float* noise, int64_t n_rays, int32_t n_samples,
                                  int32_t* keep_rows, int32_t* count, void* stream) {
  NVSR_CHECK_ARG(sigma && keep_rows && count && n_rays >= 0 && n_samples > 0);
  if (n_rays == 0) return NVSR_OK;
  const int64_t n_tiles = rows_padded(n_rays, n_samples, NVSR_ROWS_BLOCKED) / kTileRows;
  NVSR_CHECK_ARG(n_tiles * kTileRows < ((int64_t)1 << 31));
  int64_t blocks = ceil_div64(n_tiles * kTileRows, 1024);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  keep_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sigma, noise, n_rays, n_samples, n_tiles, keep_rows, count);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_sample_gather_rows(const nvsr_sampler_t* s, const nvsr_planes_t* pl, int32_t feat_layout,
                                           const int32_t* keep_rows, const int32_t* count, int64_t max_rows,
                                           void* feat_p, void* stream) {
  NVSR_CHECK_ARG(s && pl && keep_rows && count && feat_p && max_rows >= 0);
  NVSR_CHECK_ARG(s->n_rays >= 0 && s->n_samples > 0 && s->ro && s->rd && s->z_in);
  NVSR_CHECK_ARG(pl->channels > 0 && pl->channels % 8 == 0 && pl->channels <= 64);
  NVSR_CHECK_ARG(feat_layout == NVSR_FEAT_TILE_BF16 || feat_layout == NVSR_FEAT_TILE_F16);
  const bool f16 = feat_layout == NVSR_FEAT_TILE_F16;
  if (pl->dtype != (f16 ? NVSR_F16 : NVSR_BF16)) return NVSR_ERR_UNSUPPORTED;
  for (int d = 0; d < 3; ++d) {
    NVSR_CHECK_ARG(pl->plane[d] && pl->rh[d] > 0 && pl->rw[d] > 0);
    if ((reinterpret_cast<uintptr_t>(pl->plane[d]) & 31u) != 0) return NVSR_ERR_ALIGNMENT;
  }
  if (!aligned16(feat_p)) return NVSR_ERR_ALIGNMENT;
  if (s->n_rays == 0 || max_rows == 0) return NVSR_OK;
  SamplerArgs a{s->n_rays, s->n_samples, s->ro, s->rd, s->near_, s->far_, s->lindisp, s->t_vals, s->t_rand, s->z_in};
  PlaneArgs p;
  for (int d = 0; d < 3; ++d) {
    p.plane[d] = pl->plane[d], p.rh[d] = pl->rh[d], p.rw[d] = pl->rw[d];
    p.lo[d] = pl->box_lo[d], p.rng[d] = pl->box_rng[d];
    for (int j = 0; j < 6; ++j) p.proj[d][j] = pl->proj[d][j];
  }
  p.C = pl->channels;
  p.combine_sum = pl->combine == 1;
  const bool c48 = pl->channels == 48;
  auto kernel = f16 ? (c48 ? gather_rows_16<true, 6> : gather_rows_16<true, 0>)
                    : (c48 ? gather_rows_16<false, 6> : gather_rows_16<false, 0>);
  int64_t grid = (int64_t)kNumSMs * NVSR_GATHER_MINB * 4;   // grid-stride over the tiles *count turns out to need
  const int64_t max_tiles = ceil_div64(max_rows, kTileRows);
  if (grid > max_tiles) grid = max_tiles;
  kernel<<<(unsigned)grid, kGatherThreads, 0, (cudaStream_t)stream>>>(a, p, keep_rows, count, (uint8_t*)feat_p);
  NVSR_RETURN_LAST_ERROR();
}
