// Tensor-core decoder chain at fp32-grade accuracy: every operand is SPLIT into two fp16 terms (x = hi + lo, exact to
// 2^-22) and every layer runs three tcgen05.mma passes — hi.hi, lo.hi, hi.lo — into the same fp32 accumulator (the lo.lo
// term is below the accumulator's own rounding).  kind::tf32 would not help: tf32 carries the same 10 mantissa bits as
// fp16.  Used for the DENSITY chain of the 'fp16-split' precision mode: scripts/studies/fp16_error_sources.py shows the
// 16-bit mode's map error comes from sigma alone (the colour logits are 3e-5 off, sigma 7e-2), so splitting the density
// chain (43 % of the decoder FLOPs, x3) brings the maps inside the 1e-3 contract at ~1/30 of the SIMT fp32 mode's cost.
//
// One CTA per SM, two 128-row tiles in flight (8 epilogue warps: TMEM lane quadrant x column half, + 1 issuer warp):
//   features : fp32 row-major [n_rays * S][k0] (the fp32 gather), split by the threads and stored straight into TMEM
//              (A_hi / A_lo, the TS-form operand layout) — there is no shared-memory input ring, which is what lets BOTH
//              weight images of all four layers (221 KB for 48 -> 128 x4) stay resident in shared memory;
//   layers   : bias pre-stored into the accumulator (fp32, exact), 3 x K/16 MMAs issued by one elected lane, epilogue =
//              ReLU + split + tcgen05.st of both halves; the head (128 -> 1) is an fp32 dot product of the unrounded
//              last activations, as in the 16-bit chain.
#include "common.cuh"

namespace nvsr {

namespace {

constexpr int kSpThreads = 288;         // 8 epilogue warps (TMEM lane quadrant x column half) + 1 issuer warp
constexpr uint32_t kSpTmemCols = 512;   // slot s: D [256 s, +128) | A_hi [256 s + 128, +64) | A_lo [256 s + 192, +64)
constexpr uint32_t kSpSlotCols = 256;
constexpr uint32_t kSpAhi = 128, kSpAlo = 192;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16_kmajor(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (a, b) -> hi pair and lo pair: hi = fp16(x) (saturating), lo = fp16(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t* hi, uint32_t* lo) {
  const uint32_t h = pack16x2<true>(a, b);
  const float2 hf = unpack16x2<true>(h);
  *hi = h;
  *lo = pack16x2<true>(a - hf.x, b - hf.y);
}

struct SplitArgs {
  const float* feat;          // [n_rays * S][k0] fp32, ray-major rows; or (feat_tiled) the fp32 tile image
  int feat_tiled;             //   [tiles][k0/4][128 rows][4] in BLOCKED rows that nvsr_sample_gather_hilo writes
  int k0;
  const uint8_t* w_hi[4];     // fp16 images [k/8][128][8]
  const uint8_t* w_lo[4];
  const float* bias[4];
  const float* head_w;        // [head_n][128]
  const float* head_b;
  int head_n, head_ch;
  float* raw;                 // planar [4][raw_stride], BLOCKED rows
  int64_t raw_stride;
  int64_t n_rays, n_tiles;
  int S, tiles_per_blk;
};

// Two tiles ("slots") are in flight per CTA: while the eight epilogue warps work on one slot's accumulator, the tensor core
// runs the other slot's 3 x K/16 MMAs.  Warp 8 is the issuer: it waits for the slot's eight "operands written" arrivals,
// issues the layer's MMAs from one elected lane and commits them to the slot's mbarrier; the epilogue warps wait on that.
template <int K0C>   // k0 / 16
__global__ void __launch_bounds__(kSpThreads, 1) chain_split_kernel(const __grid_constant__ SplitArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_mma[2], bar_ready[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[4 * 128];
  __shared__ __align__(16) float s_headw[4 * 128];
  __shared__ __align__(16) float s_hpart[2 * 128 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t w0_bytes = (uint32_t)K0C * 16u * 256u, wh_bytes = 128u * 256u;
  constexpr uint32_t w_all = w0_bytes + 3u * wh_bytes;   // one copy (hi or lo) of the four images
  auto w_off = [&](int l) { return l == 0 ? 0u : w0_bytes + (uint32_t)(l - 1) * wh_bytes; };
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&bar_mma[s], 1), mbar_init(&bar_ready[s], 8);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kSpTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4 * 128; i += kSpThreads) {
    s_bias[i] = a.bias[i >> 7] ? __ldg(a.bias[i >> 7] + (i & 127)) : 0.f;
    s_headw[i] = (i >> 7) < a.head_n ? __ldg(a.head_w + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // tiles of this CTA: blockIdx.x + j * gridDim.x; slot = j & 1
  const int64_t my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 8) {
    // ================= issuer =================
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(&bar_w, 2u * w_all);
      for (int l = 0; l < 4; ++l) {
        const uint32_t nb = l == 0 ? w0_bytes : wh_bytes;
        bulk_g2s(smem + w_off(l), a.w_hi[l], nb, &bar_w);
        bulk_g2s(smem + w_all + w_off(l), a.w_lo[l], nb, &bar_w);
      }
    }
    if (my_tiles > 0) mbar_wait(&bar_w, 0);
    const uint32_t idesc = idesc_f16_kmajor(128);
    uint32_t ph[2] = {0u, 0u};
    // order of the steps: (pair p, layer l, slot s) for l = 0..3, s = 0..1 — the same order the epilogue warps follow
    for (int64_t p = 0; 2 * p < my_tiles; ++p) {
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (2 * p + s >= my_tiles) continue;
          mbar_wait(&bar_ready[s], ph[s]);
          ph[s] ^= 1u;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem + (uint32_t)s * kSpSlotCols;
            const uint64_t bh = smem_desc(smem_u32(smem + w_off(l)), 2048u, 128u);
            const uint64_t bl = smem_desc(smem_u32(smem + w_all + w_off(l)), 2048u, 128u);
            if (l == 0) {
#pragma unroll
              for (int ks = 0; ks < K0C; ++ks) {   // + 2 K-chunks (2 * 2048 B) per step of 16
                umma_ts(d, d + kSpAhi + (uint32_t)ks * 8u, bh + (uint64_t)(ks * 256), idesc, 1u);
                umma_ts(d, d + kSpAlo + (uint32_t)ks * 8u, bh + (uint64_t)(ks * 256), idesc, 1u);
                umma_ts(d, d + kSpAhi + (uint32_t)ks * 8u, bl + (uint64_t)(ks * 256), idesc, 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                umma_ts(d, d + kSpAhi + (uint32_t)ks * 8u, bh + (uint64_t)(ks * 256), idesc, 1u);
                umma_ts(d, d + kSpAlo + (uint32_t)ks * 8u, bh + (uint64_t)(ks * 256), idesc, 1u);
                umma_ts(d, d + kSpAhi + (uint32_t)ks * 8u, bl + (uint64_t)(ks * 256), idesc, 1u);
              }
            }
            umma_commit(&bar_mma[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane, col0 = half * 64;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph[2] = {0u, 0u};

    auto prestore_bias = [&](uint32_t d_tmem, int l) {
      uint32_t v[32];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + l * 128 + col0 + 32 * g + 4 * j);
          v[4 * j] = __float_as_uint(b4.x), v[4 * j + 1] = __float_as_uint(b4.y);
          v[4 * j + 2] = __float_as_uint(b4.z), v[4 * j + 3] = __float_as_uint(b4.w);
        }
        tmem_st32(d_tmem + 32u * g, v);
      }
    };
    auto arrive_ready = [&](int s) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ready[s]);
    };
    // features of `tile` -> A_hi (column-half 0 warps) / A_lo (half 1 warps) of slot s, bias of layer 0 -> D_s
    auto load_tile = [&](int s, int64_t tile) {
      const uint32_t sb = lane_base + (uint32_t)s * kSpSlotCols;
      const int64_t blk = tile / a.tiles_per_blk;
      const int smp = (int)(tile - blk * a.tiles_per_blk) * kBlkSamples + (r >> 3);
      const int64_t ray = blk * kBlkRays + (r & 7);
      const bool valid = ray < a.n_rays && smp < a.S;
      float4 f[K0C * 4];
      if (a.feat_tiled) {   // group j of row r: 16 bytes, a warp's 32 rows contiguous (padding rows are zero in the image)
        const float4* src = reinterpret_cast<const float4*>(a.feat) + tile * (int64_t)(K0C * 4 * kTileRows) + r;
#pragma unroll
        for (int j = 0; j < K0C * 4; ++j) f[j] = __ldg(src + j * kTileRows);
      } else {
        const float4* src = reinterpret_cast<const float4*>(a.feat + (valid ? (ray * a.S + smp) * (int64_t)(K0C * 16) : 0));
#pragma unroll
        for (int j = 0; j < K0C * 4; ++j) f[j] = valid ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < K0C; ++c) {   // 16 features = 8 TMEM columns per step
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split2(f[c * 4 + j].x, f[c * 4 + j].y, &hi[2 * j], &lo[2 * j]);
          split2(f[c * 4 + j].z, f[c * 4 + j].w, &hi[2 * j + 1], &lo[2 * j + 1]);
        }
        if (half == 0) tmem_st8(sb + kSpAhi + (uint32_t)(c * 8), hi);
        else tmem_st8(sb + kSpAlo + (uint32_t)(c * 8), lo);
      }
      prestore_bias(sb + (uint32_t)col0, 0);
    };

    // prologue: the first tile of each slot
    for (int s = 0; s < 2; ++s)
      if (s < my_tiles) {
        load_tile(s, blockIdx.x + (int64_t)s * gridDim.x);
        arrive_ready(s);
      }
    for (int64_t p = 0; 2 * p < my_tiles; ++p) {
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int64_t j = 2 * p + s;
          if (j >= my_tiles) continue;
          const int64_t tile = blockIdx.x + j * gridDim.x;
          const uint32_t sb = lane_base + (uint32_t)s * kSpSlotCols;
          const uint32_t d_tmem = sb + (uint32_t)col0;
          mbar_wait(&bar_mma[s], ph[s]);
          ph[s] ^= 1u;
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tmem_ld32(d_tmem, v0);
          tmem_ld32(d_tmem + 32u, v1);
          tmem_ld_wait();
          if (l < 3) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q)
              split2(fmaxf(__uint_as_float(v0[2 * q]), 0.f), fmaxf(__uint_as_float(v0[2 * q + 1]), 0.f), &hi[q], &lo[q]);
            tmem_st16(sb + kSpAhi + (uint32_t)(col0 >> 1), hi);
            tmem_st16(sb + kSpAlo + (uint32_t)(col0 >> 1), lo);
#pragma unroll
            for (int q = 0; q < 16; ++q)
              split2(fmaxf(__uint_as_float(v1[2 * q]), 0.f), fmaxf(__uint_as_float(v1[2 * q + 1]), 0.f), &hi[q], &lo[q]);
            tmem_st16(sb + kSpAhi + (uint32_t)(col0 >> 1) + 16u, hi);
            tmem_st16(sb + kSpAlo + (uint32_t)(col0 >> 1) + 16u, lo);
            prestore_bias(d_tmem, l + 1);
            arrive_ready(s);
          } else {
            float hacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (h < a.head_n) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  acc = fmaf(fmaxf(__uint_as_float(v0[c]), 0.f), s_headw[h * 128 + col0 + c], acc);
                  acc = fmaf(fmaxf(__uint_as_float(v1[c]), 0.f), s_headw[h * 128 + col0 + 32 + c], acc);
                }
                hacc[h] = acc;
              }
            }
            // the accumulator has been read: hand the slot its next tile first, so its layer-0 MMAs run under the rest
            if (j + 2 < my_tiles) {
              load_tile(s, blockIdx.x + (j + 2) * gridDim.x);
              arrive_ready(s);
            }
            float* hp = s_hpart + (s * 128 + r) * 4;
            if (half == 1) *reinterpret_cast<float4*>(hp) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
            named_bar_sync(1 + s * 4 + quad, 64);   // the two warps sharing this slot and lane quadrant
            if (half == 0) {
              const float4 o = *reinterpret_cast<const float4*>(hp);
              const float hv[4] = {hacc[0] + o.x, hacc[1] + o.y, hacc[2] + o.z, hacc[3] + o.w};
              for (int h = 0; h < a.head_n; ++h)
                a.raw[(int64_t)(a.head_ch + h) * a.raw_stride + tile * kTileRows + r] = hv[h] + __ldg(a.head_b + h);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kSpTmemCols) : "memory");
  }
}

}  // namespace
}  // namespace nvsr

using namespace nvsr;

static int32_t split_launch(const float* feat, int32_t tiled, int32_t k0, const void* const* w_hi, const void* const* w_lo,
                            const float* const* bias, const float* head_w, const float* head_b, int32_t head_n,
                            int32_t head_ch, int64_t n_rays, int32_t n_samples, float* raw, int64_t raw_stride,
                            void* stream) {
  NVSR_CHECK_ARG(feat && w_hi && w_lo && bias && head_w && head_b && raw && n_rays >= 0 && n_samples > 0);
  NVSR_CHECK_ARG(k0 >= 16 && (k0 % 16) == 0 && k0 <= 64 && head_n >= 1 && head_n <= 4 && head_ch >= 0 && head_ch + head_n <= 4);
  if (n_rays == 0) return NVSR_OK;
  SplitArgs a;
  for (int l = 0; l < 4; ++l) {
    NVSR_CHECK_ARG(w_hi[l] && w_lo[l]);
    if (!aligned16(w_hi[l]) || !aligned16(w_lo[l])) return NVSR_ERR_ALIGNMENT;
    a.w_hi[l] = (const uint8_t*)w_hi[l], a.w_lo[l] = (const uint8_t*)w_lo[l], a.bias[l] = bias[l];
  }
  if (!aligned16(feat)) return NVSR_ERR_ALIGNMENT;
  a.feat = feat, a.feat_tiled = tiled, a.k0 = k0, a.head_w = head_w, a.head_b = head_b, a.head_n = head_n, a.head_ch = head_ch;
  a.raw = raw, a.raw_stride = raw_stride, a.n_rays = n_rays, a.S = n_samples;
  a.tiles_per_blk = tiles_per_block(n_samples);
  a.n_tiles = ceil_div64(n_rays, kBlkRays) * a.tiles_per_blk;
  NVSR_CHECK_ARG(raw_stride >= a.n_tiles * kTileRows);
  const uint32_t smem_bytes = 2u * ((uint32_t)k0 * 256u + 3u * 128u * 256u);
  if (smem_bytes + 9728u > 227u * 1024u) return NVSR_ERR_RESOURCE;   // + the kernel's static shared memory
  void (*kernel)(SplitArgs) = nullptr;
  switch (k0 / 16) {
    case 1: kernel = chain_split_kernel<1>; break;
    case 2: kernel = chain_split_kernel<2>; break;
    case 3: kernel = chain_split_kernel<3>; break;
    case 4: kernel = chain_split_kernel<4>; break;
    default: return NVSR_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int32_t)e;
  const int64_t grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  kernel<<<(unsigned)grid, kSpThreads, smem_bytes, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_mlp_chain_split(const float* feat, int32_t k0, const void* const* w_hi, const void* const* w_lo,
                                        const float* const* bias, const float* head_w, const float* head_b, int32_t head_n,
                                        int32_t head_ch, int64_t n_rays, int32_t n_samples, float* raw, int64_t raw_stride,
                                        void* stream) {
  return split_launch(feat, 0, k0, w_hi, w_lo, bias, head_w, head_b, head_n, head_ch, n_rays, n_samples, raw, raw_stride, stream);
}

extern "C" int32_t nvsr_mlp_chain_split_tiled(const float* feat_tiles, int32_t k0, const void* const* w_hi,
                                              const void* const* w_lo, const float* const* bias, const float* head_w,
                                              const float* head_b, int32_t head_n, int32_t head_ch, int64_t n_rays,
                                              int32_t n_samples, float* raw, int64_t raw_stride, void* stream) {
  return split_launch(feat_tiles, 1, k0, w_hi, w_lo, bias, head_w, head_b, head_n, head_ch, n_rays, n_samples, raw, raw_stride,
                      stream);
}
