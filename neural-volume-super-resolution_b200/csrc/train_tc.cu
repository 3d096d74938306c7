// Backward of the decoder MLP on the tensor cores (the reference's loss.backward() through
// models.py:393-421, train_nerf.py:860-916).  Everything works on the FORWARD's operand images — no transposed
// packing, no activation transposes (both descriptor claims verified on a B200 by scripts/ubench/umma_wgrad.cu):
//
//   forward (training)   mlp_chain_tc_kernel<..., TRAIN> (mlp_tc.cu): the inference chain that additionally stores every
//                        layer's post-ReLU 16-bit activation tile image x_1 .. x_4 — the next MMA's exact operand.
//   dgrad_chain_kernel   per 128-row tile: g_3 = (d_out . W_head) * [x_4 > 0] on the CUDA cores, then per layer
//                        l = 3, 2, 1:  g_{l-1} = (g_l . W_l) * [x_l > 0]  as ONE tcgen05.mma K-loop with A = g_l in TMEM
//                        (TS form, the forward's activation layout) and B = the forward weight image of W_l read
//                        MN-major (8 consecutive k_in in 16 B, LBO 128 B, SBO n_out * 16 B); finally d_x0 = g_0 . W_0
//                        written fp32 row-major for nvsr_sample_gather_bwd.  Every g_l is also stored as a 16-bit tile
//                        image (the weight gradient's operand).  Deltas carry a power-of-two loss scale (fp16 deltas of
//                        an mse over thousands of rays underflow otherwise: scripts/studies/backward_precision.py).
//   wgrad_kernel         dW_l[n_out][k_in] = sum over rows g_l[r][n_out] * x_l[r][k_in]: persistent streaming kernel,
//                        both operands are tile images read MN-major (K = the 128 rows of a tile), the accumulator
//                        stays in TMEM over the CTA's whole slice of tiles; the bias gradient rides along as one more
//                        N = 16 block against a resident image of ones; partials are reduced with fp32 atomics.
//   ray_sum_kernel       per-ray sum over the samples of a delta image (the per-ray view-feature columns of the rgb
//                        chain's first layer are a per-ray bias in the forward: their gradients are tiny per-ray GEMMs).
#include "common.cuh"

namespace nvsr {

namespace {

constexpr int kDgThreads = 288;          // 8 epilogue warps (quad = warp & 3: TMEM lane quadrant, half = warp >> 2: columns) + issuer
constexpr uint32_t kDgTmemCols = 512;    // slot s: D [256 s, +192) | A [256 s + 192, +64)
constexpr uint32_t kDgSlotCols = 256;
constexpr uint32_t kDgAOff = 192;
constexpr uint32_t kActTileBytes = kTileRows * 128 * 2;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// kind::f16, fp16 x fp16 -> fp32, M = 128, N = n; a_mn / b_mn: the operand is MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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

struct DgradArgs {
  const uint8_t* w[4];   // forward weight images W_0 .. W_3, fp16 [k/8][128][8]
  int k0;                // input width of layer 0 (multiple of 16, <= 256): N of the last product
  const float* head_w;   // [head_n][128] fp32
  int head_n, head_ch;
  const float* d_raw;    // planar [4][raw_stride], BLOCKED rows: gradient w.r.t. the chain's raw outputs
  int64_t raw_stride;
  float scale;           // loss scale folded into every delta; d_x0 is written unscaled
  const uint8_t* act[4]; // x_1 .. x_4 images of the training forward
  uint8_t* g[4];         // out: g_0 .. g_3 images (scaled)
  uint8_t* dout_img;     // out: [tiles][2][128][8] image of the scaled head gradient (columns head_n.. are zero)
  float* d_x0;           // out: fp32 [n_rays * S][k0], ray-major rows
  int64_t n_tiles, n_rays;
  int S, tiles_per_blk;
  const int32_t* row_count;   // row-list mode (device count of listed rows): g / dout_img / d_x0 come out in LIST order
  const int32_t* row_ids;     // with row_count: act / d_raw / x0 are the forward's BLOCKED buffers, read through the list,
  uint8_t* act_c[4];          //   and their listed rows are written out in LIST order for the weight gradients
  const uint8_t* x0;          //   (act_c: x_1..x_4, x0_c: the k0-channel feature image).  NULL: act / d_raw are
  uint8_t* x0_c;              //   LIST-ordered already (nvsr_compact_rows)
  int acts_listed;            // with row_ids: only d_raw is read through the list, act is LIST-ordered already (the sparse
                              //   training forward wrote it so); no copies are written
};

// this thread's mask words: the 64 activations of image row `src` (< 0: a zero row), columns [col0, col0 + 64), eight per
// 16-byte group; `copy` (row-list mode): the same words go to row r of tile `tile` of the LIST-ordered image
__device__ __forceinline__ void load_mask64(const uint8_t* img, int64_t src, int col0, uint4 (&m)[8], uint8_t* copy,
                                            int64_t tile, int r) {
  if (src >= 0) {
    const uint4* p = reinterpret_cast<const uint4*>(img + (src >> 7) * (int64_t)kActTileBytes) + (col0 >> 3) * 128 + (src & 127);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = __ldg(p + j * 128);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (copy) {
    uint4* q = reinterpret_cast<uint4*>(copy + tile * (int64_t)kActTileBytes) + (col0 >> 3) * 128 + r;
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j * 128] = m[j];
  }
}
// activations are post-ReLU (>= +0): "positive" == any non-sign bit set
__device__ __forceinline__ bool act_pos(const uint4& m, int e) {
  const uint32_t w = e < 2 ? m.x : (e < 4 ? m.y : (e < 6 ? m.z : m.w));
  return ((e & 1) ? (w >> 16) : (w & 0xffffu)) & 0x7fffu;
}
// pack 64 fp32 deltas (already masked) to fp16 pairs, hand them to the next MMA (TMEM A region) and to the image
__device__ __forceinline__ void emit_delta64(const float (&d)[64], uint32_t a_addr, uint8_t* g_tile, int col0, int r) {
  uint32_t pk[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) pk[j] = pack16x2<true>(d[2 * j], d[2 * j + 1]);
  tmem_st32(a_addr, pk);
  uint4* p = reinterpret_cast<uint4*>(g_tile) + (col0 >> 3) * 128 + r;
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j * 128] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

// Two tiles ("slots") in flight per CTA: 8 epilogue warps (TMEM lane quadrant x column half) + warp 8 as the MMA issuer.
// While the epilogue warps mask / pack / store one slot's deltas, the tensor core runs the other slot's K loop.
__global__ void __launch_bounds__(kDgThreads, 1) dgrad_chain_kernel(const __grid_constant__ DgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_mma[2], bar_ready[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_headw[4 * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // smem: W_3 | W_2 | W_1 (32 KB each) | W_0 (k0 * 256 B)
  const uint32_t w_hidden = 128u * 128u * 2u, w0_bytes = (uint32_t)a.k0 * 256u;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&bar_mma[s], 1), mbar_init(&bar_ready[s], 8);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kDgTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4 * 128; i += kDgThreads) s_headw[i] = (i >> 7) < a.head_n ? __ldg(a.head_w + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  int64_t n_tiles = a.n_tiles, listed_rows = 0;
  if (a.row_count) {
    listed_rows = (int64_t)__ldg(a.row_count);
    const int64_t listed = (listed_rows + kTileRows - 1) / kTileRows;
    n_tiles = listed < n_tiles ? listed : n_tiles;
  }
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 8) {
    // ================= issuer =================
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(&bar_w, 3u * w_hidden + w0_bytes);
      for (int l = 3; l >= 1; --l) bulk_g2s(smem + (3 - l) * w_hidden, a.w[l], w_hidden, &bar_w);
      bulk_g2s(smem + 3 * w_hidden, a.w[0], w0_bytes, &bar_w);
    }
    if (my_tiles > 0) mbar_wait(&bar_w, 0);
    uint32_t ph[2] = {0u, 0u};
    for (int64_t p = 0; 2 * p < my_tiles; ++p) {
#pragma unroll 1
      for (int l = 3; l >= 0; --l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (2 * p + s >= my_tiles) continue;
          mbar_wait(&bar_ready[s], ph[s]);
          ph[s] ^= 1u;
          tc_fence_after();
          if (elect_one()) {
            // B = forward weight image [k_in/8][128 n_out][8] read MN-major: N = k_in, K = n_out
            const uint32_t d = tmem + (uint32_t)s * kDgSlotCols;
            const uint64_t b0 = smem_desc(smem_u32(smem + (3 - l) * w_hidden), 128u, 2048u);
            const uint32_t idesc = idesc_f16(l > 0 ? 128 : a.k0, false, true);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ts(d, d + kDgAOff + (uint32_t)ks * 8u, b0 + (uint64_t)(ks * 16), idesc, ks ? 1u : 0u);
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
    auto arrive_ready = [&](int s) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ready[s]);
    };
    // g_3 = (d_out . W_head) * [x_4 > 0] of `tile` on the CUDA cores -> slot s's A region + the g[3] / d_out images
    // the buffer row that feeds row r of `tile`: itself, or (row-list mode) the listed BLOCKED row; -1 = the list's zero tail
    auto src_row = [&](int64_t tile) -> int64_t {
      const int64_t i = tile * kTileRows + r;
      if (!a.row_ids) return i;
      return i < listed_rows ? (int64_t)__ldg(a.row_ids + i) : -1;
    };
    const bool copy_acts = a.row_ids && !a.acts_listed;
    // the row of the activation images that belongs to list row (tile, r) whose buffer row is `src`
    auto act_row = [&](int64_t tile, int64_t src) -> int64_t {
      return (a.acts_listed && src >= 0) ? tile * kTileRows + r : src;
    };
    auto prep_tile = [&](int s, int64_t tile) {
      const uint32_t a_tmem = lane_base + (uint32_t)s * kDgSlotCols + kDgAOff + (uint32_t)(col0 >> 1);
      const int64_t src = src_row(tile);
      float dv[4];
#pragma unroll
      for (int h = 0; h < 4; ++h)
        dv[h] = (h < a.head_n && src >= 0) ? __ldg(a.d_raw + (int64_t)(a.head_ch + h) * a.raw_stride + src) * a.scale : 0.f;
      uint4 m[8];
      load_mask64(a.act[3], act_row(tile, src), col0, m, copy_acts ? a.act_c[3] : nullptr, tile, r);
      if (copy_acts && a.x0_c) {   // the feature image's listed rows, for the layer-0 weight gradient
        const int chunks = a.k0 >> 3, c_half = (chunks + 1) >> 1;
        const int c0 = half == 0 ? 0 : c_half, c1 = half == 0 ? c_half : chunks;
        const uint4* px = reinterpret_cast<const uint4*>(a.x0) + (src >= 0 ? (src >> 7) * chunks * kTileRows + (src & 127) : 0);
        uint4* qx = reinterpret_cast<uint4*>(a.x0_c) + tile * chunks * kTileRows + r;
        for (int c = c0; c < c1; ++c) qx[c * kTileRows] = src >= 0 ? __ldg(px + c * kTileRows) : make_uint4(0u, 0u, 0u, 0u);
      }
      float d[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = dv[0] * s_headw[col0 + c];
#pragma unroll
        for (int h = 1; h < 4; ++h) v = fmaf(dv[h], s_headw[h * 128 + col0 + c], v);
        d[c] = act_pos(m[c >> 3], c & 7) ? v : 0.f;
      }
      emit_delta64(d, a_tmem, a.g[3] + tile * (int64_t)kActTileBytes, col0, r);
      if (half == 0) {   // the head gradient as a 16-column image (the head weight gradient's B operand)
        uint4* pimg = reinterpret_cast<uint4*>(a.dout_img + tile * (int64_t)(2 * kTileRows * 16)) + r;
        pimg[0] = make_uint4(pack16x2<true>(dv[0], dv[1]), pack16x2<true>(dv[2], dv[3]), 0u, 0u);
        pimg[128] = make_uint4(0u, 0u, 0u, 0u);
      }
    };

    for (int s = 0; s < 2; ++s)
      if (s < my_tiles) {
        prep_tile(s, blockIdx.x + (int64_t)s * gridDim.x);
        arrive_ready(s);
      }
    for (int64_t p = 0; 2 * p < my_tiles; ++p) {
#pragma unroll 1
      for (int l = 3; l >= 0; --l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int64_t j = 2 * p + s;
          if (j >= my_tiles) continue;
          const int64_t tile = blockIdx.x + j * gridDim.x;
          const uint32_t d_tmem = lane_base + (uint32_t)s * kDgSlotCols;
          const uint32_t a_tmem = d_tmem + kDgAOff + (uint32_t)(col0 >> 1);
          // the mask of the layer below does not depend on the accumulator: fetch it while the MMAs run
          uint4 m[8];
          if (l > 0) load_mask64(a.act[l - 1], act_row(tile, src_row(tile)), col0, m, copy_acts ? a.act_c[l - 1] : nullptr, tile, r);
          mbar_wait(&bar_mma[s], ph[s]);
          ph[s] ^= 1u;
          tc_fence_after();
          if (l > 0) {
            float d[64];
            {
              uint32_t v0[32], v1[32];
              tmem_ld32(d_tmem + (uint32_t)col0, v0);
              tmem_ld32(d_tmem + (uint32_t)col0 + 32u, v1);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                d[c] = act_pos(m[c >> 3], c & 7) ? __uint_as_float(v0[c]) : 0.f;
                d[32 + c] = act_pos(m[4 + (c >> 3)], c & 7) ? __uint_as_float(v1[c]) : 0.f;
              }
            }
            emit_delta64(d, a_tmem, a.g[l - 1] + tile * (int64_t)kActTileBytes, col0, r);
            arrive_ready(s);
          } else {
            // d_x0: 16-column units split over the two column halves; fp32, unscaled, ray-major rows
            const int units = a.k0 >> 4, u_half = (units + 1) >> 1;
            const int u0 = half == 0 ? 0 : u_half, u1 = half == 0 ? u_half : units;
            const int64_t blk = tile / a.tiles_per_blk;
            const int smp = (int)(tile - blk * a.tiles_per_blk) * kBlkSamples + (r >> 3);
            const int64_t ray = blk * kBlkRays + (r & 7);
            // row-list mode: row i of the list -> d_x0 row i (the list's zero-padded tail rows come out as zeros)
            const bool valid = a.row_count ? true : (ray < a.n_rays && smp < a.S);
            const float inv = 1.0f / a.scale;
            float* out = a.d_x0 + (a.row_count ? (tile * kTileRows + r) * (int64_t)a.k0 : (valid ? (ray * a.S + smp) * (int64_t)a.k0 : 0));
            for (int u = u0; u < u1; ++u) {
              uint32_t v[16];
              tmem_ld16(d_tmem + (uint32_t)(u * 16), v);
              tmem_ld_wait();
              if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(out + u * 16 + 4 * q) =
                      make_float4(__uint_as_float(v[4 * q]) * inv, __uint_as_float(v[4 * q + 1]) * inv,
                                  __uint_as_float(v[4 * q + 2]) * inv, __uint_as_float(v[4 * q + 3]) * inv);
              }
            }
            // the slot's accumulator and A region are free: its next tile's g_3 goes in
            if (j + 2 < my_tiles) {
              prep_tile(s, blockIdx.x + (j + 2) * gridDim.x);
              arrive_ready(s);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kDgTmemCols) : "memory");
  }
}

// ---- weight gradient ------------------------------------------------------------------------------------------
constexpr int kWgThreads = 128;
constexpr int kWgStages = 3;

// dW[m][n] += inv_scale * sum over the tiles' rows of A[r][m] * B[r][n]  (A: 128 channels, B: n_b channels, both tile
// images), db[m] += inv_scale * sum over rows of A[r][m] (optional).  One warp loads and issues, four read out.
// up to five products per launch (blockIdx.y): the weight gradients of one chain are independent, and with a short row
// list each needs only a few CTAs, so they run side by side instead of one launch after the other
struct WgradProblem {
  const uint8_t* a_img;
  const uint8_t* b_img;
  int n_b;
  float* dw;
  int64_t ldw;
  float* db;
};
struct WgradArgs {
  WgradProblem p[5];
  int64_t n_tiles;
  const int32_t* row_count;
  int min_tiles;
  float inv_scale;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradArgs args) {
  const WgradProblem& pr = args.p[blockIdx.y];
  const uint8_t* __restrict__ a_img = pr.a_img;
  const uint8_t* __restrict__ b_img = pr.b_img;
  const int n_b = pr.n_b;
  float* __restrict__ dw = pr.dw;
  const int64_t ldw = pr.ldw;
  float* __restrict__ db = pr.db;
  int64_t n_tiles = args.n_tiles;
  const int32_t* __restrict__ row_count = args.row_count;
  const int min_tiles = args.min_tiles;
  const float inv_scale = args.inv_scale;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kWgStages], empty[kWgStages], done;
  __shared__ uint32_t tmem_slot;
  const uint32_t a_bytes = kActTileBytes, b_bytes = (uint32_t)kTileRows * (uint32_t)n_b * 2u, st_bytes = a_bytes + b_bytes;
  uint8_t* ones = smem + kWgStages * st_bytes;   // [2][128][8] fp16 1.0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * kTileRows * 8 / 2; i += kWgThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (row_count) {   // row-list mode: only the tiles the list fills (its tail rows are zero in both images)
    const int64_t listed = ((int64_t)__ldg(row_count) + kTileRows - 1) / kTileRows;
    n_tiles = listed < n_tiles ? listed : n_tiles;
  }
  // every participating CTA ends with 128 x n_b atomics into the same accumulator: with few tiles (a short row list)
  // fewer CTAs take part, at least `min_tiles` tiles each
  int64_t ctas = n_tiles / min_tiles;
  ctas = ctas < 1 ? 1 : (ctas > (int64_t)gridDim.x ? (int64_t)gridDim.x : ctas);
  const int64_t my_tiles = blockIdx.x < ctas && blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + ctas - 1) / ctas : 0;

  if (threadIdx.x == 0 && my_tiles > 0) {
    auto load = [&](int64_t t) {
      const int s = (int)(t % kWgStages);
      const int64_t tile = blockIdx.x + t * ctas;
      mbar_arrive_expect_tx(&full[s], st_bytes);
      bulk_g2s(smem + s * st_bytes, a_img + tile * a_bytes, a_bytes, &full[s]);
      bulk_g2s(smem + s * st_bytes + a_bytes, b_img + tile * (int64_t)b_bytes, b_bytes, &full[s]);
    };
    for (int64_t t = 0; t < kWgStages && t < my_tiles; ++t) load(t);
    // both operands MN-major: 8 consecutive MN (channels) in 16 B, the next 8 rows (K) 128 B further (LBO), the next 8
    // channels 128 rows * 16 B further (SBO); 16 rows per MMA = + 256 B
    const uint32_t idesc = idesc_f16(n_b, true, true), idesc1 = idesc_f16(16, true, true);
    const uint64_t o0 = smem_desc(smem_u32(ones), 128u, 2048u);
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int s = (int)(t % kWgStages);
      const uint32_t par = (uint32_t)((t / kWgStages) & 1);
      mbar_wait(&full[s], par);
      tc_fence_after();
      const uint64_t a0 = smem_desc(smem_u32(smem + s * st_bytes), 128u, 2048u);
      const uint64_t b0 = smem_desc(smem_u32(smem + s * st_bytes + a_bytes), 128u, 2048u);
#pragma unroll
      for (int ks = 0; ks < kTileRows / 16; ++ks) {
        umma_ss(tmem, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, (t | ks) ? 1u : 0u);
        if (db) umma_ss(tmem + 256u, a0 + (uint64_t)(ks * 16), o0 + (uint64_t)(ks * 16), idesc1, (t | ks) ? 1u : 0u);
      }
      umma_commit(&empty[s]);   // arrives when the MMAs above have finished reading stage s
      // refill the stage of the PREVIOUS tile (its MMAs have had this tile's issue time to finish): the issuer never
      // waits for the MMAs it has just issued
      if (t >= 1 && t - 1 + kWgStages < my_tiles) {
        mbar_wait(&empty[(t - 1) % kWgStages], (uint32_t)(((t - 1) / kWgStages) & 1));
        load(t - 1 + kWgStages);
      }
    }
    umma_commit(&done);
  }
  __syncwarp();
  if (my_tiles > 0) {
    mbar_wait(&done, 0);
    tc_fence_after();
    // warp w reads TMEM lanes 32w .. 32w+31 = output rows m; 16 columns (n) at a time
    const int m = warp * 32 + lane;
    // 128-bit vector reductions (RED.E.ADD.F32x4) when the accumulator rows are 16-byte aligned: the partial sums of all
    // participating CTAs meet in 128 x n_b addresses, and the L2's reduction rate — not the streaming — bounds a launch
    // over a short row list (4x fewer operations: 0.30 -> see DESIGN 4.6)
    const bool vec = ((reinterpret_cast<uintptr_t>(dw) | (uintptr_t)(ldw * 4)) & 15u) == 0;
    for (int c0 = 0; c0 < n_b; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      float* row = dw + (int64_t)m * ldw + c0;
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          atomicAdd(reinterpret_cast<float4*>(row) + q,
                    make_float4(__uint_as_float(v[4 * q]) * inv_scale, __uint_as_float(v[4 * q + 1]) * inv_scale,
                                __uint_as_float(v[4 * q + 2]) * inv_scale, __uint_as_float(v[4 * q + 3]) * inv_scale));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(row + j, __uint_as_float(v[j]) * inv_scale);
      }
    }
    if (db) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 256u, v);
      tmem_ld_wait();
      atomicAdd(db + m, __uint_as_float(v[0]) * inv_scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// out[ray][c] = inv_scale * sum over the ray's samples of img[row(ray, s)][c]; one thread per (ray, 8-channel chunk)
__global__ void __launch_bounds__(256)
ray_sum_kernel(const uint8_t* __restrict__ img, int64_t n_rays, int S, int tiles_per_blk, float inv_scale, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * 16) return;
  const int64_t ray = idx >> 4;
  const int j = (int)(idx & 15);
  const int64_t blk = ray / kBlkRays;
  const int rr = (int)(ray - blk * kBlkRays);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < S; ++s) {
    const int64_t tile = blk * tiles_per_blk + s / kBlkSamples;
    const int row = (s % kBlkSamples) * kBlkRays + rr;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + tile * (int64_t)kActTileBytes) + j * 128 + row);
    const float2 f0 = unpack16x2<true>(v.x), f1 = unpack16x2<true>(v.y), f2 = unpack16x2<true>(v.z), f3 = unpack16x2<true>(v.w);
    acc[0] += f0.x, acc[1] += f0.y, acc[2] += f1.x, acc[3] += f1.y, acc[4] += f2.x, acc[5] += f2.y, acc[6] += f3.x, acc[7] += f3.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) out[ray * 128 + j * 8 + e] = acc[e] * inv_scale;
}


// ---- row-list ("sparse") backward ---------------------------------------------------------------------------------
// A sample whose raw gradient is identically zero (alpha = 0: sigma + noise <= 0, or transmittance 0) contributes
// nothing to any weight / plane gradient: the backward chains need only the other rows.  nonzero_rows_kernel lists them
// (BLOCKED row ids, order unspecified), compact_rows_kernel copies those rows of the forward's images and of d_raw into
// dense tiles in list order (tail of the last tile zero-filled, so it adds nothing to the weight gradients).
__global__ void __launch_bounds__(256)
nonzero_rows_kernel(const float* __restrict__ d_raw, int64_t stride, int64_t n_rows, int32_t* __restrict__ ids,
                    int32_t* __restrict__ count) {
  __shared__ int warp_cnt[8];
  __shared__ int block_base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 1024; base < n_rows; base += (int64_t)gridDim.x * 1024) {
    unsigned m[4];
    bool k[4];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t row = base + j * 256 + threadIdx.x;
      k[j] = false;
      if (row < n_rows) {
#pragma unroll
        for (int h = 0; h < 4; ++h) k[j] = k[j] || !(__ldg(d_raw + h * stride + row) == 0.f);   // NaN is kept
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
        const int c = warp_cnt[w];
        warp_cnt[w] = tot;
        tot += c;
      }
      block_base = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    int off = block_base + warp_cnt[wid];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k[j]) ids[off + __popc(m[j] & ((1u << lane) - 1u))] = (int32_t)(base + j * 256 + threadIdx.x);
      off += __popc(m[j]);
    }
    __syncthreads();
  }
}

struct CompactArgs {
  const uint8_t* src[12];
  uint8_t* dst[12];
  int chunks[12];     // 16-byte chunks per row (channels / 8)
  int n_img;
  const float* d_raw;      // planar [4][raw_stride]
  float* d_raw_out;        // planar [4][out_stride]
  int64_t raw_stride, out_stride;
  const int32_t* ids;
  const int32_t* count;
};

// grid (tiles, n_img + 1): block (t, k) fills tile t of image k (k == n_img: the four d_raw planes)
__global__ void __launch_bounds__(kTileRows) compact_rows_kernel(const __grid_constant__ CompactArgs a) {
  const int n = __ldg(a.count);
  const int r = threadIdx.x, k = blockIdx.y;
  for (int64_t tile = blockIdx.x; tile * kTileRows < n; tile += gridDim.x) {   // the listed tiles only (device-side count)
    const int64_t i = tile * kTileRows + r;
    const bool live = i < n;
    const int64_t rid = live ? __ldg(a.ids + i) : 0;
    if (k == a.n_img) {
#pragma unroll
      for (int h = 0; h < 4; ++h) a.d_raw_out[h * a.out_stride + i] = live ? __ldg(a.d_raw + h * a.raw_stride + rid) : 0.f;
      continue;
    }
    const int chunks = a.chunks[k];
    const uint4* src = reinterpret_cast<const uint4*>(a.src[k]) + (rid >> 7) * chunks * kTileRows + (rid & 127);
    uint4* dst = reinterpret_cast<uint4*>(a.dst[k]) + tile * chunks * kTileRows + r;
    int c = 0;
    for (; c + 4 <= chunks; c += 4) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = live ? __ldg(src + (c + q) * kTileRows) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[(c + q) * kTileRows] = v[q];
    }
    for (; c < chunks; ++c) dst[c * kTileRows] = live ? __ldg(src + c * kTileRows) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// per-ray sums of a LIST-ordered 128-channel image: out[ray(ids[i])][c] += inv_scale * img[i][c]  (out zeroed by the caller)
__global__ void __launch_bounds__(256)
ray_sum_rows_kernel(const uint8_t* __restrict__ img, const int32_t* __restrict__ ids, const int32_t* __restrict__ count,
                    int tiles_per_blk, float inv_scale, float* __restrict__ out) {
  // grid-stride over the listed rows only: the list length is known on the device, a grid sized for the capacity would
  // be mostly blocks that exit at once (65 536 of them for 1 M rows: ~20 us of pure block scheduling)
  const int64_t total = (int64_t)__ldg(count) * 16;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx >> 4;
    const int j = (int)(idx & 15);
    const int32_t rid = __ldg(ids + i);
    const int64_t ray = (int64_t)((rid >> 7) / tiles_per_blk) * kBlkRays + (rid & (kBlkRays - 1));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (i >> 7) * (int64_t)kActTileBytes) + j * 128 + (i & 127));
    const float2 f0 = unpack16x2<true>(v.x), f1 = unpack16x2<true>(v.y), f2 = unpack16x2<true>(v.z), f3 = unpack16x2<true>(v.w);
    float4* o = reinterpret_cast<float4*>(out + ray * 128 + j * 8);
    atomicAdd(o, make_float4(f0.x * inv_scale, f0.y * inv_scale, f1.x * inv_scale, f1.y * inv_scale));
    atomicAdd(o + 1, make_float4(f2.x * inv_scale, f2.y * inv_scale, f3.x * inv_scale, f3.y * inv_scale));
  }
}

}  // namespace

int32_t launch_mlp_tc(const nvsr_mlp_t* m, cudaStream_t st, void* const* act_out);

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_mlp_chain_train(const nvsr_mlp_t* m, void* const* act_out, void* stream) {
  NVSR_CHECK_ARG(m && act_out && m->n_layers == 4 && m->precision == NVSR_F16);
  NVSR_CHECK_ARG(m->in && m->raw && m->rows >= 0 && m->raw_stride >= m->rows);
  if (m->rows == 0) return NVSR_OK;
  return launch_mlp_tc(m, (cudaStream_t)stream, act_out);
}

extern "C" int32_t nvsr_mlp_dgrad(const nvsr_dgrad_t* d, void* stream) {
  NVSR_CHECK_ARG(d && d->k0 > 0 && (d->k0 % 16) == 0 && d->k0 <= 192 && d->head_n >= 1 && d->head_n <= 4);
  NVSR_CHECK_ARG(d->head_ch >= 0 && d->head_ch + d->head_n <= 4 && d->head_w && d->d_raw && d->dout_img && d->d_x0);
  NVSR_CHECK_ARG(d->n_rays > 0 && d->n_samples > 0 && d->scale > 0.f);
  DgradArgs a;
  for (int l = 0; l < 4; ++l) {
    NVSR_CHECK_ARG(d->w[l] && d->act[l] && d->g[l]);
    if (!aligned16(d->w[l]) || !aligned16(d->act[l]) || !aligned16(d->g[l])) return NVSR_ERR_ALIGNMENT;
    a.w[l] = (const uint8_t*)d->w[l], a.act[l] = (const uint8_t*)d->act[l], a.g[l] = (uint8_t*)d->g[l];
  }
  if (!aligned16(d->dout_img) || !aligned16(d->d_x0)) return NVSR_ERR_ALIGNMENT;
  a.k0 = d->k0, a.head_w = d->head_w, a.head_n = d->head_n, a.head_ch = d->head_ch;
  a.d_raw = d->d_raw, a.raw_stride = d->raw_stride, a.scale = d->scale;
  a.dout_img = (uint8_t*)d->dout_img, a.d_x0 = d->d_x0;
  a.n_rays = d->n_rays, a.S = d->n_samples, a.tiles_per_blk = tiles_per_block(d->n_samples);
  a.n_tiles = ceil_div64(d->n_rays, kBlkRays) * a.tiles_per_blk;
  a.row_count = d->row_count, a.row_ids = d->row_count ? d->row_ids : nullptr;
  a.x0 = nullptr, a.x0_c = nullptr;
  a.acts_listed = (a.row_ids && d->acts_listed) ? 1 : 0;
  for (int l = 0; l < 4; ++l) a.act_c[l] = nullptr;
  if (a.row_ids && !a.acts_listed) {
    for (int l = 0; l < 4; ++l) {
      NVSR_CHECK_ARG(d->act_list[l]);
      if (!aligned16(d->act_list[l])) return NVSR_ERR_ALIGNMENT;
      a.act_c[l] = (uint8_t*)d->act_list[l];
    }
    NVSR_CHECK_ARG((d->x0_img == nullptr) == (d->x0_list == nullptr));
    if (!aligned16(d->x0_img) || !aligned16(d->x0_list)) return NVSR_ERR_ALIGNMENT;
    a.x0 = (const uint8_t*)d->x0_img, a.x0_c = (uint8_t*)d->x0_list;
  }
  NVSR_CHECK_ARG(d->raw_stride >= a.n_tiles * kTileRows);
  const uint32_t smem_bytes = 3u * 128u * 128u * 2u + (uint32_t)d->k0 * 256u;
  cudaError_t e = cudaFuncSetAttribute(dgrad_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int32_t)e;
  const int64_t grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  dgrad_chain_kernel<<<(unsigned)grid, kDgThreads, smem_bytes, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

constexpr int kWgMinTiles = 16;   // tiles per participating CTA before another CTA joins (A/B on B200, scalar reductions: 1: 0.93, 8: 0.70, 16: 0.66 ms; with vector reductions 8 and 16 measure the same)

// n products (<= 5) of one chain in one launch
static int32_t wgrad_launch(const WgradProblem* pr, int n, int64_t n_tiles, const int32_t* row_count, float inv_scale,
                            void* stream) {
  NVSR_CHECK_ARG(n >= 1 && n <= 5 && n_tiles >= 0);
  WgradArgs a;
  uint32_t smem_bytes = 0;
  for (int i = 0; i < n; ++i) {
    const WgradProblem& q = pr[i];
    NVSR_CHECK_ARG(q.a_img && q.b_img && q.dw && q.n_b >= 16 && (q.n_b % 16) == 0 && q.n_b <= 256 && q.ldw >= q.n_b);
    if (!aligned16(q.a_img) || !aligned16(q.b_img)) return NVSR_ERR_ALIGNMENT;
    const uint32_t need = kWgStages * (kActTileBytes + (uint32_t)kTileRows * (uint32_t)q.n_b * 2u) + 2u * kTileRows * 16u;
    smem_bytes = need > smem_bytes ? need : smem_bytes;
    a.p[i] = q;
  }
  if (n_tiles == 0) return NVSR_OK;
  if (smem_bytes > 227u * 1024u) return NVSR_ERR_RESOURCE;
  a.n_tiles = n_tiles, a.row_count = row_count, a.min_tiles = kWgMinTiles, a.inv_scale = inv_scale;
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int32_t)e;
  const int64_t grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  wgrad_kernel<<<dim3((unsigned)grid, (unsigned)n), kWgThreads, smem_bytes, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_mlp_wgrad(const void* a_img, const void* b_img, int32_t n_b, int64_t n_tiles, float inv_scale,
                                  float* dw, int64_t ldw, float* db, void* stream) {
  const WgradProblem q{(const uint8_t*)a_img, (const uint8_t*)b_img, n_b, dw, ldw, db};
  return wgrad_launch(&q, 1, n_tiles, nullptr, inv_scale, stream);
}

extern "C" int32_t nvsr_mlp_wgrad_chain_rows(const void* const* g, const void* x0_img, int32_t k0, const void* const* act,
                                             const void* dout_img, int64_t n_tiles, const int32_t* row_count,
                                             float inv_scale, float* const* dw, const int64_t* ldw, float* const* db,
                                             float* dw_head, void* stream) {
  NVSR_CHECK_ARG(g && x0_img && act && dout_img && dw && ldw && db && dw_head);
  WgradProblem q[5];
  for (int l = 0; l < 4; ++l)
    q[l] = WgradProblem{(const uint8_t*)g[l], (const uint8_t*)(l == 0 ? x0_img : act[l - 1]), l == 0 ? k0 : 128, dw[l], ldw[l], db[l]};
  q[4] = WgradProblem{(const uint8_t*)act[3], (const uint8_t*)dout_img, 16, dw_head, 16, nullptr};
  return wgrad_launch(q, 5, n_tiles, row_count, inv_scale, stream);
}

extern "C" int32_t nvsr_mlp_wgrad_chain(const void* const* g, const void* x0_img, int32_t k0, const void* const* act,
                                        const void* dout_img, int64_t n_tiles, float inv_scale, float* const* dw,
                                        const int64_t* ldw, float* const* db, float* dw_head, void* stream) {
  return nvsr_mlp_wgrad_chain_rows(g, x0_img, k0, act, dout_img, n_tiles, nullptr, inv_scale, dw, ldw, db, dw_head, stream);
}

extern "C" int32_t nvsr_nonzero_rows(const float* d_raw, int64_t raw_stride, int64_t n_rows, int32_t* row_ids, int32_t* count,
                                     void* stream) {
  NVSR_CHECK_ARG(d_raw && row_ids && count && n_rows >= 0 && raw_stride >= n_rows && n_rows < ((int64_t)1 << 31));
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int32_t)e;
  if (n_rows == 0) return NVSR_OK;
  int64_t blocks = ceil_div64(n_rows, 1024);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  nonzero_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_raw, raw_stride, n_rows, row_ids, count);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_compact_rows(const void* const* src, void* const* dst, const int32_t* channels, int32_t n_img,
                                     const float* d_raw, int64_t raw_stride, float* d_raw_out, int64_t out_stride,
                                     const int32_t* row_ids, const int32_t* count, int64_t max_tiles, void* stream) {
  NVSR_CHECK_ARG(src && dst && channels && n_img >= 0 && n_img <= 12 && row_ids && count && ((d_raw == nullptr) == (d_raw_out == nullptr)));
  NVSR_CHECK_ARG(max_tiles >= 0 && max_tiles <= 0x7fffffff && (!d_raw || out_stride >= max_tiles * kTileRows));
  if (max_tiles == 0) return NVSR_OK;
  CompactArgs a;
  for (int k = 0; k < n_img; ++k) {
    NVSR_CHECK_ARG(src[k] && dst[k] && channels[k] > 0 && channels[k] % 8 == 0);
    if (!aligned16(src[k]) || !aligned16(dst[k])) return NVSR_ERR_ALIGNMENT;
    a.src[k] = (const uint8_t*)src[k], a.dst[k] = (uint8_t*)dst[k], a.chunks[k] = channels[k] / 8;
  }
  a.n_img = n_img, a.d_raw = d_raw, a.d_raw_out = d_raw_out, a.raw_stride = raw_stride, a.out_stride = out_stride;
  a.ids = row_ids, a.count = count;
  const int64_t gx = max_tiles < (int64_t)kNumSMs * 8 ? max_tiles : (int64_t)kNumSMs * 8;
  compact_rows_kernel<<<dim3((unsigned)gx, (unsigned)(n_img + (d_raw ? 1 : 0))), kTileRows, 0, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_ray_sum_rows(const void* img, const int32_t* row_ids, const int32_t* count, int64_t max_rows,
                                     int32_t n_samples, float inv_scale, float* out, void* stream) {
  NVSR_CHECK_ARG(img && row_ids && count && out && max_rows >= 0 && n_samples > 0);
  if (!aligned16(img) || !aligned16(out)) return NVSR_ERR_ALIGNMENT;
  if (max_rows == 0) return NVSR_OK;
  int64_t blocks = ceil_div64(max_rows * 16, 256);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  ray_sum_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)img, row_ids, count, tiles_per_block(n_samples), inv_scale, out);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_ray_sum(const void* img, int64_t n_rays, int32_t n_samples, float inv_scale, float* out, void* stream) {
  NVSR_CHECK_ARG(img && out && n_rays >= 0 && n_samples > 0);
  if (n_rays == 0) return NVSR_OK;
  const int64_t threads = n_rays * 16;
  ray_sum_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)img, n_rays, n_samples,
                                                                                       tiles_per_block(n_samples), inv_scale, out);
  NVSR_RETURN_LAST_ERROR();
}
