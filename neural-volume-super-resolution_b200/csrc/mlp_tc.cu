// a6: decoder MLP chain on the 5th-gen tensor cores (tcgen05.mma, accumulators AND hidden
// activations in TMEM).
//
// Persistent kernel, one CTA per SM, cta_group::1, UMMA M=128 (one 128-row tile), N = layer width.
//   warp 0 (1 lane)  : producer — bulk-copies (TMA engine) the chain's 16-bit weights into shared
//                      memory once, then one feature tile image per tile into a 2-deep ring
//   warp 1 (1 lane)  : MMA issuer — layer 0: A = feature tile in smem (SS form); hidden layers: A =
//                      previous layer's activations in TMEM (TS form); B = resident weights in smem;
//                      D = fp32 accumulator in TMEM; completion committed to an mbarrier
//   warps 2..17      : epilogue — tcgen05.ld the accumulator (warp w owns TMEM lanes 32*(w%4)..+31,
//                      i.e. 32 rows, and one quarter of the layer's columns), + bias (smem, or per-ray
//                      from global), ReLU + saturate + 16-bit pack in ONE F2FP per pair, tcgen05.st
//                      the packed row back into TMEM as the next layer's A operand.  Output heads
//                      (128 -> 1 / 3) are fp32 dot products on the CUDA cores over the UNROUNDED
//                      last activations.
// Two tiles ("slots") are in flight and ping-pong: while the epilogue warps work on slot A's layer l,
// the tensor core runs slot B's layer l.  Hidden activations never leave the SM (never even touch
// shared memory); weights are read from HBM/L2 once per CTA; the feature ring slot is released as
// soon as layer 0's MMAs have been committed.
//
// TMEM map (512 columns allocated): slot s -> D_s = [192 s, 192 s + 128) fp32 accumulator,
// A_s = [192 s + 128, 192 s + 192): 128 rows x 128 16-bit activations, two per 32-bit column
// (row = lane, K pair k/2 = column: the TS-form A layout).
//
// smem operand layout (layer-0 A and every B): K-major, no-swizzle canonical UMMA layout with
// 8x16-byte core matrices, stored [K/8][rows][8 x 16 bit]: LBO (K-chunk stride) = rows*16 B,
// SBO (8-row group stride) = 128 B.  The feature tile image written by the gather kernel and the
// weight image written by nvsr_pack_weight16 are exactly this, so both arrive with plain bulk copies.
#include "common.cuh"

namespace nvsr {

constexpr int kTcEpiWarps = 16;
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr int kTcMaxHeads = 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSlotCols = 192;  // D (128 fp32 columns) + A (64 columns of packed 16-bit pairs)
constexpr uint32_t kSlotAOff = 128;

struct TcLayer {
  const void* w;          // global 16-bit image
  const float* bias;      // global
  const float* row_bias;  // global per-ray or null
  const float* head_w;
  const float* head_b;
  int k, n, relu, head_n, head_ch, head_slot;
  uint32_t w_off;         // smem byte offset of the weight image
};

struct TcArgs {
  TcLayer layer[NVSR_MAX_LAYERS];
  int n_layers;
  const uint8_t* in;      // tile images, in_bytes each
  uint32_t in_bytes;
  int64_t rows, n_tiles;
  int samples_per_ray, row_order, tiles_per_blk;
  int64_t n_rays;
  float* raw;
  int64_t raw_stride;
  // smem carve-up (byte offsets from the 1024-aligned base)
  uint32_t in_off[2], bias_off, headw_off, hpart_off, bar_off, w_bytes_total;
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// kind::f16 instruction descriptor: (bf16|f16) x (bf16|f16) -> fp32, both operands K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_16(int n, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;  // F16F32Format: 0 = F16, 1 = BF16
  return (1u << 4) /*D=f32*/ | (fmt << 7) /*A*/ | (fmt << 10) /*B*/ | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (base_lane + t)
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
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (ReLU +) saturate-to-finite + round + pack two fp32 into one 16-bit pair: a single F2FP
template <bool F16, bool RELU>
__device__ __forceinline__ uint32_t pack_act(float lo, float hi) {
  uint32_t r;
  if constexpr (F16 && RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else if constexpr (RELU) asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// barrier block layout (uint64_t each)
enum { BAR_W = 0, BAR_IN_FULL = 1, BAR_IN_FREE = 3, BAR_ACC_FULL = 5, BAR_ACT_READY = 7, BAR_ACC_FREE = 9, BAR_COUNT = 11 };

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_chain_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* sbias = reinterpret_cast<float*>(smem + a.bias_off);    // [n_layers][128]
  float* sheadw = reinterpret_cast<float*>(smem + a.headw_off);  // [kTcMaxHeads][4][128]
  float* shpart = reinterpret_cast<float*>(smem + a.hpart_off);  // [2 parity][3 quarters][128 rows][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.n_layers;
  const int64_t G = gridDim.x;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    mbar_init(&bars[BAR_W], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[BAR_IN_FULL + s], 1);
      mbar_init(&bars[BAR_IN_FREE + s], 1);
      mbar_init(&bars[BAR_ACC_FULL + s], 1);
      mbar_init(&bars[BAR_ACT_READY + s], kTcEpiWarps);
      mbar_init(&bars[BAR_ACC_FREE + s], kTcEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  // biases / head weights -> smem (tiny, read by every epilogue thread for every tile)
  for (int i = threadIdx.x; i < L * 128; i += kTcThreads) {
    int l = i >> 7, n = i & 127;
    const TcLayer& ly = a.layer[l];
    sbias[i] = (n < ly.n && ly.bias) ? __ldg(ly.bias + n) : 0.f;
  }
  for (int i = threadIdx.x; i < kTcMaxHeads * 4 * 128; i += kTcThreads) sheadw[i] = 0.f;
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const TcLayer& ly = a.layer[l];
    if (ly.head_w) {
      for (int i = threadIdx.x; i < ly.head_n * ly.n; i += kTcThreads) {
        int h = i / ly.n, n = i - h * ly.n;
        sheadw[(ly.head_slot * 4 + h) * 128 + n] = __ldg(ly.head_w + i);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[BAR_W], a.w_bytes_total);
      for (int l = 0; l < L; ++l) {
        const TcLayer& ly = a.layer[l];
        bulk_g2s(smem + ly.w_off, ly.w, (uint32_t)(ly.k * ly.n * 2), &bars[BAR_W]);
      }
      for (int64_t it = 0;; ++it) {
        int64_t tile = blockIdx.x + it * G;
        if (tile >= a.n_tiles) break;
        int s = (int)(it & 1);
        uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&bars[BAR_IN_FREE + s], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_IN_FULL + s], a.in_bytes);
        bulk_g2s(smem + a.in_off[s], a.in + tile * (int64_t)a.in_bytes, a.in_bytes, &bars[BAR_IN_FULL + s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      mbar_wait(&bars[BAR_W], 0);
      uint32_t ph_act[2] = {0, 0};
      for (int64_t p = 0;; ++p) {
        bool valid[2];
        valid[0] = blockIdx.x + (2 * p) * G < a.n_tiles;
        valid[1] = blockIdx.x + (2 * p + 1) * G < a.n_tiles;
        if (!valid[0]) break;
        for (int l = 0; l < L; ++l) {
          const TcLayer& ly = a.layer[l];
          const uint32_t idesc = umma_idesc_16(ly.n, F16);
          const uint32_t b_lbo = (uint32_t)ly.n * 16u;
          const uint32_t b_base = smem_u32(smem + ly.w_off);
          const int ksteps = ly.k >> 4;
          for (int s = 0; s < 2; ++s) {
            if (!valid[s]) continue;
            const uint32_t d_tmem = tmem_base + (uint32_t)s * kSlotCols;
            if (l == 0) {
              mbar_wait(&bars[BAR_IN_FULL + s], (uint32_t)(p & 1));
              mbar_wait(&bars[BAR_ACC_FREE + s], (uint32_t)(p & 1) ^ 1);
              tc_fence_after();
              const uint32_t a_base = smem_u32(smem + a.in_off[s]);
              for (int ks = 0; ks < ksteps; ++ks) {
                uint64_t ad = umma_desc(a_base + (uint32_t)ks * 2u * 2048u, 2048u, 128u);
                uint64_t bd = umma_desc(b_base + (uint32_t)ks * 2u * b_lbo, b_lbo, 128u);
                umma_ss(d_tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
              }
              umma_commit(&bars[BAR_IN_FREE + s]);  // ring slot reusable once these MMAs have read it
            } else {
              mbar_wait(&bars[BAR_ACT_READY + s], ph_act[s]);
              ph_act[s] ^= 1;
              tc_fence_after();
              const uint32_t a_tmem = d_tmem + kSlotAOff;
              for (int ks = 0; ks < ksteps; ++ks) {
                uint64_t bd = umma_desc(b_base + (uint32_t)ks * 2u * b_lbo, b_lbo, 128u);
                umma_ts(d_tmem, a_tmem + (uint32_t)ks * 8u, bd, idesc, ks > 0 ? 1u : 0u);
              }
            }
            umma_commit(&bars[BAR_ACC_FULL + s]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int qtr = ew >> 2;              // which quarter of the layer's columns
    const int r = quad * 32 + lane;       // row within the tile == TMEM lane
    const uint32_t lane_field = (uint32_t)(quad * 32) << 16;
    uint32_t ph_acc[2] = {0, 0};
    uint32_t head_parity = 0;
    for (int64_t p = 0;; ++p) {
      int64_t tile_of[2] = {blockIdx.x + (2 * p) * G, blockIdx.x + (2 * p + 1) * G};
      bool valid[2] = {tile_of[0] < a.n_tiles, tile_of[1] < a.n_tiles};
      if (!valid[0]) break;
      for (int l = 0; l < L; ++l) {
        const TcLayer& ly = a.layer[l];
        const bool last = l == L - 1;
        // 32 columns per warp: a 128-wide layer uses all four quarters, a 64-wide one the first two
        const int col = qtr * 32;
        const bool active = col < ly.n;
        for (int s = 0; s < 2; ++s) {
          if (!valid[s]) continue;
          const int64_t row = tile_of[s] * kTileRows + r;
          const uint32_t d_tmem = tmem_base + lane_field + (uint32_t)s * kSlotCols;
          mbar_wait(&bars[BAR_ACC_FULL + s], ph_acc[s]);
          ph_acc[s] ^= 1;
          tc_fence_after();
          float f[32];
          if (active) {
            uint32_t v[32];
            tmem_ld32(d_tmem + (uint32_t)col, v);
            tmem_ld_wait();
            if (ly.row_bias) {
              int64_t ray = a.row_order == NVSR_ROWS_BLOCKED
                                ? (tile_of[s] / a.tiles_per_blk) * kBlkRays + (r & (kBlkRays - 1))
                                : row / a.samples_per_ray;
              if (ray >= a.n_rays) ray = a.n_rays - 1;
              const float4* rb = reinterpret_cast<const float4*>(ly.row_bias + ray * ly.n + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 b4 = __ldg(rb + j);
                f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
                f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
                f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
              }
            } else {
              const float4* sb = reinterpret_cast<const float4*>(sbias + l * 128 + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 b4 = sb[j];
                f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
                f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
                f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
              }
            }
          }
          if (last) {
            // the accumulator has been read: the slot may start its next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_ACC_FREE + s]);
          } else {
            if (active) {
              uint32_t pk[16];
              if (ly.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(f[2 * j], f[2 * j + 1]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(f[2 * j], f[2 * j + 1]);
              }
              tmem_st16(d_tmem + kSlotAOff + (uint32_t)(col >> 1), pk);
              tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_ACT_READY + s]);
          }
          if (ly.head_w) {
            float hacc[4] = {0.f, 0.f, 0.f, 0.f};
            if (active) {
              if (ly.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              const float* hw = sheadw + (ly.head_slot * 4) * 128 + col;
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                if (h < ly.head_n) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    float4 w4 = *reinterpret_cast<const float4*>(hw + h * 128 + j);
                    hacc[h] = fmaf(f[j + 0], w4.x, hacc[h]);
                    hacc[h] = fmaf(f[j + 1], w4.y, hacc[h]);
                    hacc[h] = fmaf(f[j + 2], w4.z, hacc[h]);
                    hacc[h] = fmaf(f[j + 3], w4.w, hacc[h]);
                  }
                }
              }
            }
            // combine the four column quarters of a row: quarters 1..3 -> smem -> quarter 0
            float* hp = shpart + head_parity * (3 * 128 * 4);
            if (qtr > 0)
              *reinterpret_cast<float4*>(hp + ((qtr - 1) * 128 + r) * 4) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
            named_bar_sync(1 + quad, 4 * 32);  // the four warps that share this lane quadrant
            if (qtr == 0 && row < a.rows) {
              float4 o1 = *reinterpret_cast<const float4*>(hp + (0 * 128 + r) * 4);
              float4 o2 = *reinterpret_cast<const float4*>(hp + (1 * 128 + r) * 4);
              float4 o3 = *reinterpret_cast<const float4*>(hp + (2 * 128 + r) * 4);
              float hv[4] = {(hacc[0] + o1.x) + (o2.x + o3.x), (hacc[1] + o1.y) + (o2.y + o3.y),
                             (hacc[2] + o1.z) + (o2.z + o3.z), (hacc[3] + o1.w) + (o2.w + o3.w)};
              for (int h = 0; h < ly.head_n; ++h)
                a.raw[(int64_t)(ly.head_ch + h) * a.raw_stride + row] = hv[h] + __ldg(ly.head_b + h);
            }
            head_parity ^= 1;
          }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int32_t launch_mlp_tc(const nvsr_mlp_t* m, cudaStream_t st) {
  TcArgs a;
  a.n_layers = m->n_layers;
  uint32_t off = 0;
  int heads = 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const nvsr_layer_t& L = m->layer[l];
    if (L.k <= 0 || (L.k % 16) != 0 || L.k > 256) return NVSR_ERR_UNSUPPORTED;
    if (L.n_out < 64 || L.n_out > 128 || (L.n_out % 64) != 0) return NVSR_ERR_UNSUPPORTED;
    if (l > 0 && L.k > 128) return NVSR_ERR_UNSUPPORTED;  // hidden activations live in 64 TMEM columns
    if (!L.w || (!L.bias && !L.row_bias)) return NVSR_ERR_INVALID_ARG;
    if (!aligned16(L.w) || (L.row_bias && !aligned16(L.row_bias))) return NVSR_ERR_ALIGNMENT;
    if (l > 0 && L.k != m->layer[l - 1].n_out) return NVSR_ERR_INVALID_ARG;
    TcLayer& t = a.layer[l];
    t.w = L.w, t.bias = L.bias, t.row_bias = L.row_bias, t.head_w = L.head_w, t.head_b = L.head_b;
    t.k = L.k, t.n = L.n_out, t.relu = L.relu, t.head_n = L.head_n, t.head_ch = L.head_ch, t.head_slot = 0;
    if (L.head_w) {
      if (L.head_n <= 0 || L.head_n > 4 || !L.head_b || L.head_ch < 0 || L.head_ch + L.head_n > 4) return NVSR_ERR_INVALID_ARG;
      if (heads >= kTcMaxHeads) return NVSR_ERR_UNSUPPORTED;
      t.head_slot = heads++;
    }
    t.w_off = off;
    off += (uint32_t)(L.k * L.n_out * 2);
  }
  if (!m->layer[m->n_layers - 1].head_w) return NVSR_ERR_INVALID_ARG;  // the chain must end in a head
  a.w_bytes_total = off;
  const uint32_t in_bytes = (uint32_t)m->layer[0].k * 256u;  // 128 rows * K * 2 B
  a.in_off[0] = off, off += in_bytes;
  a.in_off[1] = off, off += in_bytes;
  a.bias_off = off, off += (uint32_t)m->n_layers * 128u * 4u;
  a.headw_off = off, off += kTcMaxHeads * 4u * 128u * 4u;
  a.hpart_off = off, off += 2u * 3u * 128u * 4u * 4u;
  a.bar_off = off, off += BAR_COUNT * 8u + 16u;
  const uint32_t smem_bytes = off;
  if (smem_bytes > 227u * 1024u) return NVSR_ERR_RESOURCE;
  if (!aligned16(m->in)) return NVSR_ERR_ALIGNMENT;

  a.in = (const uint8_t*)m->in;
  a.in_bytes = in_bytes;
  a.rows = m->rows;
  a.n_tiles = ceil_div64(m->rows, kTileRows);
  a.samples_per_ray = m->samples_per_ray > 0 ? m->samples_per_ray : 1;
  a.row_order = m->row_order;
  a.tiles_per_blk = tiles_per_block(a.samples_per_ray);
  if (a.row_order == NVSR_ROWS_BLOCKED && (m->rows % kTileRows) != 0) return NVSR_ERR_INVALID_ARG;
  a.n_rays = m->n_rays > 0 ? m->n_rays : 1;
  a.raw = m->raw;
  a.raw_stride = m->raw_stride;

  auto kernel = m->precision == NVSR_F16 ? mlp_chain_tc_kernel<true> : mlp_chain_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int32_t)e;
  int64_t grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  kernel<<<(unsigned)grid, kTcThreads, smem_bytes, st>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

}  // namespace nvsr
