// a6: decoder MLP chain on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Persistent kernel, one CTA per SM, cta_group::1, UMMA M=128 (one 128-row tile), N = layer width.
//   warp 0 (1 lane)  : producer — bulk-copies (TMA engine) the chain's bf16 weights into shared
//                      memory once, then one bf16 feature tile image per tile into its slot
//   warp 1 (1 lane)  : MMA issuer — for every layer issues K/16 tcgen05.mma (A = activations in smem,
//                      B = resident weights in smem, D = TMEM), commits to an mbarrier
//   warps 2..9       : epilogue — tcgen05.ld the fp32 accumulator, + bias (global or per-ray), ReLU,
//                      pack to bf16 and write the next layer's A operand back into the slot in place;
//                      output heads (128 -> 1 / 3) are fp32 dot products on the CUDA cores
// Two tiles ("slots") are in flight and ping-pong: while the epilogue warps work on slot A's layer l,
// the tensor core runs slot B's layer l, so neither side waits on the other in steady state.
// Hidden activations never leave the SM; weights are read from HBM/L2 once per CTA.
//
// Operand layout (both A and B): K-major, no-swizzle canonical UMMA layout with 8x16-byte core
// matrices, stored [K/8][rows][8 bf16]: LBO (K-chunk stride) = rows*16 B, SBO (8-row group stride)
// = 128 B.  The feature tile image written by the gather kernel and the weight image written by
// nvsr_pack_weight_bf16 are exactly this, so both arrive with plain bulk copies.
#include "common.cuh"

namespace nvsr {

constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr int kTcMaxHeads = 2;
constexpr uint32_t kTmemCols = 256;  // 2 slots x 128 fp32 columns

struct TcLayer {
  const void* w;          // global bf16 image
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
  int samples_per_ray;
  int64_t n_rays;
  float* raw;
  int64_t raw_stride;
  // smem carve-up (byte offsets from the 1024-aligned base)
  uint32_t act_off[2], bias_off, headw_off, hpart_off, bar_off, w_bytes_total;
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
__device__ __forceinline__ void umma_f16kind(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
  float* shpart = reinterpret_cast<float*>(smem + a.hpart_off);  // [2 parity][128][4]: half 1 -> half 0

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
        bulk_g2s(smem + a.act_off[s], a.in + tile * (int64_t)a.in_bytes, a.in_bytes, &bars[BAR_IN_FULL + s]);
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
          for (int s = 0; s < 2; ++s) {
            if (!valid[s]) continue;
            if (l == 0) {
              mbar_wait(&bars[BAR_IN_FULL + s], (uint32_t)(p & 1));
              mbar_wait(&bars[BAR_ACC_FREE + s], (uint32_t)(p & 1) ^ 1);
            } else {
              mbar_wait(&bars[BAR_ACT_READY + s], ph_act[s]);
              ph_act[s] ^= 1;
            }
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + a.act_off[s]);
            const uint32_t b_base = smem_u32(smem + ly.w_off);
            const uint32_t d_tmem = tmem_base + (uint32_t)s * 128u;
            const int ksteps = ly.k >> 4;
            for (int ks = 0; ks < ksteps; ++ks) {
              uint64_t ad = umma_desc(a_base + (uint32_t)ks * 2u * 2048u, 2048u, 128u);
              uint64_t bd = umma_desc(b_base + (uint32_t)ks * 2u * b_lbo, b_lbo, 128u);
              umma_f16kind(d_tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(&bars[BAR_ACC_FULL + s]);
            if (l == L - 1) umma_commit(&bars[BAR_IN_FREE + s]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = ew >> 2;             // which half of the layer's columns
    const int r = quad * 32 + lane;       // row within the tile == TMEM lane
    uint32_t ph_acc[2] = {0, 0};
    uint32_t head_parity = 0;
    for (int64_t p = 0;; ++p) {
      int64_t tile_of[2] = {blockIdx.x + (2 * p) * G, blockIdx.x + (2 * p + 1) * G};
      bool valid[2] = {tile_of[0] < a.n_tiles, tile_of[1] < a.n_tiles};
      if (!valid[0]) break;
      for (int l = 0; l < L; ++l) {
        const TcLayer& ly = a.layer[l];
        const int ncols = ly.n >> 1;  // columns handled by this thread
        const int col_base = half * ncols;
        for (int s = 0; s < 2; ++s) {
          if (!valid[s]) continue;
          const int64_t row = tile_of[s] * kTileRows + r;
          const float* rb = nullptr;
          if (ly.row_bias) {
            int64_t ray = row / a.samples_per_ray;
            if (ray >= a.n_rays) ray = a.n_rays - 1;
            rb = ly.row_bias + ray * ly.n;
          }
          mbar_wait(&bars[BAR_ACC_FULL + s], ph_acc[s]);
          ph_acc[s] ^= 1;
          tc_fence_after();
          float hacc[4] = {0.f, 0.f, 0.f, 0.f};
          uint8_t* act = smem + a.act_off[s];
          for (int c0 = 0; c0 < ncols; c0 += 32) {
            const int col = col_base + c0;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)s * 128u + (uint32_t)col, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = rb ? __ldg(reinterpret_cast<const float4*>(rb + col + j))
                             : *reinterpret_cast<const float4*>(sbias + l * 128 + col + j);
              f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
              f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
              f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
            if (ly.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (ly.head_w) {
              const float* hw = sheadw + (ly.head_slot * 4) * 128 + col;
              for (int h = 0; h < ly.head_n; ++h) {
#pragma unroll
                for (int j = 0; j < 32; ++j) hacc[h] = fmaf(f[j], hw[h * 128 + j], hacc[h]);
              }
            }
            if (l < L - 1) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 o;
                o.x = pack16x2<F16>(f[q * 8 + 0], f[q * 8 + 1]);
                o.y = pack16x2<F16>(f[q * 8 + 2], f[q * 8 + 3]);
                o.z = pack16x2<F16>(f[q * 8 + 4], f[q * 8 + 5]);
                o.w = pack16x2<F16>(f[q * 8 + 6], f[q * 8 + 7]);
                *reinterpret_cast<uint4*>(act + (uint32_t)((col >> 3) + q) * 2048u + (uint32_t)r * 16u) = o;
              }
            }
          }
          if (l < L - 1) {
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the UMMA (async proxy) reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_ACT_READY + s]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_ACC_FREE + s]);
          }
          if (ly.head_w) {
            float* hp = shpart + head_parity * (128 * 4);
            if (half == 1) *reinterpret_cast<float4*>(hp + r * 4) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
            named_bar_sync(1, kTcEpiWarps * 32);
            if (half == 0 && row < a.rows) {
              float4 o1 = *reinterpret_cast<const float4*>(hp + r * 4);
              float hv[4] = {hacc[0] + o1.x, hacc[1] + o1.y, hacc[2] + o1.z, hacc[3] + o1.w};
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
  int max_k = 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const nvsr_layer_t& L = m->layer[l];
    if (L.k <= 0 || (L.k % 16) != 0 || L.k > 256) return NVSR_ERR_UNSUPPORTED;
    if (L.n_out < 64 || L.n_out > 128 || (L.n_out % 64) != 0) return NVSR_ERR_UNSUPPORTED;
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
    if (L.k > max_k) max_k = L.k;
  }
  if (!m->layer[m->n_layers - 1].head_w) return NVSR_ERR_INVALID_ARG;  // the chain must end in a head
  a.w_bytes_total = off;
  uint32_t act_bytes = (uint32_t)max_k * 256u;  // 128 rows * K * 2 B
  a.act_off[0] = off, off += act_bytes;
  a.act_off[1] = off, off += act_bytes;
  a.bias_off = off, off += (uint32_t)m->n_layers * 128u * 4u;
  a.headw_off = off, off += kTcMaxHeads * 4u * 128u * 4u;
  a.hpart_off = off, off += 2u * 128u * 4u * 4u;
  a.bar_off = off, off += BAR_COUNT * 8u + 16u;
  const uint32_t smem_bytes = off;
  if (smem_bytes > 227u * 1024u) return NVSR_ERR_RESOURCE;
  if (!aligned16(m->in)) return NVSR_ERR_ALIGNMENT;

  a.in = (const uint8_t*)m->in;
  a.in_bytes = (uint32_t)m->layer[0].k * 256u;
  a.rows = m->rows;
  a.n_tiles = ceil_div64(m->rows, kTileRows);
  a.samples_per_ray = m->samples_per_ray > 0 ? m->samples_per_ray : 1;
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
