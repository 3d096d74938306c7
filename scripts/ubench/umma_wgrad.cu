// Micro-test for the decoder WEIGHT-GRADIENT MMA (NOTES.md backlog #1; round-2 tool — written without GPU access,
// never run yet).  Question it answers on the hardware: can the forward's K-major tile images be fed to tcgen05.mma
// UNCHANGED as MN-major operands of the transposed product?
//
//   dW[n_out][k_in] = sum over rows r of  dY[r][n_out] * A[r][k_in]          (nn.Linear weight gradient)
//
// Both operands have the reduction index (rows) as their slow index.  The forward's tile image of a [128 rows x C]
// matrix is [C/8][128 rows][8] 16-bit: element (r, c) at ((c/8)*128 + r)*16 B + (c%8)*2 B.  Read as the TRANSPOSED
// matrix [C x rows] this is exactly the canonical MN-major, no-swizzle UMMA layout (CUTLASS mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>, INTERLEAVE: ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO)) with T = 8 elements per 16 B):
// 8 consecutive K (rows) are 16 B apart, the next group of 8 rows follows at LBO = 128 B, the next group of 8
// MN elements (channels) at SBO = 128 rows * 16 B = 2048 B.  So: A-operand = dY tile image (M = n_out = 128),
// B-operand = activation tile image (N = k_in), a_major = b_major = 1 in the instruction descriptor, one K = 16 step
// per 16 rows (descriptor start + 256 B), 8 steps per 128-row tile, the accumulator stays in TMEM over all tiles.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I neural-volume-super-resolution_b200/csrc -I include \
//        -o scripts/ubench/umma_wgrad scripts/ubench/umma_wgrad.cu && scripts/ubench/umma_wgrad [k_in=128] [tiles_per_cta=64] [swap_lbo_sbo=0]
//
// A second kernel checks the DATA gradient the same way (dgrad_kernel below: the forward WEIGHT image read as an
// MN-major B operand).  Prints the max relative error against a CPU fp64 reference and the achieved TFLOP/s / GB/s.  A protocol bug traps
// after 4 s (mbar_wait in common.cuh) instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"

using namespace nvsr;

namespace {

constexpr int kRows = 128;       // rows per tile = K of the product per tile
constexpr int kNOut = 128;       // M of the MMA
constexpr uint32_t kLbo = 128;   // next group of 8 rows (K)
constexpr uint32_t kSbo = 2048;  // next group of 8 channels (MN)

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// kind::f16, fp16 x fp16 -> fp32, M = 128, N = n; a_mn / b_mn: operand is MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
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

// One CTA of 128 threads accumulates dW over its slice of tiles.  Thread 0 is loader AND issuer (a correctness
// test, not the final schedule): two stages, stage s is refilled once the MMAs that read it have committed.
__global__ void __launch_bounds__(128, 1)
wgrad_kernel(const uint8_t* __restrict__ dy, const uint8_t* __restrict__ act, int k_in, int tiles_per_cta,
             float* __restrict__ dw_partial, uint32_t mn_lbo, uint32_t mn_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_slot;
  const uint32_t dy_bytes = kRows * kNOut * 2, act_bytes = (uint32_t)kRows * k_in * 2;
  uint8_t* dy_s[2] = {smem, smem + dy_bytes + act_bytes};
  uint8_t* act_s[2] = {smem + dy_bytes, smem + 2 * dy_bytes + act_bytes};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1), mbar_init(&full[1], 1), mbar_init(&empty[0], 1), mbar_init(&empty[1], 1), mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;

  if (threadIdx.x == 0) {
    auto load = [&](int t) {
      int s = t & 1;
      mbar_arrive_expect_tx(&full[s], dy_bytes + act_bytes);
      bulk_g2s(dy_s[s], dy + (tile0 + t) * dy_bytes, dy_bytes, &full[s]);
      bulk_g2s(act_s[s], act + (tile0 + t) * act_bytes, act_bytes, &full[s]);
    };
    load(0);
    if (tiles_per_cta > 1) load(1);
    const uint32_t idesc = idesc_f16(k_in, true, true);
    for (int t = 0; t < tiles_per_cta; ++t) {
      int s = t & 1;
      mbar_wait(&full[s], (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint64_t a0 = smem_desc(smem_u32(dy_s[s]), mn_lbo, mn_sbo), b0 = smem_desc(smem_u32(act_s[s]), mn_lbo, mn_sbo);
#pragma unroll
      for (int ks = 0; ks < kRows / 16; ++ks)   // 16 rows per MMA: descriptor start address + 256 B (>> 4 = 16)
        umma_ss(tmem, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, (t | ks) ? 1u : 0u);
      umma_commit(&empty[s]);                   // arrives when the MMAs above have finished reading stage s
      if (t + 2 < tiles_per_cta) {
        mbar_wait(&empty[s], (uint32_t)((t >> 1) & 1));
        load(t + 2);
      }
    }
    umma_commit(&done);
  }
  __syncwarp();
  mbar_wait(&done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32w..32w+31 = rows n_out of dW; 32 columns (k_in) at a time
  float* out = dw_partial + (int64_t)blockIdx.x * kNOut * k_in + (int64_t)(warp * 32 + lane) * k_in;
  for (int c0 = 0; c0 < k_in; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32 && c0 + j < k_in; ++j) out[c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// DATA-GRADIENT of one layer with the FORWARD weight image, no transposed packing:
//   dX[r][k_in] = sum over o of dY[r][o] * W[o][k_in]
// A = dY tile image, K-major exactly as the forward's layer-0 operand (K = n_out: LBO = 128 rows * 16 B between
// K-chunks, SBO = 128 B between 8-row groups, + 2 K-chunks = 4096 B per K = 16 step).  B = W as [N = k_in][K = n_out]:
// the forward weight image [k_in/8][n_out][8] read MN-major (8 consecutive k_in in 16 B, consecutive o 16 B apart,
// LBO = 128 B, SBO = n_out * 16 B = 2048 B, + 256 B per K = 16 step).  One tile at a time, serialised: a correctness test.
__global__ void __launch_bounds__(128, 1)
dgrad_kernel(const uint8_t* __restrict__ dy, const uint8_t* __restrict__ w_img, int k_in, int tiles_per_cta,
             float* __restrict__ dx, uint32_t mn_lbo, uint32_t mn_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full, mma_done;
  __shared__ uint32_t tmem_slot;
  const uint32_t dy_bytes = kRows * kNOut * 2, w_bytes = (uint32_t)kNOut * k_in * 2;
  uint8_t* dy_s = smem;
  uint8_t* w_s = smem + dy_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1), mbar_init(&mma_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_f16(k_in, false, true);
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int64_t tile = (int64_t)blockIdx.x * tiles_per_cta + t;
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&full, dy_bytes + (t == 0 ? w_bytes : 0u));
      bulk_g2s(dy_s, dy + tile * dy_bytes, dy_bytes, &full);
      if (t == 0) bulk_g2s(w_s, w_img, w_bytes, &full);
      mbar_wait(&full, (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint64_t a0 = smem_desc(smem_u32(dy_s), kRows * 16, 128), b0 = smem_desc(smem_u32(w_s), mn_lbo, mn_sbo);
#pragma unroll
      for (int ks = 0; ks < kNOut / 16; ++ks)
        umma_ss(tmem, a0 + (uint64_t)(ks * 256), b0 + (uint64_t)(ks * 16), idesc, ks ? 1u : 0u);
      umma_commit(&mma_done);
    }
    __syncwarp();
    mbar_wait(&mma_done, (uint32_t)(t & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* out = dx + (tile * kRows + warp * 32 + lane) * k_in;     // TMEM lane = row of the tile
    for (int c0 = 0; c0 < k_in; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32 && c0 + j < k_in; ++j) out[c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // the accumulator and the dY buffer are free for the next tile
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

}  // namespace

int main(int argc, char** argv) {
  const int k_in = argc > 1 ? std::atoi(argv[1]) : 128;          // 48 | 128 | 144 (N % 16 == 0, <= 256)
  const int tiles_per_cta = argc > 2 ? std::atoi(argv[2]) : 64;
  const bool swap = argc > 3 && std::atoi(argv[3]) != 0;   // 1: exchange LBO and SBO of the MN-major descriptors
  const uint32_t mn_lbo = swap ? kSbo : kLbo, mn_sbo = swap ? kLbo : kSbo;
  const int ctas = 148;
  if (k_in % 16 || k_in > 256 || k_in < 16) return std::printf("k_in must be a multiple of 16 in [16, 256]\n"), 1;
  const int64_t tiles = (int64_t)ctas * tiles_per_cta;
  const size_t dy_elems = (size_t)tiles * kRows * kNOut, act_elems = (size_t)tiles * kRows * k_in;
  std::vector<__half> dy(dy_elems), act(act_elems);
  std::vector<float> dyf(dy_elems), actf(act_elems);   // logical [tile][row][col] copies for the reference
  uint32_t rng = 12345u;
  auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return ((rng >> 9) & 0x3FFF) / 8192.0f - 1.0f; };
  auto fill = [&](std::vector<__half>& img, std::vector<float>& ref, int C) {
    for (int64_t t = 0; t < tiles; ++t)
      for (int r = 0; r < kRows; ++r)
        for (int c = 0; c < C; ++c) {
          __half h = __float2half(rnd());
          img[(((size_t)t * (C / 8) + c / 8) * kRows + r) * 8 + c % 8] = h;      // the forward's tile image
          ref[((size_t)t * kRows + r) * C + c] = __half2float(h);
        }
  };
  fill(dy, dyf, kNOut);
  fill(act, actf, k_in);
  uint8_t *d_dy, *d_act;
  float* d_dw;
  CK(cudaMalloc(&d_dy, dy_elems * 2));
  CK(cudaMalloc(&d_act, act_elems * 2));
  CK(cudaMalloc(&d_dw, (size_t)ctas * kNOut * k_in * 4));
  CK(cudaMemcpy(d_dy, dy.data(), dy_elems * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_act, act.data(), act_elems * 2, cudaMemcpyHostToDevice));
  const size_t smem = 2 * ((size_t)kRows * kNOut * 2 + (size_t)kRows * k_in * 2) + 1024;
  CK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_kernel<<<ctas, 128, smem>>>(d_dy, d_act, k_in, tiles_per_cta, d_dw, mn_lbo, mn_sbo);
  CK(cudaDeviceSynchronize());
  std::vector<float> dw((size_t)ctas * kNOut * k_in);
  CK(cudaMemcpy(dw.data(), d_dw, dw.size() * 4, cudaMemcpyDeviceToHost));
  // reference for CTA 0 and the last CTA (fp64)
  double worst = 0.0, scale = 0.0;
  for (int b : {0, ctas - 1}) {
    for (int n = 0; n < kNOut; ++n)
      for (int k = 0; k < k_in; ++k) {
        double acc = 0.0;
        for (int64_t t = (int64_t)b * tiles_per_cta; t < (int64_t)(b + 1) * tiles_per_cta; ++t)
          for (int r = 0; r < kRows; ++r)
            acc += (double)dyf[((size_t)t * kRows + r) * kNOut + n] * (double)actf[((size_t)t * kRows + r) * k_in + k];
        double got = dw[((size_t)b * kNOut + n) * k_in + k];
        worst = std::fmax(worst, std::fabs(got - acc));
        scale = std::fmax(scale, std::fabs(acc));
      }
  }
  std::printf("[MN-major LBO %u SBO %u] k_in %d, %d tiles per CTA: max abs err %.3e (largest |dW| %.3e, relative %.2e) -> %s\n", mn_lbo, mn_sbo, k_in, tiles_per_cta, worst,
              scale, worst / scale, worst <= 1e-3 * scale ? "MN-MAJOR TILE IMAGES OK" : "MISMATCH");
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 10; ++i) wgrad_kernel<<<ctas, 128, smem>>>(d_dy, d_act, k_in, tiles_per_cta, d_dw, mn_lbo, mn_sbo);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 10;
  double flop = 2.0 * tiles * kRows * kNOut * k_in, bytes = (double)(dy_elems + act_elems) * 2;
  std::printf("%.3f ms per launch: %.1f TFLOP/s, %.1f GB/s of operand traffic\n", ms, flop / ms * 1e-9, bytes / ms * 1e-6);

  // ---- dgrad: dX = dY @ W with the forward weight image [k_in/8][n_out][8] as an MN-major B operand
  std::vector<__half> w_img((size_t)kNOut * k_in);
  std::vector<float> wf((size_t)kNOut * k_in);
  for (int o = 0; o < kNOut; ++o)
    for (int i = 0; i < k_in; ++i) {
      __half h = __float2half(rnd());
      w_img[((size_t)(i / 8) * kNOut + o) * 8 + i % 8] = h;       // nvsr_pack_weight16's image of W[o][i]
      wf[(size_t)o * k_in + i] = __half2float(h);
    }
  uint8_t* d_w;
  float* d_dx;
  const int dg_tiles = 4, dg_ctas = 8;
  CK(cudaMalloc(&d_w, w_img.size() * 2));
  CK(cudaMalloc(&d_dx, (size_t)dg_ctas * dg_tiles * kRows * k_in * 4));
  CK(cudaMemcpy(d_w, w_img.data(), w_img.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem_d = (size_t)kRows * kNOut * 2 + (size_t)kNOut * k_in * 2 + 1024;
  CK(cudaFuncSetAttribute(dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
  dgrad_kernel<<<dg_ctas, 128, smem_d>>>(d_dy, d_w, k_in, dg_tiles, d_dx, mn_lbo, mn_sbo);
  CK(cudaDeviceSynchronize());
  std::vector<float> dx((size_t)dg_ctas * dg_tiles * kRows * k_in);
  CK(cudaMemcpy(dx.data(), d_dx, dx.size() * 4, cudaMemcpyDeviceToHost));
  worst = scale = 0.0;
  for (int64_t t = 0; t < (int64_t)dg_ctas * dg_tiles; ++t)
    for (int r = 0; r < kRows; ++r)
      for (int i = 0; i < k_in; ++i) {
        double acc = 0.0;
        for (int o = 0; o < kNOut; ++o) acc += (double)dyf[((size_t)t * kRows + r) * kNOut + o] * (double)wf[(size_t)o * k_in + i];
        worst = std::fmax(worst, std::fabs((double)dx[((size_t)t * kRows + r) * k_in + i] - acc));
        scale = std::fmax(scale, std::fabs(acc));
      }
  std::printf("dgrad k_in %d: max abs err %.3e (largest |dX| %.3e, relative %.2e) -> %s\n", k_in, worst, scale, worst / scale,
              worst <= 1e-3 * scale ? "FORWARD WEIGHT IMAGE AS MN-MAJOR B OK" : "MISMATCH");
  return 0;
}
