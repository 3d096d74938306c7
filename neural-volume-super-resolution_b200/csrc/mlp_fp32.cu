// a6: decoder MLP chain, fp32 SIMT parity kernel (the "1e-3" mode of the precision contract).
// One CTA owns 64 rows and walks the whole chain with activations resident in shared memory, so
// no hidden activation ever touches HBM (the reference round-trips every [n,128] activation and
// launches ~20 kernels per chunk).  Weights stream through a small smem stage in K-chunks.
#include "common.cuh"

namespace nvsr {

constexpr int kF32Rows = 64;
constexpr int kF32Threads = 256;
constexpr int kF32MaxWidth = 256;
constexpr int kF32Ld = kF32MaxWidth + 4;
constexpr int kF32Kc = 16;
constexpr int kF32WLd = 128 + 1;

struct MlpF32Args {
  nvsr_layer_t layer[NVSR_MAX_LAYERS];
  int n_layers;
  const float* in;
  int64_t rows;
  int samples_per_ray;
  int64_t n_rays;
  float* raw;
  int64_t raw_stride;
};

__global__ void __launch_bounds__(kF32Threads)
mlp_chain_f32_kernel(const __grid_constant__ MlpF32Args a) {
  extern __shared__ float sm[];
  float* buf0 = sm;
  float* buf1 = sm + kF32Rows * kF32Ld;
  float* wc = buf1 + kF32Rows * kF32Ld;  // [kF32Kc][kF32WLd]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * kF32Rows;

  // load the input tile
  const int k0 = a.layer[0].k;
  for (int i = tid; i < kF32Rows * k0; i += kF32Threads) {
    int r = i / k0, k = i - r * k0;
    int64_t row = row0 + r;
    buf0[r * kF32Ld + k] = row < a.rows ? __ldg(a.in + row * k0 + k) : 0.f;
  }
  __syncthreads();

  float* cur = buf0;
  float* nxt = buf1;
  for (int l = 0; l < a.n_layers; ++l) {
    const nvsr_layer_t& L = a.layer[l];
    const int K = L.k, N = L.n_out;
    const int ncol = N >> 4;  // columns per thread (N multiple of 16, <= 128)
    const float* W = (const float*)L.w;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int kb = 0; kb < K; kb += kF32Kc) {
      // stage W[:, kb:kb+Kc] transposed into wc[kk][n]
      for (int i = tid; i < N * kF32Kc; i += kF32Threads) {
        int n = i / kF32Kc, kk = i - n * kF32Kc;
        int k = kb + kk;
        wc[kk * kF32WLd + n] = k < K ? __ldg(W + (int64_t)n * K + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kF32Kc; ++kk) {
        int k = kb + kk;
        if (k >= K) break;
        float av[4], bv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = cur[(ty * 4 + i) * kF32Ld + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = j < ncol ? wc[kk * kF32WLd + tx + 16 * j] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
    // epilogue: bias (global or per-ray), activation, to the other buffer
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = ty * 4 + i;
      int64_t row = row0 + r;
      const float* rb = nullptr;
      if (L.row_bias) {
        int64_t ray = row / a.samples_per_ray;
        if (ray >= a.n_rays) ray = a.n_rays - 1;
        rb = L.row_bias + ray * N;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ncol) {
          int n = tx + 16 * j;
          float v = acc[i][j] + (rb ? __ldg(rb + n) : __ldg(L.bias + n));
          if (L.relu) v = fmaxf(v, 0.f);
          nxt[r * kF32Ld + n] = v;
        }
      }
    }
    __syncthreads();
    if (L.head_w) {
      for (int i = tid; i < kF32Rows * L.head_n; i += kF32Threads) {
        int r = i / L.head_n, h = i - r * L.head_n;
        int64_t row = row0 + r;
        if (row < a.rows) {
          float s = __ldg(L.head_b + h);
          const float* hw = L.head_w + (int64_t)h * N;
          for (int n = 0; n < N; ++n) s = fmaf(nxt[r * kF32Ld + n], __ldg(hw + n), s);
          a.raw[(int64_t)(L.head_ch + h) * a.raw_stride + row] = s;
        }
      }
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
    // (the next layer's first __syncthreads after staging W orders head reads vs. overwrites:
    //  nxt (old cur) is only written in the next epilogue, after at least one barrier)
  }
}

int32_t launch_mlp_f32(const nvsr_mlp_t* m, cudaStream_t st) {
  MlpF32Args a;
  a.n_layers = m->n_layers;
  for (int l = 0; l < m->n_layers; ++l) {
    const nvsr_layer_t& L = m->layer[l];
    if (L.k <= 0 || L.k > kF32MaxWidth || L.n_out <= 0 || L.n_out > 128 || (L.n_out % 16) != 0) return NVSR_ERR_UNSUPPORTED;
    if (!L.w || (!L.bias && !L.row_bias)) return NVSR_ERR_INVALID_ARG;
    if (l > 0 && L.k != m->layer[l - 1].n_out) return NVSR_ERR_INVALID_ARG;
    if (L.head_w && (L.head_n <= 0 || !L.head_b || L.head_ch < 0 || L.head_ch + L.head_n > 4)) return NVSR_ERR_INVALID_ARG;
    a.layer[l] = L;
  }
  a.in = (const float*)m->in;
  a.rows = m->rows;
  a.samples_per_ray = m->samples_per_ray > 0 ? m->samples_per_ray : 1;
  a.n_rays = m->n_rays > 0 ? m->n_rays : 1;
  a.raw = m->raw;
  a.raw_stride = m->raw_stride;
  size_t smem = (size_t)(2 * kF32Rows * kF32Ld + kF32Kc * kF32WLd) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mlp_chain_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int32_t)e;
  int64_t blocks = ceil_div64(m->rows, kF32Rows);
  if (blocks >= ((int64_t)1 << 31)) return NVSR_ERR_INVALID_ARG;
  mlp_chain_f32_kernel<<<(unsigned)blocks, kF32Threads, smem, st>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

}  // namespace nvsr
