// Ray generation / preparation, plane + weight re-packing, per-ray view-direction features.
// Small elementwise kernels: one launch each instead of the reference's ~10 ATen kernels per call.
#include "common.cuh"
#include "bilinear.cuh"

namespace nvsr {

struct Mat4 {
  float m[16];
};

// a1: nerf_helpers.py:530-549.  No FMA contraction: the reference rounds every op separately.
// c2w_dev != NULL: the pose is read from device memory (row-major 4x4 fp32) instead of the by-value copy, so a
// caller holding the pose on the GPU (the reference does, train_nerf.py:659) needs no device->host round trip.
__global__ void ray_bundle_kernel(int H, int W, float fx, float fy, Mat4 c2w, const float* __restrict__ c2w_dev,
                                  int padding, float offset, int row_begin, int n_rows, int Wp,
                                  float* __restrict__ ro, float* __restrict__ rd) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)n_rows * Wp;
  if (idx >= total) return;
  if (c2w_dev) {
#pragma unroll
    for (int i = 0; i < 12; ++i) c2w.m[i] = __ldg(c2w_dev + i);
  }
  int r = (int)(idx / Wp) + row_begin;
  int c = (int)(idx % Wp);
  // ii = (arange(W+2p) + offset) - p ; jj likewise
  float ii = __fadd_rn((float)c, offset);
  float jj = __fadd_rn((float)r, offset);
  if (padding > 0) {
    ii = __fsub_rn(ii, (float)padding);
    jj = __fsub_rn(jj, (float)padding);
  }
  float dx = __fdiv_rn(__fsub_rn(ii, (float)W * 0.5f), fx);
  float dy = -__fdiv_rn(__fsub_rn(jj, (float)H * 0.5f), fy);
  float dz = -1.0f;
  float* o = ro + idx * 3;
  float* d = rd + idx * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w.m[k * 4 + 0]), __fmul_rn(dy, c2w.m[k * 4 + 1])),
                        __fmul_rn(dz, c2w.m[k * 4 + 2]));
    d[k] = s;
    o[k] = c2w.m[k * 4 + 3];
  }
}

// a2/a3: train_utils.py:210-226 + nerf_helpers.py:578-605
__global__ void prepare_rays_kernel(const float* __restrict__ ro_in, const float* __restrict__ rd_in,
                                    int64_t n, int use_ndc, float c0, float c1, float near_f,
                                    float two_near, float neg_two_near, float* __restrict__ ro_out,
                                    float* __restrict__ rd_out, float* __restrict__ viewdirs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ox = ro_in[i * 3 + 0], oy = ro_in[i * 3 + 1], oz = ro_in[i * 3 + 2];
  float dx = rd_in[i * 3 + 0], dy = rd_in[i * 3 + 1], dz = rd_in[i * 3 + 2];
  if (viewdirs) {
    float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    viewdirs[i * 3 + 0] = __fdiv_rn(dx, nrm);
    viewdirs[i * 3 + 1] = __fdiv_rn(dy, nrm);
    viewdirs[i * 3 + 2] = __fdiv_rn(dz, nrm);
  }
  if (use_ndc) {
    float t = __fdiv_rn(-__fadd_rn(near_f, oz), dz);
    ox = __fadd_rn(ox, __fmul_rn(t, dx));
    oy = __fadd_rn(oy, __fmul_rn(t, dy));
    oz = __fadd_rn(oz, __fmul_rn(t, dz));
    float o0 = __fdiv_rn(__fmul_rn(c0, ox), oz);
    float o1 = __fdiv_rn(__fmul_rn(c1, oy), oz);
    float o2 = __fadd_rn(1.0f, __fdiv_rn(two_near, oz));
    float d0 = __fmul_rn(c0, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(ox, oz)));
    float d1 = __fmul_rn(c1, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(oy, oz)));
    float d2 = __fdiv_rn(neg_two_near, oz);
    ox = o0, oy = o1, oz = o2, dx = d0, dy = d1, dz = d2;
  }
  ro_out[i * 3 + 0] = ox, ro_out[i * 3 + 1] = oy, ro_out[i * 3 + 2] = oz;
  rd_out[i * 3 + 0] = dx, rd_out[i * 3 + 1] = dy, rd_out[i * 3 + 2] = dz;
}

// NCHW fp32 -> HWC (fp32|bf16) through a 32x33 smem transpose: coalesced on both sides.
template <typename OutT>
__global__ void pack_plane_kernel(const float* __restrict__ src, int C, int64_t HW, OutT* __restrict__ dst) {
  __shared__ float tile[32][33];
  int64_t p0 = (int64_t)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j;
    int64_t p = p0 + tx;
    tile[j][tx] = (c < C && p < HW) ? src[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int64_t p = p0 + j;
    int c = c0 + tx;
    if (c < C && p < HW) {
      float v = tile[tx][j];
      dst[p * C + c] = OutT(v);
    }
  }
}

// NCHW fp32 -> 16-bit "x-pair record" image [Rh][C/8][Rw][2][8]: the 32-byte record (y, c, x) holds the
// 8-channel chunk c of texel (y, x) and of its right neighbour (y, min(x+1, Rw-1)), so a bilinear
// footprint row is ONE 256-bit load and neighbouring rays share 128-byte lines (csrc/gather.cu).
// One thread per record: 8 coalesced-along-x reads (the neighbour's come from L1), two 16-byte stores.
// Values beyond the format's finite range saturate (the interpolation is a convex combination, so
// features stay finite).
template <bool F16>
__global__ void pack_plane16_kernel(const float* __restrict__ src, int C, int rh, int rw, uint4* __restrict__ dst) {
  const int CH = C / 8;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)rh * CH * rw;
  if (idx >= total) return;
  int x = (int)(idx % rw);
  int c8 = (int)((idx / rw) % CH);
  int y = (int)(idx / ((int64_t)rw * CH));
  const int64_t HW = (int64_t)rh * rw;
  const float* s = src + (int64_t)(c8 * 8) * HW + (int64_t)y * rw;
  const int xr = min(x + 1, rw - 1);
  float v[8], w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = s[e * HW + x], w[e] = s[e * HW + xr];
  uint4 o, q;
  o.x = pack16x2<F16>(v[0], v[1]), o.y = pack16x2<F16>(v[2], v[3]);
  o.z = pack16x2<F16>(v[4], v[5]), o.w = pack16x2<F16>(v[6], v[7]);
  q.x = pack16x2<F16>(w[0], w[1]), q.y = pack16x2<F16>(w[2], w[3]);
  q.z = pack16x2<F16>(w[4], w[5]), q.w = pack16x2<F16>(w[6], w[7]);
  dst[idx * 2] = o;
  dst[idx * 2 + 1] = q;
}

// W [n_out,k] (ld) fp32 -> 16-bit image [k_pad/8][n_out][8]
template <typename OutT>
__global__ void pack_weight_kernel(const float* __restrict__ w, int n_out, int k, int ldw, int k_pad,
                                   OutT* __restrict__ dst) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = n_out * k_pad;
  if (idx >= total) return;
  int e = idx & 7;
  int n = (idx >> 3) % n_out;
  int chunk = (idx >> 3) / n_out;
  int kk = chunk * 8 + e;
  float v = kk < k ? w[(int64_t)n * ldw + kk] : 0.f;
  dst[idx] = OutT(v);
}

// up to 8 weights in one launch (blockIdx.y = which); optionally the maximum |w| of everything packed is folded into
// *absmax (non-negative floats order like their bit patterns; a NaN compares above +Inf) — the fp16 range check of a
// training step without separate reduction kernels
struct PackWeightsArgs {
  const float* w[8];
  void* dst[8];
  int n_out[8], k[8], ldw[8], k_pad[8];
  float* absmax;
};
template <typename OutT>
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackWeightsArgs a) {
  const int i = blockIdx.y;
  const int n_out = a.n_out[i], k = a.k[i], k_pad = a.k_pad[i];
  const float* __restrict__ w = a.w[i];
  OutT* __restrict__ dst = reinterpret_cast<OutT*>(a.dst[i]);
  const int total = n_out * k_pad;
  float m = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7;
    const int n = (idx >> 3) % n_out;
    const int kk = ((idx >> 3) / n_out) * 8 + e;
    const float v = kk < k ? w[(int64_t)n * a.ldw[i] + kk] : 0.f;
    dst[idx] = OutT(v);
    const float av = fabsf(v);
    m = (av > m || av != av) ? av : m;
  }
  if (a.absmax) {
    int bits = __float_as_int(m) & 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
    if ((threadIdx.x & 31) == 0 && bits > 0) atomicMax(reinterpret_cast<int*>(a.absmax), bits);
  }
}

// a5 (view half): one thread per (ray, 4-channel chunk)
__global__ void viewdir_gather_kernel(const float* __restrict__ viewdirs, int64_t n, const float* __restrict__ vplane,
                                      int rh, int rw, int C, float az_lo, float az_rng, float el_lo, float el_rng,
                                      float* __restrict__ vfeat) {
  int chunks = C / 4;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * chunks) return;
  int64_t ray = idx / chunks;
  int ch = (int)(idx % chunks) * 4;
  float dx = viewdirs[ray * 3 + 0], dy = viewdirs[ray * 3 + 1], dz = viewdirs[ray * 3 + 2];
  // cart2az_el: el = atan2(z, sqrt(x^2+y^2)), az = atan2(y,x)
  float el = atan2f(dz, __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
  float az = atan2f(dy, dx);
  // normalize_coords: 2*(c-lo)/(hi-lo)-1
  float gaz = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(az, az_lo)), az_rng), 1.f);
  float gel = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(el, el_lo)), el_rng), 1.f);
  Bilin b = bilinear_setup(gaz, gel, rw, rh);
  const float4 v00 = *reinterpret_cast<const float4*>(vplane + ((int64_t)b.y0 * rw + b.x0) * C + ch);
  const float4 v01 = *reinterpret_cast<const float4*>(vplane + ((int64_t)b.y0 * rw + b.x1) * C + ch);
  const float4 v10 = *reinterpret_cast<const float4*>(vplane + ((int64_t)b.y1 * rw + b.x0) * C + ch);
  const float4 v11 = *reinterpret_cast<const float4*>(vplane + ((int64_t)b.y1 * rw + b.x1) * C + ch);
  float4 o;
  o.x = v00.x * b.w00 + v01.x * b.w01 + v10.x * b.w10 + v11.x * b.w11;
  o.y = v00.y * b.w00 + v01.y * b.w01 + v10.y * b.w10 + v11.y * b.w11;
  o.z = v00.z * b.w00 + v01.z * b.w01 + v10.z * b.w10 + v11.z * b.w11;
  o.w = v00.w * b.w00 + v01.w * b.w01 + v10.w * b.w10 + v11.w * b.w11;
  *reinterpret_cast<float4*>(vfeat + ray * C + ch) = o;
}

// out[ray,n] = b[n] + sum_k w[n,k]*vin[ray,k].  Persistent blocks: the weights are transposed into
// shared memory once per block (coalesced reads, padded rows -> conflict-free), then every iteration
// stages 64 rays of vin and each thread produces 8 rays x 1 output column from LDS.128 operands.
constexpr int kRbRays = 64;
__global__ void __launch_bounds__(512)
row_bias_kernel(const float* __restrict__ vin, int64_t n_rays, int K, int Kp, const float* __restrict__ w, int ldw,
                const float* __restrict__ b, int n_out, float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];  // [Kp][n_out+1] transposed weights | [kRbRays][Kp] vin tile
  const int ldn = n_out + 1;
  float* sv = sw + (size_t)Kp * ldn + ((4 - ((Kp * ldn) & 3)) & 3);
  for (int i = threadIdx.x; i < n_out * Kp; i += blockDim.x) {
    int n = i / Kp, k = i - n * Kp;
    sw[k * ldn + n] = k < K ? w[(int64_t)n * ldw + k] : 0.f;
  }
  const int groups = kRbRays / 8;  // 8 rays per thread
  for (int64_t r0 = (int64_t)blockIdx.x * kRbRays; r0 < n_rays; r0 += (int64_t)gridDim.x * kRbRays) {
    __syncthreads();
    for (int i = threadIdx.x; i < kRbRays * Kp; i += blockDim.x) {
      int rr = i / Kp, k = i - rr * Kp;
      int64_t ray = r0 + rr;
      sv[i] = (k < K && ray < n_rays) ? __ldg(vin + ray * K + k) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < groups * n_out; i += blockDim.x) {
      int n = i % n_out, g = i / n_out;
      float acc[8];
      float bn = b ? __ldg(b + n) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bn;
      for (int k = 0; k < Kp; k += 4) {
        float w0 = sw[k * ldn + n], w1 = sw[(k + 1) * ldn + n], w2 = sw[(k + 2) * ldn + n], w3 = sw[(k + 3) * ldn + n];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = *reinterpret_cast<const float4*>(sv + (g * 8 + j) * Kp + k);
          acc[j] = fmaf(w0, v.x, acc[j]);
          acc[j] = fmaf(w1, v.y, acc[j]);
          acc[j] = fmaf(w2, v.z, acc[j]);
          acc[j] = fmaf(w3, v.w, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int64_t ray = r0 + g * 8 + j;
        if (ray < n_rays) out[ray * n_out + n] = acc[j];
      }
    }
  }
}


// Same product for the shapes of the hot path (n_out a multiple of 4, n_out <= 128): one warp per group of 8
// consecutive rays, lane l owns output columns 4l..4l+3, so a ray's row is written as one coalesced 512-byte
// store per warp.  Weights sit transposed in shared memory ([k][n_out]: one conflict-free LDS.128 per lane and
// k), the group's 8 x K inputs are staged transposed per warp ([k][8 rays]: two broadcast LDS.128 per k), the
// accumulation runs over k in order with fused multiply-adds — bit-identical to row_bias_kernel.  HBM-bound:
// 4 K + 4 n_out bytes per ray.
constexpr int kRb2Warps = 8;
__global__ void __launch_bounds__(kRb2Warps * 32)
row_bias_warp_kernel(const float* __restrict__ vin, int64_t n_rays, int K, const float* __restrict__ w, int ldw,
                     const float* __restrict__ b, int n_out, float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];  // [K][n_out] transposed weights | per warp [K][8] inputs
  float* sv = sw + (size_t)K * n_out + (threadIdx.x >> 5) * (K * 8);
  for (int i = threadIdx.x; i < n_out * K; i += blockDim.x) {
    int n = i / K, k = i - n * K;
    sw[k * n_out + n] = w[(int64_t)n * ldw + k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int col = 4 * lane;
  const bool active = col < n_out;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b && active) bias = make_float4(__ldg(b + col), __ldg(b + col + 1), __ldg(b + col + 2), __ldg(b + col + 3));
  const int64_t n_groups = ceil_div64(n_rays, 8);
  for (int64_t g = (int64_t)blockIdx.x * kRb2Warps + (threadIdx.x >> 5); g < n_groups; g += (int64_t)gridDim.x * kRb2Warps) {
    const int64_t r0 = g * 8;
    __syncwarp();
    // the group's inputs are 8*K contiguous floats: coalesced read, transposed store
    for (int i = lane; i < 8 * K; i += 32) {
      int j = i / K, k = i - j * K;
      sv[k * 8 + j] = (r0 + j < n_rays) ? __ldg(vin + r0 * K + i) : 0.f;
    }
    __syncwarp();
    float4 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias;
    if (active) {
      for (int k = 0; k < K; ++k) {
        const float4 wk = *reinterpret_cast<const float4*>(sw + k * n_out + col);
        const float4 va = *reinterpret_cast<const float4*>(sv + k * 8);
        const float4 vb = *reinterpret_cast<const float4*>(sv + k * 8 + 4);
        const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j].x = fmaf(wk.x, v[j], acc[j].x), acc[j].y = fmaf(wk.y, v[j], acc[j].y);
          acc[j].z = fmaf(wk.z, v[j], acc[j].z), acc[j].w = fmaf(wk.w, v[j], acc[j].w);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < n_rays) *reinterpret_cast<float4*>(out + (r0 + j) * n_out + col) = acc[j];
    }
  }
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_ray_bundle(int32_t height, int32_t width, float focal_x, float focal_y,
                                   const float* c2w_host, int32_t padding, float offset, int32_t row_begin,
                                   int32_t row_end, float* ro, float* rd, void* stream) {
  NVSR_CHECK_ARG(height > 0 && width > 0 && c2w_host && ro && rd && padding >= 0);
  NVSR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= height + 2 * padding);
  int n_rows = row_end - row_begin;
  if (n_rows == 0) return NVSR_OK;
  Mat4 m;
  for (int i = 0; i < 16; ++i) m.m[i] = c2w_host[i];
  int Wp = width + 2 * padding;
  int64_t total = (int64_t)n_rows * Wp;
  int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  ray_bundle_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(height, width, focal_x, focal_y, m, nullptr,
                                                                          padding, offset, row_begin, n_rows, Wp, ro, rd);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_ray_bundle_dev(int32_t height, int32_t width, float focal_x, float focal_y,
                                       const float* c2w_device, int32_t padding, float offset, int32_t row_begin,
                                       int32_t row_end, float* ro, float* rd, void* stream) {
  NVSR_CHECK_ARG(height > 0 && width > 0 && c2w_device && ro && rd && padding >= 0);
  NVSR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= height + 2 * padding);
  int n_rows = row_end - row_begin;
  if (n_rows == 0) return NVSR_OK;
  Mat4 m = {};
  int Wp = width + 2 * padding;
  int64_t total = (int64_t)n_rows * Wp;
  int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  ray_bundle_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(height, width, focal_x, focal_y, m, c2w_device,
                                                                          padding, offset, row_begin, n_rows, Wp, ro, rd);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_prepare_rays(const float* ro_in, const float* rd_in, int64_t n_rays, int32_t use_ndc,
                                     int32_t height, int32_t width, double focal, double ndc_near, float* ro_out,
                                     float* rd_out, float* viewdirs, void* stream) {
  NVSR_CHECK_ARG(ro_in && rd_in && ro_out && rd_out && n_rays >= 0);
  if (n_rays == 0) return NVSR_OK;
  // python-side scalars are evaluated in double then cast to fp32 by the tensor op
  float c0 = 0.f, c1 = 0.f;
  if (use_ndc) {
    NVSR_CHECK_ARG(focal != 0.0 && height > 0 && width > 0);
    c0 = (float)(-1.0 / ((double)width / (2.0 * focal)));
    c1 = (float)(-1.0 / ((double)height / (2.0 * focal)));
  }
  int threads = 256;
  int64_t blocks = ceil_div64(n_rays, threads);
  prepare_rays_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      ro_in, rd_in, n_rays, use_ndc, c0, c1, (float)ndc_near, (float)(2.0 * ndc_near), (float)(-2.0 * ndc_near),
      ro_out, rd_out, viewdirs);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_pack_plane(const float* src_nchw, int32_t channels, int32_t rh, int32_t rw, void* dst,
                                   int32_t dst_dtype, void* stream) {
  NVSR_CHECK_ARG(src_nchw && dst && channels > 0 && rh > 0 && rw > 0);
  NVSR_CHECK_ARG(dst_dtype == NVSR_F32 || dst_dtype == NVSR_BF16 || dst_dtype == NVSR_F16);
  int64_t HW = (int64_t)rh * rw;
  dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)((channels + 31) / 32));
  dim3 block(32, 8);
  if (dst_dtype == NVSR_F32) {
    pack_plane_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(src_nchw, channels, HW, (float*)dst);
  } else {
    NVSR_CHECK_ARG(channels % 8 == 0);
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) != 0) return NVSR_ERR_ALIGNMENT;
    int64_t total = HW * (channels / 8);
    unsigned blocks = (unsigned)ceil_div64(total, 256);
    if (dst_dtype == NVSR_BF16)
      pack_plane16_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src_nchw, channels, rh, rw, (uint4*)dst);
    else
      pack_plane16_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src_nchw, channels, rh, rw, (uint4*)dst);
  }
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_pack_weight16(const float* w, int32_t n_out, int32_t k, int32_t ldw, int32_t k_pad,
                                      void* dst, int32_t dst_dtype, void* stream) {
  NVSR_CHECK_ARG(w && dst && n_out > 0 && k > 0 && ldw >= k && k_pad >= k && (k_pad % 16) == 0);
  NVSR_CHECK_ARG(dst_dtype == NVSR_BF16 || dst_dtype == NVSR_F16);
  int total = n_out * k_pad;
  if (dst_dtype == NVSR_BF16)
    pack_weight_kernel<__nv_bfloat16>
        <<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_out, k, ldw, k_pad, (__nv_bfloat16*)dst);
  else
    pack_weight_kernel<__half><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_out, k, ldw, k_pad, (__half*)dst);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_pack_weights16(int32_t count, const float* const* w, const int32_t* n_out, const int32_t* k,
                                       const int32_t* ldw, const int32_t* k_pad, void* const* dst, int32_t dst_dtype,
                                       float* absmax, void* stream) {
  NVSR_CHECK_ARG(count >= 0 && (count == 0 || (w && n_out && k && ldw && k_pad && dst)));
  NVSR_CHECK_ARG(dst_dtype == NVSR_BF16 || dst_dtype == NVSR_F16);
  for (int32_t base = 0; base < count; base += 8) {
    PackWeightsArgs a;
    const int n = count - base < 8 ? count - base : 8;
    int max_total = 0;
    for (int i = 0; i < n; ++i) {
      const int j = base + i;
      NVSR_CHECK_ARG(w[j] && dst[j] && n_out[j] > 0 && k[j] > 0 && ldw[j] >= k[j] && k_pad[j] >= k[j] && (k_pad[j] % 16) == 0);
      a.w[i] = w[j], a.dst[i] = dst[j], a.n_out[i] = n_out[j], a.k[i] = k[j], a.ldw[i] = ldw[j], a.k_pad[i] = k_pad[j];
      const int total = n_out[j] * k_pad[j];
      max_total = total > max_total ? total : max_total;
    }
    a.absmax = absmax;
    const dim3 grid((unsigned)((max_total + 255) / 256), (unsigned)n);
    if (dst_dtype == NVSR_BF16) pack_weights_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else pack_weights_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  }
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_viewdir_gather(const float* viewdirs, int64_t n_rays, const float* vplane, int32_t rh,
                                       int32_t rw, int32_t channels, float az_lo, float az_rng, float el_lo,
                                       float el_rng, float* vfeat, void* stream) {
  NVSR_CHECK_ARG(viewdirs && vplane && vfeat && n_rays >= 0 && rh > 0 && rw > 0);
  NVSR_CHECK_ARG(channels > 0 && channels % 4 == 0);
  if (!aligned16(vplane) || !aligned16(vfeat)) return NVSR_ERR_ALIGNMENT;
  if (n_rays == 0) return NVSR_OK;
  int64_t total = n_rays * (channels / 4);
  viewdir_gather_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      viewdirs, n_rays, vplane, rh, rw, channels, az_lo, az_rng, el_lo, el_rng, vfeat);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_row_bias(const float* vin, int64_t n_rays, int32_t k, const float* w, int32_t ldw,
                                 const float* b, int32_t n_out, float* out, void* stream) {
  NVSR_CHECK_ARG(vin && w && out && n_rays >= 0 && k > 0 && n_out > 0 && ldw >= k);
  NVSR_CHECK_ARG((size_t)k * n_out * sizeof(float) <= 48 * 1024);
  if (n_rays == 0) return NVSR_OK;
  if ((n_out & 3) == 0 && n_out <= 128 && aligned16(out)) {
    size_t smem2 = ((size_t)k * n_out + (size_t)kRb2Warps * k * 8) * sizeof(float);
    if (smem2 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(row_bias_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return (int32_t)e;
    }
    int64_t blocks2 = ceil_div64(ceil_div64(n_rays, 8), kRb2Warps);
    if (blocks2 > 6 * kNumSMs) blocks2 = 6 * kNumSMs;
    row_bias_warp_kernel<<<(unsigned)blocks2, kRb2Warps * 32, smem2, (cudaStream_t)stream>>>(vin, n_rays, k, w, ldw, b, n_out, out);
    NVSR_RETURN_LAST_ERROR();
  }
  const int kp = (k + 3) & ~3;
  size_t smem = ((size_t)kp * (n_out + 1) + 4 + (size_t)kRbRays * kp) * sizeof(float);
  if (smem > 200 * 1024) return NVSR_ERR_RESOURCE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(row_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int32_t)e;
  }
  int64_t blocks = ceil_div64(n_rays, kRbRays);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;  // up to 4 resident CTAs per SM (37 KB smem, 512 threads each)
  row_bias_kernel<<<(unsigned)blocks, 512, smem, (cudaStream_t)stream>>>(vin, n_rays, k, kp, w, ldw, b, n_out, out);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_abi_version(void) { return NVSR_ABI_VERSION; }

extern "C" int64_t nvsr_rows_padded(int64_t n_rays, int32_t n_samples, int32_t row_order) {
  if (n_rays < 0 || n_samples <= 0) return 0;
  return rows_padded(n_rays, n_samples, row_order);
}

extern "C" const char* nvsr_status_string(int32_t status) {
  switch (status) {
    case NVSR_OK: return "ok";
    case NVSR_ERR_INVALID_ARG: return "invalid argument";
    case NVSR_ERR_UNSUPPORTED: return "unsupported configuration";
    case NVSR_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
    case NVSR_ERR_RESOURCE: return "resource limit (shared memory / samples per ray)";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown nvsr status";
  }
}
