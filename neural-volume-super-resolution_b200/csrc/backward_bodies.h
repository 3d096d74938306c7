// Per-element bodies of the backward kernels (backward of the hot path, train_nerf.py:860-916).
//
// Host/device-neutral on purpose: backward.cu instantiates them inside the CUDA kernels (one thread per ray /
// per (row, channel chunk)), and tests/hostcheck/hostcheck.cpp compiles THE SAME SOURCE with g++ and loops over
// the indices on the CPU, so the arithmetic and the indexing of the kernels are checked against autograd of the
// reference algorithm without a GPU (tests/test_backward_bodies.py).  Nothing here touches the forward kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NVSR_HD __host__ __device__ __forceinline__
#else
#define NVSR_HD inline
#endif

namespace nvsr {
namespace bwd {

// separately rounded fp32 ops: the footprint cell of the backward must be the cell the forward read
#if defined(__CUDA_ARCH__)
NVSR_HD float mul(float a, float b) { return __fmul_rn(a, b); }
NVSR_HD float add(float a, float b) { return __fadd_rn(a, b); }
NVSR_HD float sub(float a, float b) { return __fsub_rn(a, b); }
NVSR_HD float dvd(float a, float b) { return __fdiv_rn(a, b); }
NVSR_HD float sqr(float a) { return __fsqrt_rn(a); }
// scattered plane-gradient accumulation: ONE 128-bit vector reduction per corner (red.global.add.v4.f32, sm_90+)
// instead of four scalar ones; p is 16-byte aligned (channels-last accumulators, C % 4 == 0, chunk offset % 4 == 0)
NVSR_HD void accumulate4(float* p, float w, const float* g) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(w * g[0], w * g[1], w * g[2], w * g[3]));
}
#else
// host build: compiled with -ffp-contract=off, so every op is rounded separately as well
NVSR_HD float mul(float a, float b) { return a * b; }
NVSR_HD float add(float a, float b) { return a + b; }
NVSR_HD float sub(float a, float b) { return a - b; }
NVSR_HD float dvd(float a, float b) { return a / b; }
NVSR_HD float sqr(float a) { return sqrtf(a); }
NVSR_HD void accumulate4(float* p, float w, const float* g) {
  for (int c = 0; c < 4; ++c) p[c] += w * g[c];
}
#endif

// grid_sample(bilinear, align_corners=True, padding_mode='border') footprint — mirror of bilinear.cuh
// (models.py:303,320).  d out / d texel(y,x) = the corner weight; the coordinate itself carries no gradient on
// this path (ray geometry is data, not a parameter).
struct Foot {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};
NVSR_HD Foot footprint(float gx, float gy, int Wd, int Hd) {
  float ix = mul(mul(add(gx, 1.f), 0.5f), (float)(Wd - 1));
  float iy = mul(mul(add(gy, 1.f), 0.5f), (float)(Hd - 1));
  ix = fminf(fmaxf(ix, 0.f), (float)(Wd - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hd - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  Foot f;
  f.w00 = mul(sub(fx1, ix), sub(fy1, iy));
  f.w01 = mul(sub(ix, fx0), sub(fy1, iy));
  f.w10 = mul(sub(fx1, ix), sub(iy, fy0));
  f.w11 = mul(sub(ix, fx0), sub(iy, fy0));
  f.x0 = (int)fx0, f.y0 = (int)fy0;
  f.x1 = f.x0 + 1 < Wd ? f.x0 + 1 : Wd - 1;
  f.y1 = f.y0 + 1 < Hd ? f.y0 + 1 : Hd - 1;
  if (f.x0 + 1 > Wd - 1) f.w01 = 0.f, f.w11 = 0.f;   // out-of-range corners contribute nothing in ATen
  if (f.y0 + 1 > Hd - 1) f.w10 = 0.f, f.w11 = 0.f;
  return f;
}

NVSR_HD float box_normalize(float c, float lo, float rng) { return sub(dvd(mul(2.f, sub(c, lo)), rng), 1.f); }

struct PlaneGeom {   // the geometry half of nvsr_planes_t
  int rh[3], rw[3];
  int C;
  float lo[3], rng[3];
  float proj[3][6];
};

// ------------------------------------------------------------------------------------------------------------
// backward of the tri-plane gather (models.py:289-310 project_xyz, :355-361 combine_pos_planes('avg')):
//   featP[row, d*C + c] = bilinear(plane_d)(p)[c],   featM[row, c] = (sum_d featP[row, d*C + c]) / 3
//   =>  d plane_d[y,x,c] += w_corner * (dP[row, d*C + c] + dM[row, c] / 3)
// One call handles channels [ch, ch+4) of one row.  d_plane[d]: channels-last fp32 [rh][rw][C] accumulators.
// d_feat_p / d_feat_m: row-major fp32 [rows,3C] / [rows,C], either may be NULL.
NVSR_HD void gather_bwd_row(const PlaneGeom& g, const float* ro, const float* rd, float z, int64_t ray, int64_t row,
                            int ch, const float* d_feat_p, const float* d_feat_m, float* const d_plane[3]) {
  float n[3];
  for (int k = 0; k < 3; ++k) {
    float x = add(ro[ray * 3 + k], mul(rd[ray * 3 + k], z));   // pts = ro + rd*z (train_utils.py:111)
    n[k] = box_normalize(x, g.lo[k], g.rng[k]);
  }
  float gm[4] = {0.f, 0.f, 0.f, 0.f};
  if (d_feat_m)
    for (int c = 0; c < 4; ++c) gm[c] = d_feat_m[row * g.C + ch + c] / 3.f;
  for (int d = 0; d < 3; ++d) {
    float gx = n[0] * g.proj[d][0] + n[1] * g.proj[d][2] + n[2] * g.proj[d][4];
    float gy = n[0] * g.proj[d][1] + n[1] * g.proj[d][3] + n[2] * g.proj[d][5];
    Foot f = footprint(gx, gy, g.rw[d], g.rh[d]);
    float gv[4];
    bool any = false;
    for (int c = 0; c < 4; ++c) {
      gv[c] = gm[c] + (d_feat_p ? d_feat_p[row * (3 * g.C) + d * g.C + ch + c] : 0.f);
      any = any || gv[c] != 0.f;
    }
    if (!any) continue;   // rows that received no gradient (weight 0 samples) cost no atomics
    float* pl = d_plane[d];
    int64_t rw = g.rw[d];
    int64_t o00 = ((int64_t)f.y0 * rw + f.x0) * g.C + ch, o01 = ((int64_t)f.y0 * rw + f.x1) * g.C + ch;
    int64_t o10 = ((int64_t)f.y1 * rw + f.x0) * g.C + ch, o11 = ((int64_t)f.y1 * rw + f.x1) * g.C + ch;
    if (f.w00 != 0.f) accumulate4(pl + o00, f.w00, gv);
    if (f.w01 != 0.f) accumulate4(pl + o01, f.w01, gv);
    if (f.w10 != 0.f) accumulate4(pl + o10, f.w10, gv);
    if (f.w11 != 0.f) accumulate4(pl + o11, f.w11, gv);
  }
}

// backward of the view-direction gather (nerf_helpers.py:492-496 cart2az_el, models.py:312-326 project_viewdir):
// channels [ch, ch+4) of one ray.  d_vplane: channels-last fp32 [rh][rw][C].
NVSR_HD void viewdir_gather_bwd_ray(const float* viewdirs, int64_t ray, int ch, int rh, int rw, int C, float az_lo,
                                    float az_rng, float el_lo, float el_rng, const float* d_vfeat, float* d_vplane) {
  float dx = viewdirs[ray * 3 + 0], dy = viewdirs[ray * 3 + 1], dz = viewdirs[ray * 3 + 2];
  float el = atan2f(dz, sqr(add(mul(dx, dx), mul(dy, dy))));
  float az = atan2f(dy, dx);
  Foot f = footprint(box_normalize(az, az_lo, az_rng), box_normalize(el, el_lo, el_rng), rw, rh);
  int64_t o00 = ((int64_t)f.y0 * rw + f.x0) * C + ch, o01 = ((int64_t)f.y0 * rw + f.x1) * C + ch;
  int64_t o10 = ((int64_t)f.y1 * rw + f.x0) * C + ch, o11 = ((int64_t)f.y1 * rw + f.x1) * C + ch;
  float gv[4];
  bool any = false;
  for (int c = 0; c < 4; ++c) {
    gv[c] = d_vfeat[ray * C + ch + c];
    any = any || gv[c] != 0.f;
  }
  if (!any) return;
  if (f.w00 != 0.f) accumulate4(d_vplane + o00, f.w00, gv);
  if (f.w01 != 0.f) accumulate4(d_vplane + o01, f.w01, gv);
  if (f.w10 != 0.f) accumulate4(d_vplane + o10, f.w10, gv);
  if (f.w11 != 0.f) accumulate4(d_vplane + o11, f.w11, gv);
}

// ------------------------------------------------------------------------------------------------------------
// backward of volume_render_radiance_field (volume_rendering_utils.py:15-51) for one ray.
//   c_i = sigmoid(raw_i[0:3]);  s_i = relu(raw_i[3] + nz_i);  a_i = 1 - exp(-s_i * dist_i);
//   q_i = 1 - a_i + 1e-10;  T_i = prod_{j<i} q_j;  w_i = a_i * T_i
//   rgb = sum_i w_i c_i (+ 1 - acc if white);  acc = sum_i w_i;  depth = sum_i w_i t_i
// With the upstream gradients g_rgb[3], g_acc, g_depth and g_w[i] (any of the last three may be absent):
//   dL/dw_i  = g_rgb . c_i + g_acc + g_depth * t_i + g_w[i]  - (white ? sum(g_rgb) : 0)
//   dL/draw_i[0:3] = w_i * g_rgb * c_i * (1 - c_i)
//   dL/da_i  = dL/dw_i * T_i - (sum_{k>i} dL/dw_k * w_k) / q_i          (cumprod backward, q_i >= 1e-10 > 0)
//   dL/draw_i[3] = (raw_i[3] + nz_i > 0) ? dL/da_i * dist_i * exp(-s_i * dist_i) : 0
// raw / d_raw: interleaved [S,4] of this ray (the reference's radiance_field[N,S,4]); z: [S] depths, or [S+1]
// interval edges when mip != 0 (no 1e10 tail, t_i = interval midpoints, volume_rendering_utils.py:20-27,41-42).
// Pass 1 walks front to back and parks T_i / w_i in d_raw (this thread owns those 16 bytes); pass 2 walks back to
// front with the running suffix sum, so nothing is divided by a product of small factors.
NVSR_HD void composite_bwd_ray(const float* raw, const float* z, float rd_norm, const float* noise, int S, int white,
                               int mip, const float* g_rgb, const float* g_acc, const float* g_depth,
                               const float* g_w, float* d_raw) {
  float T = 1.f;
  for (int i = 0; i < S; ++i) {
    float dist = (i + 1 < S || mip) ? sub(z[i + 1], z[i]) : 1e10f;
    dist = mul(dist, rd_norm);
    float pre = raw[i * 4 + 3] + (noise ? noise[i] : 0.f);
    float s = pre > 0.f ? pre : 0.f;
    float a = sub(1.f, expf(-mul(s, dist)));
    d_raw[i * 4 + 0] = T;
    d_raw[i * 4 + 1] = mul(a, T);
    T = mul(T, add(sub(1.f, a), 1e-10f));
  }
  float gsum = white ? g_rgb[0] + g_rgb[1] + g_rgb[2] : 0.f;
  float ga = g_acc ? g_acc[0] : 0.f, gd = g_depth ? g_depth[0] : 0.f;
  float suffix = 0.f;   // sum_{k>i} dL/dw_k * w_k
  for (int i = S - 1; i >= 0; --i) {
    float Ti = d_raw[i * 4 + 0], w = d_raw[i * 4 + 1];
    float dist = (i + 1 < S || mip) ? sub(z[i + 1], z[i]) : 1e10f;
    dist = mul(dist, rd_norm);
    float pre = raw[i * 4 + 3] + (noise ? noise[i] : 0.f);
    float s = pre > 0.f ? pre : 0.f;
    float e = expf(-mul(s, dist));
    float a = sub(1.f, e);
    float q = add(sub(1.f, a), 1e-10f);
    float t = mip ? 0.5f * (z[i] + z[i + 1]) : z[i];
    float dw = ga + gd * t + (g_w ? g_w[i] : 0.f) - gsum;
    for (int c = 0; c < 3; ++c) {
      float col = 1.f / (1.f + expf(-raw[i * 4 + c]));
      dw += g_rgb[c] * col;
      d_raw[i * 4 + c] = w * g_rgb[c] * col * (1.f - col);
    }
    float da = dw * Ti - suffix / q;
    d_raw[i * 4 + 3] = pre > 0.f ? da * dist * e : 0.f;
    suffix += dw * w;
  }
}

}  // namespace bwd
}  // namespace nvsr
