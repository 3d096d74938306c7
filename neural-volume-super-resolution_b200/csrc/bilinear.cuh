// Bilinear corner setup shared by the plane-gather kernels.
#pragma once
#include "common.cuh"

namespace nvsr {

// grid_sample(bilinear, align_corners=True, padding_mode='border') coordinate -> corner data.
// ATen: ix = ((g+1)/2)*(size-1); clip to [0,size-1]; nw=floor.  Separately rounded ops.
struct Bilin {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;  // nw, ne, sw, se
};
__device__ __forceinline__ Bilin bilinear_setup(float gx, float gy, int Wd, int Hd) {
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(Wd - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(Hd - 1));
  ix = fminf(fmaxf(ix, 0.f), (float)(Wd - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hd - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  Bilin b;
  b.w00 = __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(fy1, iy));
  b.w01 = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(fy1, iy));
  b.w10 = __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(iy, fy0));
  b.w11 = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(iy, fy0));
  b.x0 = (int)fx0, b.y0 = (int)fy0;
  // out-of-range corners contribute 0 in ATen; their weight is exactly 0 there, so clamping the
  // index is equivalent and keeps the load in bounds.
  b.x1 = min(b.x0 + 1, Wd - 1);
  b.y1 = min(b.y0 + 1, Hd - 1);
  if (b.x0 + 1 > Wd - 1) b.w01 = 0.f, b.w11 = 0.f;
  if (b.y0 + 1 > Hd - 1) b.w10 = 0.f, b.w11 = 0.f;
  return b;
}

// normalize_coords (models.py:264-265): 2*(c-lo)/rng - 1, each op rounded separately
__device__ __forceinline__ float box_normalize(float c, float lo, float rng) {
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(c, lo)), rng), 1.f);
}

}  // namespace nvsr
