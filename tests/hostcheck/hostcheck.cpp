// Host build of the backward kernels' per-element bodies (csrc/backward_bodies.h): the SAME source the CUDA
// kernels of csrc/backward.cu instantiate, compiled with g++ (-ffp-contract=off) and looped over the kernels' own
// index decomposition.  TEST INFRASTRUCTURE (tests/test_backward_bodies.py); never loaded by the product path.
#include <math.h>
#include <stdint.h>

#include "backward_bodies.h"
#include "frame_bodies.h"

using namespace nvsr::bwd;

extern "C" {

// mirrors gather_bwd_kernel: idx -> (row, chunk), ray = row / S
void hc_gather_bwd(const int* rh, const int* rw, int C, const float* lo, const float* rng, const float* proj /*[3][6]*/,
                   const float* ro, const float* rd, const float* z, int64_t n_rays, int S, const float* d_feat_p,
                   const float* d_feat_m, float* dp0, float* dp1, float* dp2) {
  PlaneGeom g;
  for (int d = 0; d < 3; ++d) {
    g.rh[d] = rh[d], g.rw[d] = rw[d], g.lo[d] = lo[d], g.rng[d] = rng[d];
    for (int k = 0; k < 6; ++k) g.proj[d][k] = proj[d * 6 + k];
  }
  g.C = C;
  float* dpl[3] = {dp0, dp1, dp2};
  const int chunks = C / 4;
  for (int64_t idx = 0; idx < n_rays * S * chunks; ++idx) {
    int64_t row = idx / chunks;
    int ch = (int)(idx % chunks) * 4;
    gather_bwd_row(g, ro, rd, z[row], row / S, row, ch, d_feat_p, d_feat_m, dpl);
  }
}

void hc_viewdir_gather_bwd(const float* viewdirs, int64_t n, int rh, int rw, int C, float az_lo, float az_rng, float el_lo,
                           float el_rng, const float* d_vfeat, float* d_vplane) {
  const int chunks = C / 4;
  for (int64_t idx = 0; idx < n * chunks; ++idx)
    viewdir_gather_bwd_ray(viewdirs, idx / chunks, (int)(idx % chunks) * 4, rh, rw, C, az_lo, az_rng, el_lo, el_rng, d_vfeat,
                           d_vplane);
}

// mirrors composite_bwd_kernel: one "thread" per ray
void hc_composite_bwd(const float* raw, const float* z, const float* rd, const float* noise, int64_t n, int S, int white,
                      int mip, const float* g_rgb, const float* g_acc, const float* g_depth, const float* g_w, float* d_raw) {
  const int Z = S + (mip ? 1 : 0);
  for (int64_t ray = 0; ray < n; ++ray) {
    float dx = rd[ray * 3 + 0], dy = rd[ray * 3 + 1], dz = rd[ray * 3 + 2];
    float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    composite_bwd_ray(raw + ray * S * 4, z + ray * Z, nrm, noise ? noise + ray * S : nullptr, S, white, mip, g_rgb + ray * 3,
                      g_acc ? g_acc + ray : nullptr, g_depth ? g_depth + ray : nullptr, g_w ? g_w + ray * S : nullptr,
                      d_raw + ray * S * 4);
  }
}

// mirrors frame_to_u8_kernel: one "thread" per quad
void hc_frame_to_u8(const float* in, int64_t n_elems, uint8_t* out) {
  for (int64_t i = 0; i < (n_elems + 3) / 4; ++i) nvsr::frame::to_u8_quad(in, out, i, n_elems);
}
}
