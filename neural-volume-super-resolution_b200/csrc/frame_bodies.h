// Per-element body of the frame-sink kernel, host/device-neutral like backward_bodies.h:
// csrc/frames.cu instantiates it on the device, tests/hostcheck compiles it with g++ for the CPU check.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NVSR_HD __host__ __device__ __forceinline__
#else
#define NVSR_HD inline
#endif

namespace nvsr {
namespace frame {

// write_image (train_nerf.py:270,273): np.array(255*torch.clamp(im,0,1).cpu()).astype(np.uint8) — fp32 product,
// truncating cast.  NaN: clamp propagates it and the x86 cast yields 0; fmaxf(NaN, 0) = 0 gives the same byte.
NVSR_HD uint8_t to_u8(float x) {
  float c = fminf(fmaxf(x, 0.f), 1.f);
  return (uint8_t)(255.f * c);
}

// elements [4*i, 4*i+4) of a flat fp32 array -> 4 bytes (one 32-bit store when all four exist)
NVSR_HD void to_u8_quad(const float* in, uint8_t* out, int64_t i, int64_t n_elems) {
  int64_t e = i * 4;
  if (e + 4 <= n_elems) {
    uint32_t v = (uint32_t)to_u8(in[e]) | ((uint32_t)to_u8(in[e + 1]) << 8) | ((uint32_t)to_u8(in[e + 2]) << 16) |
                 ((uint32_t)to_u8(in[e + 3]) << 24);
    *reinterpret_cast<uint32_t*>(out + e) = v;   // out is 4-byte aligned (checked by the entry point)
  } else {
    for (; e < n_elems; ++e) out[e] = to_u8(in[e]);
  }
}

}  // namespace frame
}  // namespace nvsr
