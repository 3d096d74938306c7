// Frame sink: the step right after the render path.  The reference moves the fp32 frame to the
// host and converts there (train_nerf.py:270,273: 255*clamp(im,0,1) -> uint8 -> imageio); here the conversion happens
// on the device, so the device->host copy carries 1 byte per channel instead of 4.
#include "common.cuh"
#include "frame_bodies.h"

namespace nvsr {

__global__ void __launch_bounds__(256) frame_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out,
                                                          int64_t n_elems) {
  int64_t quads = (n_elems + 3) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (int64_t)gridDim.x * blockDim.x)
    frame::to_u8_quad(in, out, i, n_elems);
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_frame_to_u8(const float* rgb, int64_t n_elems, uint8_t* out, void* stream) {
  NVSR_CHECK_ARG(n_elems >= 0 && (n_elems == 0 || (rgb && out)));
  if (n_elems == 0) return NVSR_OK;
  if (reinterpret_cast<uintptr_t>(out) & 3u) return NVSR_ERR_ALIGNMENT;
  int64_t quads = (n_elems + 3) / 4;
  int64_t blocks = ceil_div64(quads, 256);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;   // grid-stride beyond 8 CTAs per SM
  frame_to_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, out, n_elems);
  NVSR_RETURN_LAST_ERROR();
}
