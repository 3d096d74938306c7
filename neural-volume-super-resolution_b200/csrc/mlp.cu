// C-ABI entry of the decoder chain: dispatch on the precision contract.
#include "common.cuh"

namespace nvsr {
int32_t launch_mlp_f32(const nvsr_mlp_t* m, cudaStream_t st);
int32_t launch_mlp_tc(const nvsr_mlp_t* m, cudaStream_t st, void* const* act_out);
}  // namespace nvsr

extern "C" int32_t nvsr_mlp_chain(const nvsr_mlp_t* m, void* stream) {
  NVSR_CHECK_ARG(m && m->n_layers > 0 && m->n_layers <= NVSR_MAX_LAYERS);
  NVSR_CHECK_ARG(m->in && m->raw && m->rows >= 0 && m->raw_stride >= m->rows);
  if (m->rows == 0) return NVSR_OK;
  if (m->precision == NVSR_F32) return nvsr::launch_mlp_f32(m, (cudaStream_t)stream);
  if (m->precision == NVSR_BF16 || m->precision == NVSR_F16) return nvsr::launch_mlp_tc(m, (cudaStream_t)stream, nullptr);
  return NVSR_ERR_UNSUPPORTED;
}
