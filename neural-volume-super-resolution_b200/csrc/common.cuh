// Shared device/host helpers for libnvsr_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nvsr.h"

#define NVSR_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return NVSR_ERR_INVALID_ARG; \
  } while (0)

#define NVSR_RETURN_LAST_ERROR()            \
  do {                                      \
    cudaError_t e__ = cudaGetLastError();   \
    return e__ == cudaSuccess ? NVSR_OK : (int32_t)e__; \
  } while (0)

namespace nvsr {

constexpr int kTileRows = NVSR_TILE_ROWS;  // rows per decoder tile (UMMA M)
constexpr int kNumSMs = 148;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- BLOCKED row order (nvsr.h): a tile = 8 consecutive rays x 16 consecutive samples ----------
constexpr int kBlkRays = NVSR_BLK_RAYS, kBlkSamples = NVSR_BLK_SAMPLES;
static_assert(kBlkRays * kBlkSamples == NVSR_TILE_ROWS, "a block is one decoder tile");
__host__ __device__ inline int tiles_per_block(int S) { return (S + kBlkSamples - 1) / kBlkSamples; }
__host__ __device__ inline int64_t rows_padded(int64_t n_rays, int S, int order) {
  if (order == NVSR_ROWS_BLOCKED) return ceil_div64(n_rays, kBlkRays) * tiles_per_block(S) * kTileRows;
  return n_rays * S;
}
// (tile, row-in-tile) -> (ray, sample)
__device__ __forceinline__ void blocked_decode(int64_t tile, int r, int TS, int64_t* ray, int* s) {
  int64_t b = tile / TS;
  int sb = (int)(tile - b * TS);
  *s = sb * kBlkSamples + (r >> 3);
  *ray = b * kBlkRays + (r & 7);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- bf16 packing -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  // cvt.rn.bf16x2.f32 d, a, b : d.hi = a, d.lo = b
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 16-bit operand type of the tensor-core path: bf16 (8-bit significand, fp32 range) or fp16 (11-bit
// significand, max 65504).  cvt.satfinite: an overflow saturates to the largest finite value instead of
// inf, in the conversion instruction itself (F2FP.SATFINITE.*.PACK_AB, no separate clamps).
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
  uint32_t r;
  if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool F16>
__device__ __forceinline__ float2 unpack16x2(uint32_t v) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  else return make_float2(bf16lo_to_f32(v), bf16hi_to_f32(v));
}
inline bool is_16bit(int dtype) { return dtype == NVSR_BF16 || dtype == NVSR_F16; }

// ---- shared-memory address / mbarrier ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// try_wait suspends the warp in hardware for a bounded, implementation-defined time while the phase is
// incomplete, so a polling warp issues about one instruction per ~200 cycles.  (An explicit suspend-time
// hint compiles to NANOSLEEP.SYNCS and measured ~1% slower on the decoder: -DNVSR_WAIT_HINT_NS=<ns>.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
#ifdef NVSR_WAIT_HINT_NS
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
#ifdef NVSR_WAIT_HINT_NS
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)NVSR_WAIT_HINT_NS)
#else
      : "r"(smem_u32(bar)), "r"(parity)
#endif
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Wait on the barrier phase.  A protocol bug must surface as a launch failure, never as a hung GPU:
// after ~4 s without progress the kernel traps (cudaErrorLaunchFailure on the host).
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;   // fast path: already complete / completes within the hint
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

// ---- bulk async copies (TMA engine, no tensor map) --------------------------------------------
// global -> shared, completion counted on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / UMMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace nvsr
