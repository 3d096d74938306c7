// a6: decoder MLP chain on the 5th-gen tensor cores (tcgen05.mma, accumulators AND hidden
// activations in TMEM).
//
// Persistent kernel, one CTA per SM, cta_group::1, UMMA M=128 (one 128-row tile), N = layer width.
// Two tiles ("slots") are in flight, each owned by 8 warps, so the slots advance independently and
// the tensor core runs one slot's layer while the other slot is in its epilogue.  There is no
// producer warp and no MMA-issuer warp:
//   epilogue (all 16 warps, 8 per slot): warp w owns TMEM lanes 32*(w%4)..+31 (32 rows) and columns
//       [64h, 64h+64): tcgen05.ld the accumulator, ReLU + saturate + 16-bit pack in ONE F2FP per
//       pair, tcgen05.st the packed row into the slot's A region.  No bias add is ever executed:
//       in the two tri-plane chains a layer's bias is ONE extra K = 16 MMA step (constant "ones"
//       pattern in TMEM x a 4 KB image of the biases split into three 16-bit terms, see kPatCol);
//       only the PER-RAY bias of the rgb chain's first layer (and every bias of the generic chain)
//       is pre-stored with tcgen05.st into the accumulator columns just read, and the MMAs
//       accumulate onto it.  Output heads (128 -> 1 / 3) are fp32 dot products on the CUDA cores
//       over the UNROUNDED last activations.
//   MMA issue: the LAST of a slot's 8 warps to finish a layer (shared-memory arrival counter) issues
//       the next layer's tcgen05.mma itself from one elected lane — layer 0: A = feature tile in smem
//       (SS form); hidden layers: A = previous activations in TMEM (TS form); B = resident weights
//       in smem; D = fp32 accumulator in TMEM — and commits them to the slot's mbarrier.
//   loads: weights are bulk-copied (TMA engine) into shared memory once; per tile ONE thread of the
//       slot issues the bulk copies of the next feature tile image (and of its 8 per-ray bias rows)
//       the moment layer 0's accumulator is complete — exactly when the ring buffer is free.
// Hidden activations never leave the SM (never even touch shared memory); weights are read from
// HBM/L2 once per CTA; a ring buffer is released as soon as layer 0's MMAs have been committed.
//
// TMEM map (512 columns allocated): slot s -> D_s = [192 s, 192 s + 128) fp32 accumulator,
// A_s = [192 s + 128, 192 s + 192): 128 rows x 128 16-bit activations, two per 32-bit column
// (row = lane, K pair k/2 = column: the TS-form A layout); [384, 416): the four bias "ones" patterns.
//
// smem operand layout (layer-0 A and every B): K-major, no-swizzle canonical UMMA layout with
// 8x16-byte core matrices, stored [K/8][rows][8 x 16 bit]: LBO (K-chunk stride) = rows*16 B,
// SBO (8-row group stride) = 128 B.  The feature tile image written by the gather kernel and the
// weight image written by nvsr_pack_weight16 are exactly this, so both arrive with plain bulk copies.
#include "common.cuh"

namespace nvsr {

constexpr int kTcEpiWarpsPerSlot = 8;
constexpr int kTcThreads = 32 * 2 * kTcEpiWarpsPerSlot;  // 16 warps = 4 per scheduler -> 128 registers each
constexpr int kTcMaxHeadRows = 4;   // total head outputs of a chain (r,g,b,sigma)
constexpr int kRbRowsMax = 8;       // staged per-ray bias rows per tile
constexpr int kRbPitch = 132;       // floats per staged row (528 B: conflict-free for 8 rows)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSlotCols = 192;  // D (128 fp32 columns) + A (64 columns of packed 16-bit pairs)
constexpr uint32_t kSlotAOff = 128;
// Fixed chains: a layer's bias enters as ONE extra K = 16 MMA step instead of being pre-stored into the accumulator
// (tcgen05.st of 128 fp32 columns per row and layer, the largest single item of the epilogue): A = a constant
// "ones" pattern in TMEM (shared by both slots), B = a 4 KB image holding every layer's bias split into three 16-bit
// terms (hi + mid + lo: exact to 2^-30 relative in fp16, 2^-24 in bf16 — below the fp32 accumulator's own rounding).
// Layer l uses K rows 4l .. 4l+2 of the image; pattern l has ones in exactly those K positions.
constexpr uint32_t kPatCol = 2 * kSlotCols;   // TMEM columns [384, 416): 4 patterns x 8 columns (16 K values each)
constexpr uint32_t kBiasImgBytes = 2u * 128u * 16u;   // [2 K-chunks][128 n][8] 16-bit
// rgb chain (3 outputs): the head runs on the tensor core too — one N = 16 MMA group per tile over the 16-bit last
// activations (A region) against a [128 K][16 N] image of the head weights split hi + lo (rows 0-2 / 3-5), into 16
// spare TMEM columns per slot; it is issued together with the slot's next layer-0 MMAs and read back (8 columns per row)
// under them.  This takes the head dot products, the half-combining through shared memory and its named barrier off the
// slot's critical path (~1 500 of the head epilogue's ~1 850 cycles, measured with -DNVSR_TC_TIMING): rgb chain 1 065 ->
// 1 216 TFLOP/s alone.  The head weights stay exact to 2^-22; the activations enter rounded to 16 bit: colour logits
// 3.9e-5 instead of 3.0e-5 off in fp16.  The DENSITY head stays an fp32 dot product over the unrounded activations: tried
// on the tensor core as well — sigma error +50-70 % (0.073 -> 0.126 fp16, 0.55 -> 1.05 bf16: sigma's head weights are
// large) for 2 % of the density chain; rejected.  The generic chain (mip decoder) keeps the fp32 head too.
constexpr uint32_t kHeadCol = kPatCol + 32u;          // TMEM columns [416, 432) slot 0, [432, 448) slot 1 (fixed chains;
                                                      // the generic chains' eight 2-term patterns fill [384, 448))
constexpr uint32_t kHeadImgBytes = 16u * 16u * 16u;   // [16 K-chunks][16 n][8] 16-bit

struct TcLayer {
  const void* w;          // global 16-bit image
  const float* bias;      // global
  const float* row_bias;  // global per-ray or null
  const float* head_w;
  const float* head_b;
  int k, n, relu, head_n, head_ch, head_row;  // head_row: first row of this head in the smem head table
  uint32_t w_off;         // smem byte offset of the weight image
};

struct TcArgs {
  TcLayer layer[NVSR_MAX_LAYERS];
  int n_layers;
  const uint8_t* in;      // tile images, in_bytes each
  uint32_t in_bytes;
  int64_t rows, n_tiles;
  int samples_per_ray, row_order, tiles_per_blk;
  int64_t n_rays;
  float* raw;
  int64_t raw_stride;
  const int32_t* row_ids;    // sparse mode: list entry i (input row i) is BLOCKED row row_ids[i] of the frame chunk
  const int32_t* row_count;  // sparse mode: number of list entries (device memory)
  int rb_layer;           // layer with a per-ray bias (-1: none)
  int rb_staged;          // 1: the producer stages the tile's bias rows in smem (BLOCKED order)
  // smem carve-up (byte offsets from the 1024-aligned base)
  uint32_t in_off[2], rb_off[2], bias_off, headw_off, hpart_off, bar_off, w_bytes_total, bimg_off, himg_off;
  // training forward (TRAIN kernels): act[l] = tile images [tile][16][128][8] of layer l's post-ReLU 16-bit output —
  // exactly the values the next layer's MMA reads (the last layer's: the head input rounded to 16 bit)
  uint8_t* act[4];
};

#ifdef NVSR_TC_TIMING
// Debug build only (scripts/ab_build.sh timing -DNVSR_TC_TIMING): per-warp cycle sums of the phases of a layer step in
// CTA 0 — [warp][0] wait for the accumulator, [1] epilogue body, [2] arrive (+ MMA issue on the last warp),
// [3] rest of the step (tile loads, head write-out), [4] layer steps counted, [5] steps in which this warp was the issuer,
// [6] cycles of those issues.  Read back with nvsr_debug_tc_timing().
__device__ unsigned long long g_tc_timing[16][8][8];   // [warp][layer & 7][phase]
#define TC_T(var) const long long var = clock64()
#else
#define TC_T(var)
#endif

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// kind::f16 instruction descriptor: (bf16|f16) x (bf16|f16) -> fp32, both operands K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_16(int n, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;  // F16F32Format: 0 = F16, 1 = BF16
  return (1u << 4) /*D=f32*/ | (fmt << 7) /*A*/ | (fmt << 10) /*B*/ | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (base_lane + t)
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (ReLU +) saturate-to-finite + round + pack two fp32 into one 16-bit pair: a single F2FP
template <bool F16, bool RELU>
__device__ __forceinline__ uint32_t pack_act(float lo, float hi) {
  uint32_t r;
  if constexpr (F16 && RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else if constexpr (RELU) asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// barrier block layout (uint64_t each, [2] = one per slot)
enum { BAR_W = 0, BAR_IN_FULL = 1, BAR_RB_FULL = 3, BAR_ACC_FULL = 5, BAR_DONE = 7, BAR_COUNT = 9 };

__device__ __forceinline__ void load_bias32(const float* src, uint32_t (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b4 = *reinterpret_cast<const float4*>(src + 4 * j);
    v[4 * j + 0] = __float_as_uint(b4.x), v[4 * j + 1] = __float_as_uint(b4.y);
    v[4 * j + 2] = __float_as_uint(b4.z), v[4 * j + 3] = __float_as_uint(b4.w);
  }
}

// One 32-column pass of a layer epilogue for one thread (= one row of the tile):
//   read    : tcgen05.ld the accumulator columns
//   pack    : (ReLU,) saturate, round to 16 bit, tcgen05.st as the next layer's A operand
//   head_n  : accumulate head_n fp32 dot products of the (ReLU'd) unrounded activations
//   bias    : tcgen05.st the next accumulation's bias into the same columns (1: smem/generic ptr, 2: global)
template <bool F16>
__device__ __forceinline__ void epi_pass(uint32_t d_addr, uint32_t a_addr, bool read, bool pack, bool relu, int head_n,
                                         const float* hw, float (&hacc)[4], int bias_mode, const float* bsrc) {
  uint32_t v[32];
  if (read) {
    tmem_ld32(d_addr, v);
    tmem_ld_wait();
    if (pack) {
      uint32_t pk[16];
      if (relu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
      }
      tmem_st16(a_addr, pk);
    }
    if (head_n > 0) {
      if (relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (h < head_n) {
          const float4* hw4 = reinterpret_cast<const float4*>(hw + h * 128);
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 w4 = hw4[j];
            acc0 = fmaf(__uint_as_float(v[4 * j + 0]), w4.x, acc0);
            acc1 = fmaf(__uint_as_float(v[4 * j + 1]), w4.y, acc1);
            acc0 = fmaf(__uint_as_float(v[4 * j + 2]), w4.z, acc0);
            acc1 = fmaf(__uint_as_float(v[4 * j + 3]), w4.w, acc1);
          }
          hacc[h] += acc0 + acc1;
        }
      }
    }
  }
  if (bias_mode == 1) {
    load_bias32(bsrc, v);
    tmem_st32(d_addr, v);
  } else if (bias_mode == 2) {
    const float4* g = reinterpret_cast<const float4*>(bsrc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b4 = __ldg(g + j);
      v[4 * j + 0] = __float_as_uint(b4.x), v[4 * j + 1] = __float_as_uint(b4.y);
      v[4 * j + 2] = __float_as_uint(b4.z), v[4 * j + 3] = __float_as_uint(b4.w);
    }
    tmem_st32(d_addr, v);
  }
}

// packed fp32x2 helpers for the head dot products (FFMA2: two MACs per instruction)
__device__ __forceinline__ unsigned long long tc_pack2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ unsigned long long tc_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float tc_sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// HN fp32 head dot products over 32 accumulator columns of one row; the ReLU is applied in place
// (v is dead afterwards), column pairs are the FFMA2 operands as they sit in the registers
template <int HN>
__device__ __forceinline__ void head_dot32(uint32_t (&v)[32], const float* hw, float (&hacc)[4]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
#pragma unroll
  for (int h = 0; h < HN; ++h) {
    const float4* hw4 = reinterpret_cast<const float4*>(hw + h * 128);
    unsigned long long acc0 = 0ull, acc1 = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w4 = hw4[j];
      acc0 = tc_fma2(tc_pack2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), tc_pack2(w4.x, w4.y), acc0);
      acc1 = tc_fma2(tc_pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), tc_pack2(w4.z, w4.w), acc1);
    }
    hacc[h] += tc_sum2(acc0) + tc_sum2(acc1);
  }
}

// Epilogue of one layer of a FIXED chain for one thread (= one row, this warp's 64 columns): both
// 32-column accumulator loads are issued back to back and waited for once, so the second load's latency
// hides behind the first group's pack/store.  ReLU always; HN > 0: last layer (heads, no A operand).
// act (training forward only, else nullptr): this thread's 16-byte slot of the warp's first 8-column chunk in the
// layer's activation tile image; chunk c of the warp's 64 columns lies c * 128 slots further
__device__ __forceinline__ void store_act16(uint4* act, int chunk0, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) act[(chunk0 + c) * 128] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

template <bool F16, int HN>
__device__ __forceinline__ void epi_fixed(uint32_t d_addr, uint32_t a_addr, const float* hw, float (&hacc)[4],
                                          bool bias, const float* bsrc, uint4* act = nullptr) {
  if constexpr (HN == 0) {
    uint32_t v0[32], v1[32];
    tmem_ld32(d_addr, v0);
    tmem_ld32(d_addr + 32u, v1);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
    tmem_st16(a_addr, pk);
    if (act) store_act16(act, 0, pk);
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
    tmem_st16(a_addr + 16u, pk);
    if (act) store_act16(act, 4, pk);
    if (bias) {
      load_bias32(bsrc, v0);
      tmem_st32(d_addr, v0);
      load_bias32(bsrc + 32, v1);
      tmem_st32(d_addr + 32u, v1);
    }
  } else {
    // head layer: both groups are loaded up front as well (a TMEM load costs a few hundred cycles while the
    // other slot's MMAs run); each group's registers are reused for its bias once its dot products are done
    uint32_t v0[32], v1[32];
    tmem_ld32(d_addr, v0);
    tmem_ld32(d_addr + 32u, v1);
    tmem_ld_wait();
    head_dot32<HN>(v0, hw, hacc);
    if (act) {   // v0 holds the ReLU'd activations now
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
      store_act16(act, 0, pk);
    }
    if (bias) {
      load_bias32(bsrc, v0);
      tmem_st32(d_addr, v0);
    }
    head_dot32<HN>(v1, hw + 32, hacc);
    if (act) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
      store_act16(act, 4, pk);
    }
    if (bias) {
      load_bias32(bsrc + 32, v1);
      tmem_st32(d_addr + 32u, v1);
    }
  }
}

// Epilogue of one layer of a GENERIC chain for one thread (= one row, this warp's 64 columns; layer widths are multiples
// of 64, so a warp owns either all 64 or none): the fixed chains' structure — both 32-column accumulator loads issued back
// to back and waited for once — with the layer's shape as run-time (warp-uniform) flags.  Measured on the mip decoder
// (-DNVSR_TC_TIMING): the 32-columns-at-a-time version cost ~1 040 cycles per hidden layer against ~220 in the fixed chains.
template <bool F16>
__device__ __forceinline__ void epi_pass64(uint32_t d_addr, uint32_t a_addr, bool read, bool pack, bool relu, int head_n,
                                           const float* hw, float (&hacc)[4], int bias_mode, const float* bsrc) {
  uint32_t v0[32], v1[32];
  if (read) {
    tmem_ld32(d_addr, v0);
    tmem_ld32(d_addr + 32u, v1);
    tmem_ld_wait();
    if (pack) {
      uint32_t pk[16];
      if (relu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
        tmem_st16(a_addr, pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, true>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        tmem_st16(a_addr + 16u, pk);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
        tmem_st16(a_addr, pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_act<F16, false>(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        tmem_st16(a_addr + 16u, pk);
      }
    }
    if (head_n > 0) {
      if (relu) {   // head_dot32 applies the ReLU in place
        if (head_n == 1) head_dot32<1>(v0, hw, hacc), head_dot32<1>(v1, hw + 32, hacc);
        else if (head_n == 3) head_dot32<3>(v0, hw, hacc), head_dot32<3>(v1, hw + 32, hacc);
        else head_dot32<4>(v0, hw, hacc), head_dot32<4>(v1, hw + 32, hacc);   // rows >= head_n of the table are zero
      } else {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (h < head_n) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc = fmaf(__uint_as_float(v0[j]), hw[h * 128 + j], acc);
              acc = fmaf(__uint_as_float(v1[j]), hw[h * 128 + 32 + j], acc);
            }
            hacc[h] += acc;
          }
        }
      }
    }
  }
  if (bias_mode == 1) {
    load_bias32(bsrc, v0);
    tmem_st32(d_addr, v0);
    load_bias32(bsrc + 32, v1);
    tmem_st32(d_addr + 32u, v1);
  } else if (bias_mode == 2) {
    const float4* g = reinterpret_cast<const float4*>(bsrc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(g + j), c4 = __ldg(g + 8 + j);
      v0[4 * j + 0] = __float_as_uint(b4.x), v0[4 * j + 1] = __float_as_uint(b4.y);
      v0[4 * j + 2] = __float_as_uint(b4.z), v0[4 * j + 3] = __float_as_uint(b4.w);
      v1[4 * j + 0] = __float_as_uint(c4.x), v1[4 * j + 1] = __float_as_uint(c4.y);
      v1[4 * j + 2] = __float_as_uint(c4.z), v1[4 * j + 3] = __float_as_uint(c4.w);
    }
    tmem_st32(d_addr, v0);
    tmem_st32(d_addr + 32u, v1);
  }
}

// LC > 0: "uniform" chain known at compile time — LC layers, all 128 wide with ReLU, one head of HN
// rows on the last layer, per-ray bias on layer 0 iff RB0 (staged rows, BLOCKED order).  Both decoders
// of the tri-plane model are of this shape (LC = 4).  LC == 0: generic chain described at run time.
template <bool F16, int LC, int HN, int RB0, bool TRAIN = false>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_chain_tc_kernel(const __grid_constant__ TcArgs a) {
  constexpr bool kFixed = LC > 0;
  constexpr bool kHeadTC = kFixed && HN == 3;   // rgb chain: head on the tensor core (see kHeadCol)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* sbias = reinterpret_cast<float*>(smem + a.bias_off);    // [n_layers][128]
  float* sheadw = reinterpret_cast<float*>(smem + a.headw_off);  // [kTcMaxHeadRows][128]
  float* shpart = reinterpret_cast<float*>(smem + a.hpart_off);  // [2 slots][128 rows][4]: half 1 -> half 0

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = kFixed ? LC : (LC < 0 ? -LC : a.n_layers);  // LC < 0: generic chain of -LC layers, loop unrolled
  const int64_t G = gridDim.x;
  // RB0 (fixed chains): 0 = no per-ray bias, 1 = layer-0 bias rows staged per tile (dense BLOCKED order),
  // 2 = layer-0 bias read per row from global memory (sparse row list)
  const int rb_layer = kFixed ? (RB0 ? 0 : -1) : a.rb_layer;
  const bool rb_staged = kFixed ? (RB0 == 1) : (a.rb_staged != 0);
  int64_t rows = a.rows, n_tiles = a.n_tiles;
  if (a.row_count) {  // sparse mode: the list length is known on the device only
    rows = *a.row_count;
    n_tiles = (rows + kTileRows - 1) / kTileRows;
  }

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    mbar_init(&bars[BAR_W], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[BAR_IN_FULL + s], rb_staged ? 4 : kTcEpiWarpsPerSlot);   // arrivals per tile: see issue_tile_loads
      mbar_init(&bars[BAR_RB_FULL + s], 4);
      mbar_init(&bars[BAR_ACC_FULL + s], 1);
      mbar_init(&bars[BAR_DONE + s], kTcEpiWarpsPerSlot);
    }
    tmem_slot[1] = tmem_slot[2] = 0;  // per-slot arrival counters
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  // biases / head weights -> smem (tiny, read by every epilogue thread for every tile)
  if constexpr (!kHeadTC) {
    for (int i = threadIdx.x; i < kTcMaxHeadRows * 128; i += kTcThreads) sheadw[i] = 0.f;
    __syncthreads();
    for (int l = 0; l < L; ++l) {
      const TcLayer& ly = a.layer[l];
      if (ly.head_w) {
        for (int i = threadIdx.x; i < ly.head_n * ly.n; i += kTcThreads) {
          int h = i / ly.n, n = i - h * ly.n;
          sheadw[(ly.head_row + h) * 128 + n] = __ldg(ly.head_w + i);
        }
      }
    }
  } else {
    // head weight image [16 K-chunks][16 n][8]: n = h holds fp16(W_head[h][k]), n = 3 + h the remainder, n >= 6 zero
    uint32_t* himg32 = reinterpret_cast<uint32_t*>(smem + a.himg_off);
    for (int i = threadIdx.x; i < (int)(kHeadImgBytes / 4); i += kTcThreads) himg32[i] = 0u;
    __syncthreads();
    uint16_t* himg = reinterpret_cast<uint16_t*>(smem + a.himg_off);
    const TcLayer& lh = a.layer[LC - 1];
    for (int i = threadIdx.x; i < HN * 128; i += kTcThreads) {
      const int h = i >> 7, k = i & 127;
      const float w = __ldg(lh.head_w + h * 128 + k);
      const uint32_t hi = pack16x2<F16>(w, 0.f);
      const uint32_t lo = pack16x2<F16>(w - unpack16x2<F16>(hi).x, 0.f);
      himg[((k >> 3) * 16 + h) * 8 + (k & 7)] = (uint16_t)(hi & 0xffffu);
      himg[((k >> 3) * 16 + 3 + h) * 8 + (k & 7)] = (uint16_t)(lo & 0xffffu);
    }
    fence_proxy_async_smem();
  }
  if constexpr (kFixed) {
    // bias image: K row 4l + t of column n = term t of layer l's bias[n] (hi, mid, lo); K rows 4l + 3 stay zero
    uint16_t* bimg = reinterpret_cast<uint16_t*>(smem + a.bimg_off);
    for (int i = threadIdx.x; i < LC * 128; i += kTcThreads) {
      const int l = i >> 7, n = i & 127;
      float rem = a.layer[l].bias ? __ldg(a.layer[l].bias + n) : 0.f;
      uint16_t term[4];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint32_t pk = pack16x2<F16>(rem, 0.f);
        term[t] = (uint16_t)(pk & 0xffffu);
        rem -= unpack16x2<F16>(pk).x;
      }
      term[3] = 0;
      const int k0 = 4 * l;   // chunk k0 / 8, position k0 % 8 within the 16-byte group
      *reinterpret_cast<uint2*>(bimg + ((k0 >> 3) * 128 + n) * 8 + (k0 & 7)) =
          make_uint2((uint32_t)term[0] | ((uint32_t)term[1] << 16), (uint32_t)term[2] | ((uint32_t)term[3] << 16));
    }
    fence_proxy_async_smem();
  } else {
    // generic chains (up to 8 layers): K rows 2l, 2l + 1 of column n = hi, lo of layer l's bias[n] (exact to 2^-22 in
    // fp16, 2^-16 in bf16 — far below the operand rounding); a layer with a per-ray bias keeps the pre-stored accumulator
    uint32_t* bimg32 = reinterpret_cast<uint32_t*>(smem + a.bimg_off);
    for (int i = threadIdx.x; i < (int)(kBiasImgBytes / 4); i += kTcThreads) bimg32[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < L * 128; i += kTcThreads) {
      const int l = i >> 7, n = i & 127;
      const TcLayer& ly = a.layer[l];
      const float b = (n < ly.n && ly.bias) ? __ldg(ly.bias + n) : 0.f;
      const uint32_t hi = pack16x2<F16>(b, 0.f);
      const uint32_t lo = pack16x2<F16>(b - unpack16x2<F16>(hi).x, 0.f);
      const int k0 = 2 * l;
      bimg32[(((k0 >> 3) * 128 + n) * 8 + (k0 & 7)) >> 1] = (hi & 0xffffu) | (lo << 16);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (!kFixed) {
    // "ones" patterns of the generic chains: pattern l = columns [kPatCol + 8l, +8), ones at K = 2l, 2l + 1
    if (warp < 4) {
      const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t pat[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pat[j] = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) pat[8 * q + (4 * g + q)] = one | (one << 16);
        tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + kPatCol + 32u * g, pat);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if constexpr (kFixed) {
    // "ones" patterns: pattern l = columns [kPatCol + 8l, +8) of every row, ones at K = 4l, 4l+1, 4l+2
    if (warp < 4) {
      const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
      uint32_t pat[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pat[j] = 0u;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        pat[8 * l + 2 * l] = one | (one << 16);   // K = 4l (low half), 4l + 1 (high half)
        pat[8 * l + 2 * l + 1] = one;             // K = 4l + 2
      }
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + kPatCol, pat);
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // Loads of one tile into slot s: the feature tile image and, when staged, the bias rows of its 8 rays
  // (consecutive rows of row_bias in the BLOCKED order).  Issuing a bulk copy costs its thread well over a
  // hundred cycles, so the work is spread over lane 0 of the slot's 8 warps (wi = warp index within the slot).
  auto issue_tile_loads = [&](int s, int64_t tile, int wi) {
    if (!rb_staged) {
      const uint32_t piece = a.in_bytes / kTcEpiWarpsPerSlot;  // K * 32 bytes: a multiple of 16
      mbar_arrive_expect_tx(&bars[BAR_IN_FULL + s], piece);
      bulk_g2s(smem + a.in_off[s] + wi * piece, a.in + tile * (int64_t)a.in_bytes + wi * piece, piece, &bars[BAR_IN_FULL + s]);
      return;
    }
    // staged per-ray bias rows: warps 0-3 copy a quarter of the image each, warps 4-7 two bias rows each — one
    // arrival and at most two copies per warp (measured: with image piece + bias row on EVERY warp the loads cost each
    // warp ~1 100 cycles after the layer-0 arrival, more than layer 1's MMAs hide)
    if (wi < 4) {
      const uint32_t piece = a.in_bytes / 4;   // K * 64 bytes
      mbar_arrive_expect_tx(&bars[BAR_IN_FULL + s], piece);
      bulk_g2s(smem + a.in_off[s] + wi * piece, a.in + tile * (int64_t)a.in_bytes + wi * piece, piece, &bars[BAR_IN_FULL + s]);
    } else {
      const TcLayer& rl = a.layer[rb_layer];
      const int64_t ray0 = (tile / a.tiles_per_blk) * kBlkRays + 2 * (wi - 4);
      const uint32_t row_bytes = (uint32_t)(rl.n * 4);
      const int n_rows = ray0 + 1 < a.n_rays ? 2 : (ray0 < a.n_rays ? 1 : 0);
      float* dst = reinterpret_cast<float*>(smem + a.rb_off[s]) + 2 * (wi - 4) * kRbPitch;
      // padding rays of the last ray block: a zero bias row, so that their (unused) activations stay finite — the dense
      // weight gradient multiplies them by a zero delta, and 0 x Inf would poison the sum
      for (int k = n_rows; k < 2; ++k)
        for (int i = 0; i < rl.n; i += 4) *reinterpret_cast<float4*>(dst + k * kRbPitch + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n_rows == 0) {
        mbar_arrive(&bars[BAR_RB_FULL + s]);
      } else {
        mbar_arrive_expect_tx(&bars[BAR_RB_FULL + s], row_bytes * (uint32_t)n_rows);
        for (int k = 0; k < n_rows; ++k)
          bulk_g2s(dst + k * kRbPitch, rl.row_bias + (ray0 + k) * rl.n, row_bytes, &bars[BAR_RB_FULL + s]);
      }
    }
  };

  {
    // ================= slot s: epilogue warps, the last one to finish a layer issues the next MMAs ======
    const int ew = warp;
    const int s = ew >> 3;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (ew & 7) >> 2;       // column ownership: half h owns columns [64h, 64h+64) of every layer
    const int r = quad * 32 + lane;       // row within the tile == TMEM lane
    const int col0 = half * 64;
    const int wi = ew & 7;                // warp index within the slot (also its share of the tile loads)
    const uint32_t d_base = tmem_base + (uint32_t)s * kSlotCols;  // lane 0: the MMA's D / A operands
    const uint32_t d_tmem = d_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
    const uint32_t a_tmem = d_base + ((uint32_t)(quad * 32) << 16) + kSlotAOff + (uint32_t)(col0 >> 1);
    const float* rb_row = reinterpret_cast<const float*>(smem + a.rb_off[s]) + (r & (kBlkRays - 1)) * kRbPitch + col0;
    float* hp = shpart + (s * 128 + r) * 4;
    uint64_t* bar_acc_full = &bars[BAR_ACC_FULL + s];
    uint64_t* bar_rb_full = &bars[BAR_RB_FULL + s];
    uint32_t* done_cnt = tmem_slot + 1 + s;  // arrivals of the slot's 8 warps (monotonic; every 8th issues)
    uint64_t* bar_done = &bars[BAR_DONE + s];
    (void)done_cnt, (void)bar_done;
    const uint64_t adesc0 = umma_desc(smem_u32(smem + a.in_off[s]), 2048u, 128u);
    uint32_t ph_acc = 0, ph_rb = 0;
#ifdef NVSR_TC_TIMING
    unsigned long long t_sum[8][8] = {};
    int t_layer = 0;
#endif

    // MMAs of layer l of the slot's tile number `use` (whole warp; one elected lane issues).
    //   layer 0: A = feature tile in smem (SS form); l > 0: A = activations in TMEM (TS form).
    // Every MMA accumulates onto the bias the epilogue pre-loaded into D_s.
    // the rgb head of the tile whose last activations sit in the slot's A region: 8 K steps of an N = 16 MMA
    auto issue_head_mmas = [&]() {
      const uint32_t d_head = tmem_base + kHeadCol + 16u * (uint32_t)s;
      const uint64_t hdesc = umma_desc(smem_u32(smem + a.himg_off), 256u, 128u);   // LBO = 16 n * 16 B, + 2 chunks per K step
      const uint32_t idesc16 = umma_idesc_16(16, F16);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_ts(d_head, d_base + kSlotAOff + (uint32_t)ks * 8u, hdesc + (uint64_t)(ks * 32), idesc16, ks > 0 ? 1u : 0u);
    };
    auto issue_head_only = [&]() {
      tc_fence_after();
      if (elect_one()) {
        issue_head_mmas();
        umma_commit(bar_acc_full);
      }
      __syncwarp();
    };
    auto issue_layer = [&](int l, uint32_t use, bool with_head = false) {
      const TcLayer& ly = a.layer[l];
      const int n = kFixed ? 128 : ly.n;
      const uint32_t idesc = umma_idesc_16(n, F16);
      const uint32_t b_lbo = (uint32_t)n * 16u;
      const uint64_t bdesc0 = umma_desc(smem_u32(smem + ly.w_off), b_lbo, 128u);
      const uint32_t b_step = (2u * b_lbo) >> 4;  // descriptor address units per K step of 16
      const int ksteps = ly.k >> 4;
      if (l == 0) {
        mbar_wait(&bars[BAR_W], 0);
        mbar_wait(&bars[BAR_IN_FULL + s], use & 1);
      }
      tc_fence_after();
      if (elect_one()) {
        if constexpr (kHeadTC) {
          if (with_head) issue_head_mmas();
        }
        if constexpr (kFixed) {
          // bias as a K step (see kPatCol); a layer whose bias is per ray (RB0, layer 0) keeps the pre-stored
          // accumulator, every other layer starts from a clean accumulator (first MMA overwrites)
          const bool kstep_bias = !(RB0 != 0 && l == 0);
          if (l == 0) {
            for (int ks = 0; ks < ksteps; ++ks)
              umma_ss(d_base, adesc0 + (uint64_t)(ks * 256), bdesc0 + (uint64_t)(ks * b_step), idesc,
                      (ks > 0 || !kstep_bias) ? 1u : 0u);
          } else {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ts(d_base, d_base + kSlotAOff + (uint32_t)ks * 8u, bdesc0 + (uint64_t)(ks * b_step), idesc, ks > 0 ? 1u : 0u);
          }
          if (kstep_bias)
            umma_ts(d_base, tmem_base + kPatCol + 8u * (uint32_t)l, umma_desc(smem_u32(smem + a.bimg_off), 2048u, 128u), idesc, 1u);
        } else {
          const bool kstep_bias = l != rb_layer;   // a per-ray bias was pre-stored into the accumulator
          if (l == 0) {
            for (int ks = 0; ks < ksteps; ++ks)
              umma_ss(d_base, adesc0 + (uint64_t)(ks * 256), bdesc0 + (uint64_t)(ks * b_step), idesc, (ks > 0 || !kstep_bias) ? 1u : 0u);
          } else {
            for (int ks = 0; ks < ksteps; ++ks)
              umma_ts(d_base, d_base + kSlotAOff + (uint32_t)ks * 8u, bdesc0 + (uint64_t)(ks * b_step), idesc,
                      (ks > 0 || !kstep_bias) ? 1u : 0u);
          }
          if (kstep_bias)
            umma_ts(d_base, tmem_base + kPatCol + 8u * (uint32_t)l, umma_desc(smem_u32(smem + a.bimg_off), 2048u, 128u), idesc, 1u);
        }
        umma_commit(bar_acc_full);
      }
      __syncwarp();
    };
    // This warp's TMEM writes for the next accumulation are done: count it; the slot's 8th arrival
    // issues the MMAs of (layer nl, tile number nuse) — no issuer warp, no wake-up hop.
    auto arrive_then_issue = [&](bool issue_next, int nl, uint32_t nuse, bool head = false) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      uint32_t last_in = 0;
      if (lane == 0) {
#ifndef NVSR_TC_ATOMIC_ARRIVE
        // count the arrival on an mbarrier (release, no MEMBAR): the state it returns holds the pending count
        // before this arrival, 1 = this warp completes the phase and is the issuer; the test_wait on that
        // (now complete) phase is the acquire side towards the other seven warps' arrivals
        uint64_t state;
        uint32_t pending;
        asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar_done)) : "memory");
        asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
        last_in = pending == 1;
        if (last_in) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(bar_done)), "l"(state) : "memory");
          (void)ok;
        }
#else
        uint32_t old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(done_cnt)) : "memory");
        last_in = (old & (kTcEpiWarpsPerSlot - 1)) == kTcEpiWarpsPerSlot - 1;
#endif
      }
      last_in = __shfl_sync(0xffffffffu, last_in, 0);
#ifdef NVSR_TC_TIMING
      const long long ti0 = clock64();
#endif
      if (last_in && issue_next) issue_layer(nl, nuse, head);
      else if (last_in && head) issue_head_only();
#ifdef NVSR_TC_TIMING
      if (last_in && issue_next) t_sum[t_layer][5] += 1, t_sum[t_layer][6] += clock64() - ti0;
#endif
    };

    // source of the bias rows of `layer` for this thread: mode 1 = smem pointer, 2 = global pointer
    auto bias_src = [&](int layer, int64_t tile, int* mode) -> const float* {
      if (layer != rb_layer) {
        *mode = 1;
        return sbias + layer * 128 + col0;
      }
      if (rb_staged) {
        *mode = 1;
        return rb_row;
      }
      const TcLayer& ly = a.layer[layer];
      int64_t row = tile * kTileRows + r;
      int64_t ray;
      if (a.row_ids) {  // sparse list: the entry names a BLOCKED row of the chunk -> its ray
        const uint32_t rid = row < rows ? (uint32_t)__ldg(a.row_ids + row) : 0u;
        ray = (int64_t)((rid >> 7) / (uint32_t)a.tiles_per_blk) * kBlkRays + (rid & (kBlkRays - 1));
      } else {
        ray = rows < 0x7fffffff ? (int64_t)((uint32_t)row / (uint32_t)a.samples_per_ray) : row / a.samples_per_ray;
      }
      if (ray >= a.n_rays) ray = a.n_rays - 1;
      *mode = 2;
      return ly.row_bias + ray * ly.n + col0;
    };

    // tensor-core head read-back of `tile` (kHeadTC): this thread's row, columns hi 0-2 + lo 3-5 of the slot's head block
    auto write_tc_head = [&](int64_t tile) {
      uint32_t hv[8];
      tmem_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + kHeadCol + 16u * (uint32_t)s, hv);
      tmem_ld_wait();
      const TcLayer& lh = a.layer[L - 1];
      int64_t row = tile * kTileRows + r;
      if (row < rows) {
        if (a.row_ids) row = __ldg(a.row_ids + row);  // sparse list: write the row the entry stands for
#pragma unroll
        for (int h = 0; h < (HN > 0 ? HN : 1); ++h)
          a.raw[(int64_t)(lh.head_ch + h) * a.raw_stride + row] =
              __uint_as_float(hv[h]) + __uint_as_float(hv[3 + h]) + __ldg(lh.head_b + h);
      }
    };

    // prologue: weights (once per CTA) and the first tile's loads, then D_s <- its layer-0 bias
    const int64_t first = blockIdx.x + (int64_t)s * G;
    // (sparse mode: a CTA that the list leaves without a tile must not start copies it will never wait for)
    if (warp == 0 && lane == 0 && blockIdx.x < n_tiles) {
      mbar_arrive_expect_tx(&bars[BAR_W], a.w_bytes_total);
      for (int l = 0; l < L; ++l) {
        const TcLayer& ly = a.layer[l];
        bulk_g2s(smem + ly.w_off, ly.w, (uint32_t)(ly.k * ly.n * 2), &bars[BAR_W]);
      }
    }
    if (first < n_tiles) {
      if (lane == 0) issue_tile_loads(s, first, wi);
      if (rb_layer == 0 && rb_staged) {
        mbar_wait(bar_rb_full, ph_rb);
        ph_rb ^= 1;
      }
      if (kFixed ? RB0 != 0 : rb_layer == 0) {
        int mode;
        const float* bsrc = bias_src(0, first, &mode);
        const int n0 = kFixed ? 128 : a.layer[0].n;
        float dummy[4];
        for (int c = 0; c < 64; c += 32)
          if (col0 + c < n0) epi_pass<F16>(d_tmem + (uint32_t)c, 0u, false, false, false, 0, nullptr, dummy, mode, bsrc + c);
      }
      arrive_then_issue(true, 0, 0u);
    }

    for (uint32_t use = 0;; ++use) {
      const int64_t tile = blockIdx.x + (int64_t)(2 * use + s) * G;
      if (tile >= n_tiles) break;
      const int64_t next_tile = tile + 2 * G;
      const bool next_valid = next_tile < n_tiles;
      // fixed chains: the layer loop is fully unrolled, so `last`, the bias source and the next layer are
      // compile-time per copy (~100 instructions per layer and warp instead of ~350: the kernel is largely
      // issue-bound, measured -18% / -8% on the density / rgb chains); the generic chain stays rolled
#ifdef NVSR_TC_ROLLED
#pragma unroll 1
#else
#pragma unroll(kFixed ? 4 : (LC < 0 ? -LC : 1))
#endif
      for (int l = 0; l < L; ++l) {
        // ---- everything that does not depend on the accumulator: done before the wait ----
        const TcLayer& ly = a.layer[l];
        const bool last = l == L - 1;
        const int n_cur = kFixed ? 128 : ly.n;
        const bool relu = kFixed ? true : (ly.relu != 0);
        const int head_n = kFixed ? (last ? HN : 0) : (ly.head_w ? ly.head_n : 0);
        const float* hw = sheadw + (kFixed ? 0 : ly.head_row) * 128 + col0;
        // what gets pre-loaded into D_s once this layer's accumulator has been read
        const int nl = last ? 0 : l + 1;
        const int n_next = (last && !next_valid) ? 0 : (kFixed ? 128 : a.layer[nl].n);
        const bool nl_rb = n_next > 0 && nl == rb_layer && rb_staged;
        int mode = 0;
        const float* bsrc = (n_next > 0 && (kFixed ? (RB0 != 0 && last) : nl == rb_layer)) ? bias_src(nl, last ? next_tile : tile, &mode) : nullptr;
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
        if (nl_rb) {
          mbar_wait(bar_rb_full, ph_rb);
          ph_rb ^= 1;
        }
        TC_T(tt0);
        mbar_wait(bar_acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        TC_T(tt1);
#ifndef NVSR_TC_OLD_EPI
        if constexpr (kFixed) {
          // fixed chains: bias rows always come from shared memory (mode 1)
          // (hidden-layer biases ride in the MMA: only the per-ray layer-0 bias of the next tile is pre-stored)
          uint4* act = nullptr;
          if constexpr (TRAIN)
            act = reinterpret_cast<uint4*>(a.act[l] + tile * (int64_t)(kTileRows * 128 * 2)) + (col0 >> 3) * 128 + r;
          if constexpr (kHeadTC) {
            // layer 0 of a later tile: the previous tile's head (issued with this layer's MMAs) is complete too —
            // 4 of the slot's warps read it back (8 columns per row: hi 0-2 + lo 3-5) and write the heads out
            if (l == 0 && use > 0 && half == 0) write_tc_head(tile - 2 * G);
            epi_fixed<F16, 0>(d_tmem, a_tmem, hw, hacc, last && RB0 != 0 && n_next > 0, bsrc, act);
          } else if (last) epi_fixed<F16, HN>(d_tmem, a_tmem, hw, hacc, RB0 != 0 && n_next > 0, bsrc, act);
          else epi_fixed<F16, 0>(d_tmem, a_tmem, hw, hacc, false, bsrc, act);
        } else
#endif
        {
          const bool read = col0 < n_cur;
          const bool bias = col0 < n_next && nl == rb_layer;   // every other bias rides in the MMA
          if (read || bias) epi_pass64<F16>(d_tmem, a_tmem, read, !last, relu, head_n, hw, hacc, bias ? mode : 0, bsrc);
        }
#ifdef NVSR_TC_TIMING
        t_layer = l & 7;
#endif
        TC_T(tt2);
        arrive_then_issue(n_next > 0, nl, last ? use + 1 : use, kHeadTC && last);
        TC_T(tt3);
        // layer 0's MMAs have completed: the slot's ring buffer (and, since all 8 warps finished reading
        // them before those MMAs were issued, its staged bias rows) may take the slot's next tile.  Issued
        // after the arrival: the warp would only be waiting for layer 1's accumulator now.
        if (l == 0 && next_valid && lane == 0) issue_tile_loads(s, next_tile, wi);

        if (!kHeadTC && head_n > 0) {
          // combine the two column halves of a row: half 1 -> smem -> half 0
          if (half == 1) *reinterpret_cast<float4*>(hp) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
          named_bar_sync(1 + s * 4 + quad, 64);  // the two warps sharing this slot and lane quadrant
          int64_t row = tile * kTileRows + r;
          if (half == 0 && row < rows) {
            if (a.row_ids) row = __ldg(a.row_ids + row);  // sparse list: write the row the entry stands for
            float4 o = *reinterpret_cast<const float4*>(hp);
            float hv[4] = {hacc[0] + o.x, hacc[1] + o.y, hacc[2] + o.z, hacc[3] + o.w};
#pragma unroll
            for (int h = 0; h < 4; ++h)
              if (h < head_n) a.raw[(int64_t)(ly.head_ch + h) * a.raw_stride + row] = hv[h] + __ldg(ly.head_b + h);
          }
        }
#ifdef NVSR_TC_TIMING
        t_sum[l & 7][0] += tt1 - tt0, t_sum[l & 7][1] += tt2 - tt1, t_sum[l & 7][2] += tt3 - tt2, t_sum[l & 7][3] += clock64() - tt3, t_sum[l & 7][4] += 1;
#endif
      }
    }
    if constexpr (kHeadTC) {
      // drain: the head of the slot's last tile was committed alone
      if (first < n_tiles) {
        const int64_t last_tile = first + ((n_tiles - 1 - first) / (2 * G)) * (2 * G);
        mbar_wait(bar_acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        if (half == 0) write_tc_head(last_tile);
      }
    }
#ifdef NVSR_TC_TIMING
    if (blockIdx.x == 0 && lane == 0)
      for (int l = 0; l < 8; ++l)
        for (int i = 0; i < 8; ++i) g_tc_timing[warp][l][i] = t_sum[l][i];
#endif
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int32_t launch_mlp_tc(const nvsr_mlp_t* m, cudaStream_t st, void* const* act_out) {
  TcArgs a;
  for (int l = 0; l < 4; ++l) a.act[l] = act_out ? (uint8_t*)act_out[l] : nullptr;
  a.n_layers = m->n_layers;
  a.rb_layer = -1;
  uint32_t off = 0;
  int head_rows = 0;
  bool uniform = true;  // all layers 128 wide with ReLU, one head on the last layer
  for (int l = 0; l < m->n_layers; ++l) {
    const nvsr_layer_t& L = m->layer[l];
    if (L.k <= 0 || (L.k % 16) != 0 || L.k > 256) return NVSR_ERR_UNSUPPORTED;
    if (L.n_out < 64 || L.n_out > 128 || (L.n_out % 64) != 0) return NVSR_ERR_UNSUPPORTED;
    if (l > 0 && L.k > 128) return NVSR_ERR_UNSUPPORTED;  // hidden activations live in 64 TMEM columns
    if (!L.w || (!L.bias && !L.row_bias)) return NVSR_ERR_INVALID_ARG;
    if (!aligned16(L.w) || (L.row_bias && !aligned16(L.row_bias))) return NVSR_ERR_ALIGNMENT;
    if (l > 0 && L.k != m->layer[l - 1].n_out) return NVSR_ERR_INVALID_ARG;
    if (L.row_bias) {
      if (a.rb_layer >= 0) return NVSR_ERR_UNSUPPORTED;  // one per-ray bias layer per chain
      a.rb_layer = l;
    }
    TcLayer& t = a.layer[l];
    t.w = L.w, t.bias = L.bias, t.row_bias = L.row_bias, t.head_w = L.head_w, t.head_b = L.head_b;
    t.k = L.k, t.n = L.n_out, t.relu = L.relu, t.head_n = L.head_n, t.head_ch = L.head_ch, t.head_row = 0;
    if (L.head_w) {
      if (L.head_n <= 0 || L.head_n > 4 || !L.head_b || L.head_ch < 0 || L.head_ch + L.head_n > 4) return NVSR_ERR_INVALID_ARG;
      if (head_rows + L.head_n > kTcMaxHeadRows) return NVSR_ERR_UNSUPPORTED;
      t.head_row = head_rows;
      head_rows += L.head_n;
    } else {
      t.head_n = 0;
    }
    uniform = uniform && L.n_out == 128 && L.relu && ((L.head_w != nullptr) == (l == m->n_layers - 1));
    t.w_off = off;
    off += (uint32_t)(L.k * L.n_out * 2);
  }
  const nvsr_layer_t& lastL = m->layer[m->n_layers - 1];
  if (!lastL.head_w) return NVSR_ERR_INVALID_ARG;  // the chain must end in a head
  a.w_bytes_total = off;
  const uint32_t in_bytes = (uint32_t)m->layer[0].k * 256u;  // 128 rows * K * 2 B
  a.in_off[0] = off, off += in_bytes;
  a.in_off[1] = off, off += in_bytes;
  // per-ray bias rows are staged through smem when the 8 rays of a tile are consecutive (BLOCKED order)
  const bool sparse = m->row_ids != nullptr;
  if (sparse && (!m->row_count || m->row_order != NVSR_ROWS_BLOCKED)) return NVSR_ERR_INVALID_ARG;
  a.row_ids = m->row_ids, a.row_count = sparse ? m->row_count : nullptr;
  a.rb_staged = (a.rb_layer == 0 && m->row_order == NVSR_ROWS_BLOCKED && !sparse) ? 1 : 0;
  a.rb_off[0] = a.rb_off[1] = off;
  if (a.rb_staged) {
    a.rb_off[1] = off + kRbRowsMax * kRbPitch * 4u;
    off += 2u * kRbRowsMax * kRbPitch * 4u;
  }
  // which specialisation will run (decided below with the same predicates): the fixed chains keep their biases in an
  // image, the rgb one also its head weights — they do not need the fp32 tables / the half-combining scratch
  const bool fixed_any = uniform && m->n_layers == 4 && (lastL.head_n == 1 ? a.rb_layer < 0 : (lastL.head_n == 3 && a.rb_layer == 0)) &&
                         (a.rb_layer < 0 || a.rb_staged || sparse);
  a.bias_off = off;                                          // (fp32 bias table: no longer used, every bias is in the image)
  a.bimg_off = off, off += kBiasImgBytes;                    // (16-byte aligned: every size above is a multiple of 16)
  const bool fixed_rgb = fixed_any && lastL.head_n == 3;
  a.himg_off = off, off += fixed_rgb ? kHeadImgBytes : 0u;
  a.headw_off = off, off += fixed_rgb ? 0u : kTcMaxHeadRows * 128u * 4u;
  a.hpart_off = off, off += fixed_rgb ? 0u : 2u * 128u * 4u * 4u;
  a.bar_off = off, off += BAR_COUNT * 8u + 16u;  // barriers, then {tmem base, arrival counter x2}
  const uint32_t smem_bytes = off;
  if (smem_bytes > 227u * 1024u) return NVSR_ERR_RESOURCE;
  if (!aligned16(m->in)) return NVSR_ERR_ALIGNMENT;

  a.in = (const uint8_t*)m->in;
  a.in_bytes = in_bytes;
  a.rows = m->rows;
  a.n_tiles = ceil_div64(m->rows, kTileRows);
  a.samples_per_ray = m->samples_per_ray > 0 ? m->samples_per_ray : 1;
  a.row_order = m->row_order;
  a.tiles_per_blk = tiles_per_block(a.samples_per_ray);
  if (a.row_order == NVSR_ROWS_BLOCKED && !sparse && (m->rows % kTileRows) != 0) return NVSR_ERR_INVALID_ARG;
  a.n_rays = m->n_rays > 0 ? m->n_rays : 1;
  a.raw = m->raw;
  a.raw_stride = m->raw_stride;

  const bool f16 = m->precision == NVSR_F16;
  void (*kernel)(TcArgs) = nullptr;
  // compile-time specialisations: the two decoders of the tri-plane model
  const bool rb_ok = a.rb_layer < 0 || a.rb_staged;
  if (act_out) {
    // training forward: the two fp16 tri-plane chains only, BLOCKED rows, dense or over a row list (sparse path: input
    // and activation images in LIST order, heads to the listed rows)
    for (int l = 0; l < 4; ++l)
      if (!act_out[l] || !aligned16(act_out[l])) return NVSR_ERR_INVALID_ARG;
    if (!(uniform && m->n_layers == 4 && f16 && m->row_order == NVSR_ROWS_BLOCKED)) return NVSR_ERR_UNSUPPORTED;
    if (lastL.head_n == 1 && a.rb_layer < 0) kernel = mlp_chain_tc_kernel<true, 4, 1, 0, true>;   // dense or over a row list
    else if (lastL.head_n == 3 && a.rb_layer == 0 && sparse) kernel = mlp_chain_tc_kernel<true, 4, 3, 2, true>;
    else if (lastL.head_n == 3 && a.rb_layer == 0 && a.rb_staged) kernel = mlp_chain_tc_kernel<true, 4, 3, 1, true>;
    else return NVSR_ERR_UNSUPPORTED;
  } else if (uniform && m->n_layers == 4 && rb_ok && lastL.head_n == 1 && a.rb_layer < 0)
    kernel = f16 ? mlp_chain_tc_kernel<true, 4, 1, 0> : mlp_chain_tc_kernel<false, 4, 1, 0>;
  else if (uniform && m->n_layers == 4 && lastL.head_n == 3 && a.rb_layer == 0 && sparse)
    kernel = f16 ? mlp_chain_tc_kernel<true, 4, 3, 2> : mlp_chain_tc_kernel<false, 4, 3, 2>;
  else if (uniform && m->n_layers == 4 && rb_ok && lastL.head_n == 3 && a.rb_layer == 0)
    kernel = f16 ? mlp_chain_tc_kernel<true, 4, 3, 1> : mlp_chain_tc_kernel<false, 4, 3, 1>;
  else
#ifndef NVSR_TC_MIP_ROLLED
  if (m->n_layers == 6)  // the mip decoder (FlexibleNeRFModel): run-time layer shapes, layer loop unrolled
    kernel = f16 ? mlp_chain_tc_kernel<true, -6, 0, 0> : mlp_chain_tc_kernel<false, -6, 0, 0>;
  else
#endif
    kernel = f16 ? mlp_chain_tc_kernel<true, 0, 0, 0> : mlp_chain_tc_kernel<false, 0, 0, 0>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int32_t)e;
  int64_t grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  kernel<<<(unsigned)grid, kTcThreads, smem_bytes, st>>>(a);
  NVSR_RETURN_LAST_ERROR();
}

}  // namespace nvsr

#ifdef NVSR_TC_TIMING
extern "C" int32_t nvsr_debug_tc_timing(unsigned long long* host_out /* [16][8][8] */) {
  return (int32_t)cudaMemcpyFromSymbol(host_out, nvsr::g_tc_timing, sizeof(unsigned long long) * 16 * 8 * 8);
}
#endif
