// §8f rank 2 — the step immediately BEFORE the gather: finishing a plane super-resolution.
//
// PlanesSR.forward (models.py:884-926) ends with  out = EDSR(pad(LR))[crop] + interpolate(LR, x scale, bilinear,
// align_corners)  and then caches the plane on the CPU (:925) and re-uploads it for every network chunk (:893); the
// gather then needs it re-packed.  sr_finalize_kernel fuses everything after the conv chain: it reads the conv chain's
// output where it lies (channels-last, any float type, with the HR_overpadding crop as an offset), computes the bilinear
// up-sampling of the LR plane on the fly (ATen's align_corners formulas), adds, and writes DIRECTLY the image the gather
// reads — 16-bit x-pair records or fp32 channels-last — plus, optionally, the fp32 NCHW plane the reference's own
// consumers expect.  The SR plane is born in its final layout, on the device, once per scene.
#include "common.cuh"

namespace nvsr {

struct SrArgs {
  const void* diff;      // conv-chain output, channels-last [Hd][Wd][C] (+ crop offset), element type diff_dtype
  int diff_dtype;        // NVSR_F32 | NVSR_BF16 | NVSR_F16
  int64_t diff_row_stride, diff_px_stride;  // in elements
  int crop;              // HR_overpadding: out(y,x) reads diff(y + crop, x + crop)
  const float* lr;       // LR plane NCHW fp32 [C][rh][rw]
  int C, rh, rw, scale;  // output is [rh*scale][rw*scale]
  int align_corners;
  void* packed;          // gather image or NULL
  int packed_dtype;      // NVSR_F32: channels-last [RH][RW][C]; 16-bit: x-pair records [RH][C/8][RW][2][8]
  float* nchw;           // optional fp32 [C][RH][RW]
};

__device__ __forceinline__ float load_diff(const SrArgs& a, int64_t idx) {
  if (a.diff_dtype == NVSR_F32) return reinterpret_cast<const float*>(a.diff)[idx];
  if (a.diff_dtype == NVSR_F16) return __half2float(reinterpret_cast<const __half*>(a.diff)[idx]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.diff)[idx]);
}

// ATen upsample_bilinear2d source index (UpSample.h: area_pixel_compute_source_index)
__device__ __forceinline__ void src_index(int dst, int in, int out, int align_corners, int* i0, int* i1, float* l0, float* l1) {
  float src;
  if (align_corners) {
    const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * (float)dst;
  } else {
    const float sc = (float)in / (float)out;
    src = fmaxf(__fsub_rn(__fmul_rn(sc, __fadd_rn((float)dst, 0.5f)), 0.5f), 0.f);
  }
  const int lo = min((int)src, in - 1);
  *i0 = lo, *i1 = lo + ((lo < in - 1) ? 1 : 0);
  *l1 = __fsub_rn(src, (float)lo), *l0 = __fsub_rn(1.f, *l1);
}

// value of the super-resolved plane at (c, y, x)
__device__ __forceinline__ float sr_value(const SrArgs& a, int c, int y, int x, int y0, int y1, float ly0, float ly1) {
  int x0, x1;
  float lx0, lx1;
  src_index(x, a.rw, a.rw * a.scale, a.align_corners, &x0, &x1, &lx0, &lx1);
  const float* p = a.lr + (int64_t)c * a.rh * a.rw;
  const float top = __fadd_rn(__fmul_rn(lx0, __ldg(p + y0 * a.rw + x0)), __fmul_rn(lx1, __ldg(p + y0 * a.rw + x1)));
  const float bot = __fadd_rn(__fmul_rn(lx0, __ldg(p + y1 * a.rw + x0)), __fmul_rn(lx1, __ldg(p + y1 * a.rw + x1)));
  const float up = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
  const float d = load_diff(a, (int64_t)(y + a.crop) * a.diff_row_stride + (int64_t)(x + a.crop) * a.diff_px_stride + c);
  return __fadd_rn(d, up);   // torch.add(difference, residual_plane)  (models.py:917)
}

// one thread per (y, 8-channel chunk, x): the x-pair record of that texel, the fp32 channels-last chunk, the NCHW values
template <bool F16>
__global__ void __launch_bounds__(256) sr_finalize_kernel(SrArgs a) {
  const int RH = a.rh * a.scale, RW = a.rw * a.scale, CH = a.C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)RH * CH * RW) return;
  const int x = (int)(idx % RW);
  const int ch = (int)((idx / RW) % CH);
  const int y = (int)(idx / ((int64_t)RW * CH));
  int y0, y1;
  float ly0, ly1;
  src_index(y, a.rh, RH, a.align_corners, &y0, &y1, &ly0, &ly1);
  float v[8], vr[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = sr_value(a, ch * 8 + e, y, x, y0, y1, ly0, ly1);
  if (a.nchw) {
#pragma unroll
    for (int e = 0; e < 8; ++e) a.nchw[((int64_t)(ch * 8 + e) * RH + y) * RW + x] = v[e];
  }
  if (!a.packed) return;
  if (a.packed_dtype == NVSR_F32) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.packed) + ((int64_t)y * RW + x) * a.C + ch * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  const int xr = min(x + 1, RW - 1);   // the record's right half: the same chunk of the texel to the right
#pragma unroll
  for (int e = 0; e < 8; ++e) vr[e] = xr == x ? v[e] : sr_value(a, ch * 8 + e, y, xr, y0, y1, ly0, ly1);
  uint4 lo, hi;
  lo.x = pack16x2<F16>(v[0], v[1]), lo.y = pack16x2<F16>(v[2], v[3]), lo.z = pack16x2<F16>(v[4], v[5]), lo.w = pack16x2<F16>(v[6], v[7]);
  hi.x = pack16x2<F16>(vr[0], vr[1]), hi.y = pack16x2<F16>(vr[2], vr[3]), hi.z = pack16x2<F16>(vr[4], vr[5]), hi.w = pack16x2<F16>(vr[6], vr[7]);
  uint4* rec = reinterpret_cast<uint4*>(a.packed) + (((int64_t)y * CH + ch) * RW + x) * 2;
  rec[0] = lo, rec[1] = hi;
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_sr_finalize(const void* diff, int32_t diff_dtype, int64_t diff_row_stride, int64_t diff_px_stride,
                                    int32_t crop, const float* lr_nchw, int32_t channels, int32_t rh, int32_t rw,
                                    int32_t scale, int32_t align_corners, void* packed, int32_t packed_dtype, float* nchw_out,
                                    void* stream) {
  NVSR_CHECK_ARG(diff && lr_nchw && (packed || nchw_out) && channels > 0 && channels % 8 == 0 && rh > 0 && rw > 0 && scale >= 1);
  NVSR_CHECK_ARG(crop >= 0 && diff_px_stride >= channels && diff_row_stride >= diff_px_stride * (int64_t)rw * scale);
  NVSR_CHECK_ARG(diff_dtype == NVSR_F32 || is_16bit(diff_dtype));
  NVSR_CHECK_ARG(!packed || packed_dtype == NVSR_F32 || is_16bit(packed_dtype));
  if (packed && (reinterpret_cast<uintptr_t>(packed) & 31u) != 0) return NVSR_ERR_ALIGNMENT;
  SrArgs a{diff, diff_dtype, diff_row_stride, diff_px_stride, crop, lr_nchw, channels, rh, rw, scale, align_corners ? 1 : 0,
           packed, packed_dtype, nchw_out};
  const int64_t total = (int64_t)rh * scale * (channels / 8) * rw * scale;
  const int64_t blocks = ceil_div64(total, 256);
  NVSR_CHECK_ARG(blocks < ((int64_t)1 << 31));
  if (packed && packed_dtype == NVSR_BF16) sr_finalize_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  else sr_finalize_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  NVSR_RETURN_LAST_ERROR();
}
