// a7+a8: alpha compositing with cumulative transmittance, inverse-CDF resampling and sort-merge.
//
// composite_kernel: one warp per block of 8 consecutive rays.  Lane = (ray-in-block rl = lane & 7,
// quarter q = lane >> 3); per 16-sample tile a lane owns 4 consecutive samples (16t + 4q .. +3) of its
// ray.  In the BLOCKED raw order (what the tcgen05 decoder writes: row = (s % 16) * 8 + ray % 8 inside a
// 128-row tile) every load instruction of a warp then covers four full 32-byte sectors, without any
// shared-memory staging or block-wide barrier.  The transmittance cumprod runs sequentially over the
// lane's 4 samples and as a 2-step shuffle scan over the 4 quarters, with a running carry; it is
// accumulated in fp64 and rounded to fp32 per prefix — what the reference's CPU path does (ATen
// accumulates float cumprod/cumsum in double) — so weights, cdf edges and with them the searchsorted
// bin indices track the oracle.  The map sums accumulate per tile in fp32 and across tiles in fp64.
// On the coarse pass the warp then resamples its 8 rays one after the other, all 32 lanes on one ray:
// pdf/cdf in ATen's summation orders, a fixed-trip-count branch-free searchsorted, and the
// sort(cat(z_vals, z_samples)) as a rank merge whose ranks start from the bin index the inversion just
// produced (z_samples[j] lies between the two bin centres around it), written out coalesced through a
// position bitmap.
#include "common.cuh"

namespace nvsr {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_up_d(double v, int off) { return __shfl_up_sync(kFull, v, off); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

// inclusive scans over the 32 lanes
__device__ __forceinline__ double warp_scan_add(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// total order used by the slow-path merge: NaN sorts last (torch.sort), ties by index
__device__ __forceinline__ bool key_less(float a, int ia, float b, int ib) {
  bool na = isnan(a), nb = isnan(b);
  if (na || nb) return (!na && nb) || (na && nb && ia < ib);
  return a < b || (a == b && ia < ib);
}

// searchsorted(cdf[0..n), u, side='right'): first index with cdf[idx] > u, n if none
__device__ __forceinline__ int upper_bound_f(const float* a, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] > u) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
// nerf_helpers.py:686-700 given cdf/bins (B entries) in shared memory
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u, int* ind_out) {
  int ind = upper_bound_f(cdf, B, u);
  int below = max(0, ind - 1);
  int above = min(B - 1, ind);
  float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  float bb = bins[below], ba = bins[above];
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// torch.sum(x, -1) of ATen's CPU kernel for a contiguous inner reduction of n fp32 values
// (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum -> row_sum -> multi_row_sum): 8-lane
// vectors, 4 interleaved vector accumulators, a 4-level cascade every 16 steps, then the leftover
// vectors, the scalar tail and the 8 lane partials added sequentially.  Floating-point sums are
// order dependent; reproducing THIS order makes `total` — and with it every cdf edge and every
// searchsorted index — bit-identical to the reference's CPU path for identical weights (probed:
// 100% agreement with torch.sum for n in 14..2050).  The 32 (accumulator, lane) pairs map onto the
// 32 lanes of the warp.  x(i) = w[i] + 1e-5 (nerf_helpers.py:672).
__device__ __forceinline__ float aten_sum_w(const float* w, int n, int lane) {
  auto X = [&](int i) { return __fadd_rn(w[i], 1e-5f); };
  if (n < 8) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, X(i));
    return s;
  }
  const int vec_size = n >> 3, size_ilp = vec_size >> 2;
  const int k = lane >> 3, L = lane & 7;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int i = 0;
  while (i + 16 <= size_ilp) {
    for (int j = 0; j < 16; ++j, ++i) acc[0] = __fadd_rn(acc[0], X(((i << 2) + k) * 8 + L));
    for (int j = 1; j < 4; ++j) {
      acc[j] = __fadd_rn(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (15 << (j * 4))) != 0) break;
    }
  }
  for (; i < size_ilp; ++i) acc[0] = __fadd_rn(acc[0], X(((i << 2) + k) * 8 + L));
  for (int j = 1; j < 4; ++j) acc[0] = __fadd_rn(acc[0], acc[j]);
  if (k == 0)
    for (int v = size_ilp << 2; v < vec_size; ++v) acc[0] = __fadd_rn(acc[0], X(v * 8 + L));
  float p = acc[0];
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 8));
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 16));
  p = __fadd_rn(p, __shfl_sync(kFull, acc[0], L + 24));
  float fin = 0.f;
  for (int idx = vec_size << 3; idx < n; ++idx) fin = __fadd_rn(fin, X(idx));
  for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, __shfl_sync(kFull, p, l));
  return fin;
}

// pdf/cdf of sample_pdf_2 (nerf_helpers.py:673-676) from weights w[0..nw) in shared memory:
// cdf[0]=0, cdf[i] = sum_{m<=i-1} (w[m]+1e-5)/total, i = 1..nw  -> nw+1 entries.
// cumsum: ATen accumulates fp32 cumsum in double on the CPU and rounds each prefix to fp32.
__device__ __forceinline__ void build_cdf(const float* w, int nw, float* cdf, int lane) {
  const float total = aten_sum_w(w, nw, lane);
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < nw; base += 32) {
    int i = base + lane;
    double pdf = (i < nw) ? (double)__fdiv_rn(__fadd_rn(w[i], 1e-5f), total) : 0.0;
    double inc = warp_scan_add(pdf, lane);
    if (i < nw) cdf[i + 1] = (float)(carry + inc);
    carry += shfl_d(inc, 31);
  }
}

struct CompositeArgs {
  int64_t n_rays;
  int S;
  const float* raw;
  int64_t raw_stride;
  const float* z;
  const float* rd;
  const float* noise;
  int white_bkgd, mip;
  float *rgb, *disp, *acc, *depth, *weights;
  int n_fine;
  const float* u;
  int u_per_ray;
  int64_t* inds;
  float* z_samples;
  float* z_merged;
};

// searchsorted(a[0..n), u, side='right') = #(a <= u) for sorted a, with a fixed trip count (top = the
// largest power of two <= n) and no divergent branches.  NaN u -> n, as torch.
__device__ __forceinline__ int upper_bound_fixed(const float* a, int n, int top, float u) {
  int pos = 0;
  if (top == 32) {  // 32 <= n < 64: the reference's 64-sample coarse pass
#pragma unroll
    for (int step = 32; step > 0; step >>= 1) {
      int p = pos + step;
      if (p <= n && !(a[p - 1] > u)) pos = p;
    }
  } else {
    for (int step = top; step > 0; step >>= 1) {
      int p = pos + step;
      if (p <= n && !(a[p - 1] > u)) pos = p;
    }
  }
  return pos;
}

// sigmoid for the colour channels: MUFU.EX2 + MUFU.RCP (abs error < 3e-7, far inside the 1e-3 map
// tolerance).  The density path below keeps expf: its weights feed the bit-exact index path.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// row pitch (floats) of the per-warp [8 rays][n] shared arrays: a multiple of 4 (16-byte stores) with
// pitch % 32 == 4, so the 8 rays of a quarter-warp store phase hit 8 distinct 16-byte bank groups.
__host__ __device__ inline int pitch4(int n) { return ((n - 4 + 31) / 32) * 32 + 4; }

// per-warp shared floats (coarse pass only): zv[8][P1] | w[8][P] | cdf[S] | bins[S] | zs[n_fine] | bitmap
__host__ __device__ inline int composite_warp_floats(int S, int S1, int n_fine) {
  if (n_fine <= 0) return 0;
  const int n = 8 * pitch4(S1) + 8 * pitch4(S) + 2 * S + n_fine + (S1 + n_fine + 31) / 32;
  return (n + 3) & ~3;  // every warp's slice stays 16-byte aligned
}

constexpr int kCompWarps = 4;  // warps per CTA at the usual sample counts (fewer when shared memory is short)
// Resident CTAs per SM (measured, profiles/README.md): the resampling pass is latency-bound on shared memory
// and short dependent chains and wants warps (7 CTAs, 72 registers; 8 CTAs spill); the compositing-only fine pass wants
// the next tile's 21 loads per lane in registers without spills (4 CTAs, 128 registers).
#ifndef NVSR_COMP_MINB_RESAMPLE
#define NVSR_COMP_MINB_RESAMPLE 7
#endif
#ifndef NVSR_COMP_MINB_PLAIN
#define NVSR_COMP_MINB_PLAIN 4
#endif
constexpr int kCompMinCtasResample = NVSR_COMP_MINB_RESAMPLE, kCompMinCtasPlain = NVSR_COMP_MINB_PLAIN;

// One lane's share of a 16-sample tile: 4 consecutive samples x (r,g,b,sigma), and depths s0 .. s0+4.
struct TileData {
  float c[4][4];  // [sample][channel]
  float zz[5];
};

template <bool BLOCKED>
__device__ __forceinline__ void load_tile(const CompositeArgs& a, TileData& d, const float* raw_blk, const float* zrow,
                                          int t, int q, int S, int S1, bool valid, bool vec_z) {
  const int s0 = t * kBlkSamples + 4 * q;
  if (vec_z && s0 + 4 <= S1) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(zrow + s0));
    d.zz[0] = v.x, d.zz[1] = v.y, d.zz[2] = v.z, d.zz[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) d.zz[j] = (s0 + j < S1) ? __ldg(zrow + s0 + j) : 0.f;
  }
  d.zz[4] = (s0 + 4 < S1) ? __ldg(zrow + s0 + 4) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool live = valid && s0 + j < S;  // padding rows are never read
    const float* rp = BLOCKED ? raw_blk + t * kTileRows + (4 * q + j) * kBlkRays : raw_blk + s0 + j;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) d.c[j][ch] = live ? __ldg(rp + ch * a.raw_stride) : 0.f;
  }
}

template <bool BLOCKED, bool RESAMPLE>
__global__ void __launch_bounds__(kCompWarps * 32, RESAMPLE ? kCompMinCtasResample : kCompMinCtasPlain)
composite_kernel(CompositeArgs a, int warps_per_cta) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int rl = lane & 7, q = lane >> 3;
  const int S = a.S;
  const int S1 = S + (a.mip ? 1 : 0);  // depth entries per ray
  const int nf = RESAMPLE ? a.n_fine : 0;
  const int P = pitch4(S), P1 = pitch4(S1);
  float* zsm = sm + (size_t)wid * composite_warp_floats(S, S1, nf);
  float* wsm = zsm + 8 * P1;
  float* cdf = wsm + 8 * P;
  float* bins = cdf + S;
  float* zs = bins + S;
  unsigned* bm = reinterpret_cast<unsigned*>(zs + nf);
  const int TS = tiles_per_block(S);
  const int64_t n_blocks = ceil_div64(a.n_rays, kBlkRays);
  const bool vec_z = (S1 & 3) == 0 && (reinterpret_cast<uintptr_t>(a.z) & 15u) == 0;
  const bool vec_w = (S & 3) == 0 && (reinterpret_cast<uintptr_t>(a.weights) & 15u) == 0;
  // resampling constants
  const int B = S - 1;  // bins (z_mid; mip: mids of mids)
  int top = 1;
  while (top * 2 <= B) top *= 2;
  const int M = S1 + nf;
  const int n_words = (M + 31) >> 5;
  // shared deterministic u (perturb == 0): the same for every ray, kept in registers
  const bool u_regs = RESAMPLE && nf <= 128 && !a.u_per_ray;
  float ureg[4] = {0.f, 0.f, 0.f, 0.f};
  if (u_regs) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nf) ureg[k] = __ldg(a.u + lane + 32 * k);
  }

  for (int64_t blk = (int64_t)blockIdx.x * warps_per_cta + wid; blk < n_blocks;
       blk += (int64_t)gridDim.x * warps_per_cta) {
    const int64_t ray = blk * kBlkRays + rl;
    const bool valid = ray < a.n_rays;
    const int64_t rayc = valid ? ray : a.n_rays - 1;  // clamped: loads stay in bounds, results are dropped
    const float* zrow = a.z + rayc * S1;
    const float* raw_blk = BLOCKED ? a.raw + blk * TS * (int64_t)kTileRows + rl : a.raw + rayc * S;
    TileData cur;
    load_tile<BLOCKED>(a, cur, raw_blk, zrow, 0, q, S, S1, valid, vec_z);
    const float dx = __ldg(a.rd + rayc * 3), dy = __ldg(a.rd + rayc * 3 + 1), dz = __ldg(a.rd + rayc * 3 + 2);
    const float dnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));

    double carry = 1.0;  // product of (1-alpha+1e-10) over the previous tiles
    double sr = 0.0, sg = 0.0, sb = 0.0, sd = 0.0, sa = 0.0;
    bool z_sorted = true;  // this lane's depths are non-decreasing (NaN -> false)
    for (int t = 0; t < TS; ++t) {
      // fine pass: the next tile's loads are in flight while this one is composited (the resampling
      // variant runs at 64 registers for occupancy and has no room for a second tile)
      TileData nxt;
      if (!RESAMPLE && t + 1 < TS) load_tile<BLOCKED>(a, nxt, raw_blk, zrow, t + 1, q, S, S1, valid, vec_z);
      const int s0 = t * kBlkSamples + 4 * q;
      const float(&zz)[5] = cur.zz;
      if (RESAMPLE) {  // keep the depths for the resampling pass (entries beyond S1 are zeros, never used)
        float* zd = zsm + rl * P1 + s0;
        if (s0 + 4 <= P1) *reinterpret_cast<float4*>(zd) = make_float4(zz[0], zz[1], zz[2], zz[3]);
        if (q == 3 && s0 + 4 < S1) zd[4] = zz[4];  // mip: the last edge when S is a multiple of 16
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (s0 + j + 1 < S1) z_sorted &= (zz[j] <= zz[j + 1]);
      }
      // ---- radiance samples ----
      float alpha[4], cr[4], cg[4], cb[4], zc[4];
      double tt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = s0 + j;
        const bool live = valid && s < S;
        alpha[j] = 0.f, cr[j] = 0.f, cg[j] = 0.f, cb[j] = 0.f, zc[j] = 0.f, tt[j] = 1.0;
        if (live) {
          float sig = cur.c[j][3];
          float dist;
          if (a.mip) {
            dist = __fsub_rn(zz[j + 1], zz[j]);
            zc[j] = __fmul_rn(0.5f, __fadd_rn(zz[j], zz[j + 1]));
          } else {
            dist = (s + 1 < S) ? __fsub_rn(zz[j + 1], zz[j]) : 1e10f;
            zc[j] = zz[j];
          }
          dist = __fmul_rn(dist, dnorm);
          if (a.noise) sig = __fadd_rn(sig, __ldg(a.noise + ray * S + s));
          // sigma + noise <= 0: relu gives 0, alpha = 1 - exp(-0) = 0 and the weight alpha*T is exactly 0, so the
          // sample's colour contributes exactly 0 whatever it is (a finite sigmoid in the reference).  Its colour
          // channels are not touched here: the sparse decoder path does not even evaluate them (nvsr_keep_rows).
          const bool lit = !(sig <= 0.f);
          cr[j] = lit ? sigmoid_fast(cur.c[j][0]) : 0.f;
          cg[j] = lit ? sigmoid_fast(cur.c[j][1]) : 0.f;
          cb[j] = lit ? sigmoid_fast(cur.c[j][2]) : 0.f;
          sig = fmaxf(sig, 0.f);
          alpha[j] = __fsub_rn(1.f, expf(__fmul_rn(-sig, dist)));
          tt[j] = (double)__fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
        }
      }
      // ---- exclusive cumprod: over the 4 quarters of the ray (lanes rl, rl+8, rl+16, rl+24), then in-lane ----
      double incl = (tt[0] * tt[1]) * (tt[2] * tt[3]);
      {
        double n1 = shfl_up_d(incl, 8);
        if (q >= 1) incl *= n1;
        double n2 = shfl_up_d(incl, 16);
        if (q >= 2) incl *= n2;
      }
      double excl = shfl_up_d(incl, 8);
      double run = carry * (q ? excl : 1.0);
      carry *= shfl_d(incl, 24 + rl);
      float w[4];
      float pr = 0.f, pg = 0.f, pb = 0.f, pd = 0.f, pa = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[j] = __fmul_rn(alpha[j], (float)run);
        run *= tt[j];
        pr = fmaf(w[j], cr[j], pr), pg = fmaf(w[j], cg[j], pg), pb = fmaf(w[j], cb[j], pb);
        pd = fmaf(w[j], zc[j], pd), pa += w[j];
      }
      sr += (double)pr, sg += (double)pg, sb += (double)pb, sd += (double)pd, sa += (double)pa;
      if (a.weights && valid) {
        float* wd = a.weights + ray * S + s0;
        if (vec_w && s0 + 4 <= S) *reinterpret_cast<float4*>(wd) = make_float4(w[0], w[1], w[2], w[3]);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (s0 + j < S) wd[j] = w[j];
        }
      }
      if (RESAMPLE && s0 + 4 <= P) *reinterpret_cast<float4*>(wsm + rl * P + s0) = make_float4(w[0], w[1], w[2], w[3]);
      if (t + 1 < TS) {
        if (RESAMPLE) load_tile<BLOCKED>(a, cur, raw_blk, zrow, t + 1, q, S, S1, valid, vec_z);
        else cur = nxt;
      }
    }
    // ---- per-ray maps: reduce the 4 quarters, quarter 0 writes ----
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      sr += __shfl_xor_sync(kFull, sr, o), sg += __shfl_xor_sync(kFull, sg, o), sb += __shfl_xor_sync(kFull, sb, o);
      sd += __shfl_xor_sync(kFull, sd, o), sa += __shfl_xor_sync(kFull, sa, o);
    }
    if (q == 0 && valid) {
      float accv = (float)sa, depthv = (float)sd;
      float vr = (float)sr, vg = (float)sg, vb = (float)sb;
      // 1/max(1e-10, depth/acc): torch.max propagates NaN (acc == 0 -> 0/0)
      float qd = __fdiv_rn(depthv, accv);
      float m = isnan(qd) ? qd : fmaxf(1e-10f, qd);
      float dispv = __fdiv_rn(1.f, m);
      if (a.white_bkgd) {
        float bg = __fsub_rn(1.f, accv);
        vr = __fadd_rn(vr, bg), vg = __fadd_rn(vg, bg), vb = __fadd_rn(vb, bg);
      }
      a.rgb[ray * 3] = vr, a.rgb[ray * 3 + 1] = vg, a.rgb[ray * 3 + 2] = vb;
      a.disp[ray] = dispv, a.acc[ray] = accv, a.depth[ray] = depthv;
    }
    if (!RESAMPLE) continue;

    // ---- hierarchical resampling, one ray at a time on all 32 lanes:
    //      train_utils.py:144-156 + nerf_helpers.py:668-702 ----
    const unsigned zs_ballot = __ballot_sync(kFull, z_sorted);  // ray r: bits r, r+8, r+16, r+24
    __syncwarp();
    for (int r = 0; r < kBlkRays; ++r) {
      const int64_t rr = blk * kBlkRays + r;
      if (rr >= a.n_rays) break;
      const float* zv = zsm + r * P1;
      const float* wv = wsm + r * P;
      for (int i = lane; i < B; i += 32) {
        float m0 = __fmul_rn(0.5f, __fadd_rn(zv[i + 1], zv[i]));
        if (a.mip) {
          float m1 = __fmul_rn(0.5f, __fadd_rn(zv[i + 2], zv[i + 1]));
          m0 = __fmul_rn(0.5f, __fadd_rn(m1, m0));
        }
        bins[i] = m0;
      }
      for (int k = lane; k < n_words; k += 32) bm[k] = 0u;
      build_cdf(wv + 1, S - 2, cdf, lane);  // weights[...,1:-1]
      __syncwarp();
      bool sorted = ((zs_ballot >> r) & 0x01010101u) == 0x01010101u;  // z_vals non-decreasing
      float prev_last = 0.f;
      for (int j0 = 0; j0 < nf; j0 += 32) {
        const int j = j0 + lane;
        const bool act = j < nf;
        float x = 0.f;
        if (act) {
          const float u = u_regs ? ureg[j0 >> 5] : (a.u_per_ray ? __ldg(a.u + rr * nf + j) : __ldg(a.u + j));
          const int ind = upper_bound_fixed(cdf, B, top, u);
          const int below = max(0, ind - 1), above = min(B - 1, ind);
          const float cb_ = cdf[below], ca_ = cdf[above];
          float denom = __fsub_rn(ca_, cb_);
          if (denom < 1e-5f) denom = 1.f;
          const float tq = __fdiv_rn(__fsub_rn(u, cb_), denom);
          const float bb = bins[below], ba = bins[above];
          x = __fadd_rn(bb, __fmul_rn(tq, __fsub_rn(ba, bb)));
          zs[j] = x;
          if (a.inds) a.inds[rr * nf + j] = ind;
          if (a.z_samples) a.z_samples[rr * nf + j] = x;
          // rank of z_samples[j] in the merge = j + #(z_vals <= z_samples[j]).  The sample lies between the
          // bin centres around the bin it was drawn from, so the count starts at that bin index and is
          // corrected by a couple of probes (exact whenever z_vals is sorted; the guess only sets the probe
          // count).  The position goes into the bitmap; it is used only if both inputs turn out sorted.
          int c = min(max(ind + a.mip, 1), S1);
          while (c < S1 && zv[c] <= x) ++c;
          while (c > 0 && zv[c - 1] > x) --c;
          const int pos = j + c;
          atomicOr(&bm[pos >> 5], 1u << (pos & 31));
        }
        // z_samples non-decreasing?  neighbours live in the next lane / the next round's lane 0
        const float up = __shfl_down_sync(kFull, x, 1);
        if (lane < 31 && j + 1 < nf) sorted &= (x <= up);
        const float first = __shfl_sync(kFull, x, 0);
        if (j0 > 0 && lane == 0) sorted &= (prev_last <= first);
        prev_last = __shfl_sync(kFull, x, 31);
      }
      __syncwarp();
      sorted = __all_sync(kFull, sorted);
      float* out = a.z_merged + rr * (int64_t)M;
      if (sorted) {
        int ones = 0;  // z_samples entries before this word
        for (int k = 0; k < n_words; ++k) {
          const unsigned word = bm[k];
          const int p = (k << 5) + lane;
          const int before = ones + __popc(word & ((1u << lane) - 1u));
          if (p < M) out[p] = ((word >> lane) & 1u) ? zs[before] : zv[p - before];
          ones += __popc(word);
        }
      } else {
        // rare (rounding-induced inversions, random u, NaNs): all-pairs ranking, still exact
        for (int e = lane; e < M; e += 32) {
          float x = e < S1 ? zv[e] : zs[e - S1];
          int rank = 0;
          for (int k = 0; k < M; ++k) {
            float y = k < S1 ? zv[k] : zs[k - S1];
            rank += key_less(y, k, x, e) ? 1 : 0;
          }
          out[rank] = x;
        }
      }
      __syncwarp();  // cdf / bins / zs / bitmap are reused by the next ray
    }
  }
}

// stand-alone sample_pdf: bins [n,B], weights [n,B-1] (or cdf_in [n,B])
__global__ void __launch_bounds__(256)
sample_pdf_kernel(const float* __restrict__ bins_g, const float* __restrict__ w_g, const float* __restrict__ cdf_in,
                  int64_t n_rays, int B, const float* __restrict__ u_g, int u_per_ray, int nf,
                  int64_t* __restrict__ inds, float* __restrict__ samples, float* __restrict__ cdf_out,
                  int warps_per_cta) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* bins = sm + (size_t)wid * (3 * B + 2);
  float* cdf = bins + B;
  float* w = cdf + B + 1;
  for (int64_t ray = (int64_t)blockIdx.x * warps_per_cta + wid; ray < n_rays;
       ray += (int64_t)gridDim.x * warps_per_cta) {
    __syncwarp();
    for (int i = lane; i < B; i += 32) bins[i] = __ldg(bins_g + ray * B + i);
    if (cdf_in) {
      for (int i = lane; i < B; i += 32) cdf[i] = __ldg(cdf_in + ray * B + i);
    } else {
      for (int i = lane; i < B - 1; i += 32) w[i] = __ldg(w_g + ray * (B - 1) + i);
      __syncwarp();
      build_cdf(w, B - 1, cdf, lane);
    }
    __syncwarp();
    if (cdf_out)
      for (int i = lane; i < B; i += 32) cdf_out[ray * B + i] = cdf[i];
    for (int j = lane; j < nf; j += 32) {
      float u = u_per_ray ? __ldg(u_g + ray * nf + j) : __ldg(u_g + j);
      int ind;
      float smp = invert_cdf(cdf, bins, B, u, &ind);
      if (inds) inds[ray * nf + j] = ind;
      if (samples) samples[ray * nf + j] = smp;
    }
  }
}


// out[ray] = sort(cat(a[ray], b[ray])) ascending (train_utils.py:144-156: the merged depths of the fine pass when the
// inverse-CDF samples are NOT sorted, i.e. with perturbation; the det=True frame path rank-merges inside composite_kernel).
// One warp per ray, a bitonic network over 32 * M keys held in registers: element e = reg * 32 + lane, so strides < 32
// are lane exchanges (SHFL.BFLY) and strides >= 32 are register exchanges.  Keys are the order-preserving unsigned image
// of the floats (total order; a NaN sorts last like torch.sort), padding = 0xffffffff.
template <int M>
__global__ void __launch_bounds__(128) sort_cat_kernel(const float* __restrict__ a, int sa, const float* __restrict__ b, int sb,
                                                       int64_t n_rays, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int S = sa + sb;
  uint32_t v[M];
#pragma unroll
  for (int r = 0; r < M; ++r) {
    const int e = r * 32 + lane;
    uint32_t key = 0xffffffffu;
    if (e < S) {
      const uint32_t bits = __float_as_uint(e < sa ? __ldg(a + ray * sa + e) : __ldg(b + ray * sb + (e - sa)));
      key = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);
    }
    v[r] = key;
  }
#pragma unroll
  for (int k = 2; k <= 32 * M; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= 32) {
        const int rj = j >> 5;
#pragma unroll
        for (int r = 0; r < M; ++r) {
          if ((r & rj) == 0) {
            const bool up = (((r * 32) & k) == 0);   // k >= 64 here: the block direction depends on the register index only
            const uint32_t lo = min(v[r], v[r | rj]), hi = max(v[r], v[r | rj]);
            v[r] = up ? lo : hi;
            v[r | rj] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < M; ++r) {
          const int e = r * 32 + lane;
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool up = ((e & k) == 0), lower = ((lane & j) == 0);
          v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < M; ++r) {
    const int e = r * 32 + lane;
    if (e < S) {
      const uint32_t key = v[r];
      out[ray * S + e] = __uint_as_float(key ^ ((key >> 31) ? 0x80000000u : 0xffffffffu));
    }
  }
}

template <typename K>
static int32_t pick_warps(K kernel, size_t floats_per_warp, int max_warps, int* warps_out, size_t* smem_out) {
  size_t per_warp = floats_per_warp * sizeof(float);
  int warps = max_warps;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  if (per_warp * warps > 200 * 1024) return NVSR_ERR_RESOURCE;
  size_t smem = per_warp * warps;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int32_t)e;
  }
  *warps_out = warps;
  *smem_out = smem;
  return NVSR_OK;
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int32_t nvsr_composite(const nvsr_composite_t* c, void* stream) {
  NVSR_CHECK_ARG(c && c->n_rays >= 0 && c->n_samples > 0 && c->n_samples <= NVSR_MAX_SAMPLES);
  NVSR_CHECK_ARG(c->raw && c->z && c->rd && c->rgb && c->disp && c->acc && c->depth);
  if (c->n_fine > 0) {
    NVSR_CHECK_ARG(c->u && c->z_merged && c->n_samples >= 3 && c->n_fine <= NVSR_MAX_SAMPLES);
  }
  if (c->n_rays == 0) return NVSR_OK;
  NVSR_CHECK_ARG(c->row_order == NVSR_ROWS_RAY_MAJOR || c->row_order == NVSR_ROWS_BLOCKED);
  NVSR_CHECK_ARG(c->raw_stride >= rows_padded(c->n_rays, c->n_samples, c->row_order));
  CompositeArgs a{c->n_rays, c->n_samples, c->raw, c->raw_stride, c->z, c->rd, c->noise, c->white_bkgd, c->mip ? 1 : 0,
                  c->rgb, c->disp, c->acc, c->depth, c->weights, c->n_fine > 0 ? c->n_fine : 0, c->u, c->u_per_ray,
                  c->inds, c->z_samples, c->z_merged};
  const bool blocked = c->row_order == NVSR_ROWS_BLOCKED;
  auto kernel = a.n_fine > 0 ? (blocked ? composite_kernel<true, true> : composite_kernel<false, true>)
                             : (blocked ? composite_kernel<true, false> : composite_kernel<false, false>);
  int warps = kCompWarps;
  size_t smem = 0;
  if (a.n_fine > 0) {
    int32_t st = pick_warps(kernel, composite_warp_floats(a.S, a.S + a.mip, a.n_fine), kCompWarps, &warps, &smem);
    if (st != NVSR_OK) return st;
  }
  int64_t ray_blocks = ceil_div64(c->n_rays, kBlkRays);
  int64_t blocks = ceil_div64(ray_blocks, warps);
  int64_t max_blocks = (int64_t)kNumSMs * 32;
  if (blocks > max_blocks) blocks = max_blocks;
  kernel<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(a, warps);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_sample_pdf(const float* bins, const float* weights, const float* cdf_in, int64_t n_rays,
                                   int32_t n_bins, const float* u, int32_t u_per_ray, int32_t n_samples,
                                   int64_t* inds, float* samples, float* cdf_out, void* stream) {
  NVSR_CHECK_ARG(bins && (weights || cdf_in) && u && n_rays >= 0 && n_bins >= 2 && n_bins <= NVSR_MAX_SAMPLES);
  NVSR_CHECK_ARG(n_samples > 0);
  if (n_rays == 0) return NVSR_OK;
  int warps;
  size_t smem;
  int32_t st = pick_warps(sample_pdf_kernel, (size_t)3 * n_bins + 2, 8, &warps, &smem);
  if (st != NVSR_OK) return st;
  int64_t blocks = ceil_div64(n_rays, warps);
  int64_t max_blocks = (int64_t)kNumSMs * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  sample_pdf_kernel<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, cdf_in, n_rays, n_bins, u, u_per_ray, n_samples, inds, samples, cdf_out, warps);
  NVSR_RETURN_LAST_ERROR();
}

extern "C" int32_t nvsr_sort_cat(const float* a, int32_t sa, const float* b, int32_t sb, int64_t n_rays, float* out, void* stream) {
  NVSR_CHECK_ARG(out && sa >= 0 && sb >= 0 && sa + sb > 0 && n_rays >= 0 && (a || sa == 0) && (b || sb == 0));
  if (sa + sb > 512) return NVSR_ERR_UNSUPPORTED;
  if (n_rays == 0) return NVSR_OK;
  const unsigned blocks = (unsigned)ceil_div64(n_rays, 4);
  const int S = sa + sb;
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 64) sort_cat_kernel<2><<<blocks, 128, 0, st>>>(a, sa, b, sb, n_rays, out);
  else if (S <= 128) sort_cat_kernel<4><<<blocks, 128, 0, st>>>(a, sa, b, sb, n_rays, out);
  else if (S <= 256) sort_cat_kernel<8><<<blocks, 128, 0, st>>>(a, sa, b, sb, n_rays, out);
  else sort_cat_kernel<16><<<blocks, 128, 0, st>>>(a, sa, b, sb, n_rays, out);
  NVSR_RETURN_LAST_ERROR();
}
