// Convenience entry of SURVEY.md §8(b): nvsr_render_rays runs the whole coarse -> fine pipeline of
// predict_and_render_radiance (train_utils.py:71-182) for one batch of prepared rays of a tri-plane scene with ONE host
// call — the same stage entry points, in the same order and on the same stream as the Python host side issues them,
// working out of a caller-owned workspace (nvsr_workspace_bytes), so a frame needs no allocation at all:
//   viewdir gather -> [coarse] row bias, sampler + gather, density chain, rgb chain, composite + sample_pdf + merge
//                  -> [fine]   (view features again if the fine model reads another view plane) row bias, gather at
//                              the merged depths, density chain, rgb chain, composite.
// Host code only (no kernel of its own): every launch is a stage kernel that is parity-tested in isolation.
#include "common.cuh"

namespace nvsr {

struct WsLayout {
  int64_t vfeat, rbias, feat_p, feat_m, raw, z_coarse, z_merged, total;
  int64_t raw_stride;
};

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

static int32_t ws_layout(const nvsr_render_t* r, WsLayout* L) {
  if (!r || r->n_rays < 0 || r->n_coarse <= 0 || r->n_fine < 0 || !r->planes_coarse || !r->dec_coarse) return NVSR_ERR_INVALID_ARG;
  if (r->n_fine > 0 && (!r->planes_fine || !r->dec_fine)) return NVSR_ERR_INVALID_ARG;
  const bool f32 = r->precision == NVSR_F32;
  if (!f32 && !is_16bit(r->precision)) return NVSR_ERR_INVALID_ARG;
  const int C = r->planes_coarse->channels;
  const int S_max = r->n_coarse + r->n_fine;
  const int order = f32 ? NVSR_ROWS_RAY_MAJOR : NVSR_ROWS_BLOCKED;
  const int64_t rows = rows_padded(r->n_rays, S_max, order);
  const int64_t e = f32 ? 4 : 2;
  const int n_out0 = r->dec_coarse->rgb[0].n_out;
  int64_t off = 0;
  L->vfeat = off, off = align256(off + r->n_rays * (int64_t)C * 4);
  L->rbias = off, off = align256(off + r->n_rays * (int64_t)n_out0 * 4);
  L->feat_p = off, off = align256(off + rows * 3 * C * e);
  L->feat_m = off, off = align256(off + rows * (int64_t)C * e);
  L->raw_stride = ceil_div64(rows, kTileRows) * kTileRows;
  L->raw = off, off = align256(off + 4 * L->raw_stride * 4);
  L->z_coarse = off, off = align256(off + r->n_rays * (int64_t)r->n_coarse * 4);
  L->z_merged = off, off = align256(off + r->n_rays * (int64_t)S_max * 4);
  L->total = off;
  return NVSR_OK;
}

static int32_t run_pass(const nvsr_render_t* r, const WsLayout& L, uint8_t* ws, bool fine, void* stream) {
  const nvsr_planes_t* pl = fine ? r->planes_fine : r->planes_coarse;
  const nvsr_decoder_t* dec = fine ? r->dec_fine : r->dec_coarse;
  const float* vplane = fine ? r->vplane_fine : r->vplane_coarse;
  const bool f32 = r->precision == NVSR_F32;
  const int order = f32 ? NVSR_ROWS_RAY_MAJOR : NVSR_ROWS_BLOCKED;
  const int layout = f32 ? NVSR_FEAT_ROWMAJOR_F32 : (r->precision == NVSR_F16 ? NVSR_FEAT_TILE_F16 : NVSR_FEAT_TILE_BF16);
  const int S = fine ? r->n_coarse + r->n_fine : r->n_coarse;
  const int C = pl->channels;
  float* vfeat = reinterpret_cast<float*>(ws + L.vfeat);
  float* rbias = reinterpret_cast<float*>(ws + L.rbias);
  float* raw = reinterpret_cast<float*>(ws + L.raw);
  float* z_coarse = reinterpret_cast<float*>(ws + L.z_coarse);
  float* z_merged = reinterpret_cast<float*>(ws + L.z_merged);
  int32_t st;
  // per-ray view features (a5, view half) — shared by both passes when they read the same view plane
  if (!fine || r->vplane_fine != r->vplane_coarse) {
    st = nvsr_viewdir_gather(r->viewdirs, r->n_rays, vplane, r->vrh, r->vrw, C, r->az_lo, r->az_rng, r->el_lo, r->el_rng,
                             vfeat, stream);
    if (st != NVSR_OK) return st;
  }
  st = nvsr_row_bias(vfeat, r->n_rays, C, dec->view_w, dec->view_ldw, dec->view_b, dec->rgb[0].n_out, rbias, stream);
  if (st != NVSR_OK) return st;
  nvsr_sampler_t s;
  s.n_rays = r->n_rays, s.n_samples = S, s.ro = r->ro, s.rd = r->rd, s.near_ = r->near_, s.far_ = r->far_;
  s.lindisp = r->lindisp, s.t_vals = fine ? nullptr : r->t_vals, s.t_rand = fine ? nullptr : r->t_rand;
  s.z_in = fine ? z_merged : nullptr;
  st = nvsr_sample_gather(&s, pl, layout, ws + L.feat_p, ws + L.feat_m, fine ? nullptr : z_coarse, stream);
  if (st != NVSR_OK) return st;
  const int64_t rows = rows_padded(r->n_rays, S, order);
  nvsr_mlp_t m;
  m.precision = r->precision, m.rows = rows, m.samples_per_ray = S, m.n_rays = r->n_rays, m.raw = raw;
  m.raw_stride = L.raw_stride, m.row_order = order, m.row_ids = nullptr, m.row_count = nullptr;
  m.n_layers = dec->n_density, m.in = ws + L.feat_m;
  for (int l = 0; l < dec->n_density; ++l) m.layer[l] = dec->density[l];
  st = nvsr_mlp_chain(&m, stream);
  if (st != NVSR_OK) return st;
  m.n_layers = dec->n_rgb, m.in = ws + L.feat_p;
  for (int l = 0; l < dec->n_rgb; ++l) m.layer[l] = dec->rgb[l];
  m.layer[0].row_bias = rbias;
  st = nvsr_mlp_chain(&m, stream);
  if (st != NVSR_OK) return st;
  nvsr_composite_t c;
  c.n_rays = r->n_rays, c.n_samples = S, c.raw = raw, c.raw_stride = L.raw_stride, c.row_order = order;
  c.z = fine ? z_merged : z_coarse, c.rd = r->rd, c.noise = fine ? r->noise_f : r->noise_c;
  c.white_bkgd = r->white_bkgd, c.mip = 0;
  c.rgb = fine ? r->rgb_f : r->rgb_c, c.disp = fine ? r->disp_f : r->disp_c, c.acc = fine ? r->acc_f : r->acc_c;
  c.depth = fine ? r->depth_f : r->depth_c, c.weights = nullptr;
  c.n_fine = (!fine && r->n_fine > 0) ? r->n_fine : 0, c.u = r->u, c.u_per_ray = r->u_per_ray;
  c.inds = nullptr, c.z_samples = nullptr, c.z_merged = z_merged;
  return nvsr_composite(&c, stream);
}

}  // namespace nvsr

using namespace nvsr;

extern "C" int64_t nvsr_workspace_bytes(const nvsr_render_t* r) {
  WsLayout L;
  if (ws_layout(r, &L) != NVSR_OK) return -1;
  return L.total;
}

extern "C" int32_t nvsr_render_rays(const nvsr_render_t* r, void* stream) {
  WsLayout L;
  int32_t st = ws_layout(r, &L);
  if (st != NVSR_OK) return st;
  NVSR_CHECK_ARG(r->ro && r->rd && r->viewdirs && r->t_vals && r->vplane_coarse && r->workspace);
  NVSR_CHECK_ARG(r->rgb_c && r->disp_c && r->acc_c && r->depth_c);
  NVSR_CHECK_ARG(r->dec_coarse->n_density > 0 && r->dec_coarse->n_density <= NVSR_MAX_LAYERS && r->dec_coarse->n_rgb > 0 &&
                 r->dec_coarse->n_rgb <= NVSR_MAX_LAYERS && r->dec_coarse->view_w && r->dec_coarse->view_b);
  if (r->n_fine > 0) {
    NVSR_CHECK_ARG(r->u && r->vplane_fine && r->rgb_f && r->disp_f && r->acc_f && r->depth_f);
    NVSR_CHECK_ARG(r->planes_fine->channels == r->planes_coarse->channels);
    NVSR_CHECK_ARG(r->dec_fine->n_density > 0 && r->dec_fine->n_density <= NVSR_MAX_LAYERS && r->dec_fine->n_rgb > 0 &&
                   r->dec_fine->n_rgb <= NVSR_MAX_LAYERS && r->dec_fine->view_w && r->dec_fine->view_b);
    NVSR_CHECK_ARG(r->dec_fine->rgb[0].n_out == r->dec_coarse->rgb[0].n_out);
  }
  if (r->workspace_bytes < L.total) return NVSR_ERR_RESOURCE;
  if ((reinterpret_cast<uintptr_t>(r->workspace) & 255u) != 0) return NVSR_ERR_ALIGNMENT;
  if (r->n_rays == 0) return NVSR_OK;
  uint8_t* ws = static_cast<uint8_t*>(r->workspace);
  st = run_pass(r, L, ws, false, stream);
  if (st != NVSR_OK || r->n_fine == 0) return st;
  return run_pass(r, L, ws, true, stream);
}
