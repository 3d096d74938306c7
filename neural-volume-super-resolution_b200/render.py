"""Drop-in render path: same call surface as the reference's train_utils.py, kernels underneath.

    run_one_iter_of_nerf(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode, ...)
    eval_nerf(height, width, focal_length, model_coarse, model_fine, ray_origins, ray_directions, ...)
    install(train_utils_module)     rebinds train_utils.run_one_iter_of_nerf

Per ray chunk the pipeline is 5 launches per pass instead of the reference's few hundred ATen ops:
  viewdir gather -> [coarse] row-bias, sampler+gather, density chain, rgb chain, composite+resample
                 -> [fine]   row-bias, sampler+gather, density chain, rgb chain, composite
Nothing of pts / embedded / hidden activations is materialised in HBM; the only intermediates are
the feature tiles (gather -> decoder) and raw [4, rows] (decoder -> composite).

Engagement rule: forward-only.  Under autograd (training) the kernels are not engaged: `install`
keeps calling the reference's own function there; calling this module's functions directly with
grad enabled raises.  Unsupported model configurations raise NotImplementedError — no fallback.
"""
import os
import weakref

import torch

from . import _lib, ops, scene
from ._lib import NVSR_BF16, NVSR_F16, NVSR_F32

_PRECISION = {"bf16": NVSR_BF16, "fp16": NVSR_F16, "fp32": NVSR_F32, "fp16-split": NVSR_F16}
_state = {
    "precision": _PRECISION[os.environ.get("NVSR_PRECISION", "fp16")],
    # 'fp16-split': fp16 everywhere except the DENSITY chain, which runs on the tensor cores with every operand split
    # into two fp16 terms (three MMA passes per layer, csrc/mlp_split.cu) on features interpolated in fp32 from the planes stored as fp16 hi + lo halves: the 16-bit modes'
    # map error is sigma's (the colour logits are 3e-5 off), so this mode meets the 1e-3 contract on tcgen05
    "split_density": os.environ.get("NVSR_PRECISION", "fp16") == "fp16-split",
    # rays per chunk of the frame loop.  The reference chunks for memory (131 072 points per network call,
    # train_utils.py:228-234); here a chunk only bounds the temporaries (384 B of features per sample row),
    # and 180 GB of HBM take half a frame at once: fewer, longer launches (measured -2.8 % per frame against
    # 32 768-ray chunks).  NVSR_MAX_CHUNK_ROWS caps rays x samples of one chunk (64 Mi rows ~ 25 GB).
    "ray_chunk": int(os.environ.get("NVSR_RAY_CHUNK", "327680")),
    "max_chunk_rows": int(os.environ.get("NVSR_MAX_CHUNK_ROWS", str(64 << 20))),
    # sparse colour path (16-bit modes): the rgb decoder is evaluated only for samples with sigma (+ noise) > 0 —
    # every other sample has alpha = 0 and weight exactly 0, so its colour cannot reach any map (exact, not a
    # tolerance: volume_rendering_utils.py:29-44).  NVSR_SPARSE_RGB=0 evaluates every sample.
    "sparse_rgb": os.environ.get("NVSR_SPARSE_RGB", "1") != "0",
    # dense frames (every sample through both decoders) go through ONE C call per chunk, nvsr_render_rays, working out
    # of a cached workspace (no per-frame allocation); NVSR_ONE_CALL=0 issues the same stage calls from Python instead
    # (bit-identical maps; what the sparse colour path, traces and bench.py's per-kernel event timing use anyway)
    "one_call": os.environ.get("NVSR_ONE_CALL", "1") != "0",
}


def set_precision(name):
    """'bf16' / 'fp16': tcgen05 decoder with 16-bit planes/features/weights/activations and fp32
    accumulation (same tensor-core rate; fp16 rounds 8x finer, bf16 has fp32's range);
    'fp32': SIMT fp32 everywhere — the 1e-3 parity contract;
    'fp16-split': the fp16 mode with the density chain on split operands (tcgen05, three passes per layer, fp32
    features): meets the 1e-3 contract at tensor-core speed (tri-plane model; the mip model runs plain fp16)."""
    _state["precision"] = _PRECISION[name]
    _state["split_density"] = name == "fp16-split"


def get_precision():
    if _state["split_density"]:
        return "fp16-split"
    return {NVSR_BF16: "bf16", NVSR_F16: "fp16", NVSR_F32: "fp32"}[_state["precision"]]


def set_ray_chunk(n):
    _state["ray_chunk"] = int(n)


def set_sparse_rgb(on):
    """Evaluate the rgb decoder only where alpha can be non-zero (exact; default on) or everywhere."""
    _state["sparse_rgb"] = bool(on)


def _is_planes_model(m):
    return hasattr(m, "planes_") or hasattr(m, "coord_projector")


_t_vals_cache = {}


def _t_vals(n, device):
    # torch.linspace has its own rounding rule; take it from torch itself.  Cached per
    # (n, device): an upload from pageable memory every pass would synchronise the host with the stream.
    key = (int(n), str(device))
    t = _t_vals_cache.get(key)
    if t is None:
        t = _t_vals_cache[key] = torch.linspace(0.0, 1.0, n).to(device=device, dtype=torch.float32)
    return t


def _slice(t, i0, i1):
    return None if t is None else t[i0:i1]


_pass_cache = scene._Cache()


def _params_sig(model):
    return tuple((p.data_ptr(), p._version) for p in model.parameters()) + (scene._GENERATION[0],)


def _planes_pass(model, scene_id, precision):
    """Cached _PlanesPass: rebuilt only when a plane tensor or a decoder weight changed."""
    model.set_cur_scene_id(scene_id)
    srcs = [scene._source_plane(model, d) for d in range(4)]
    sig = tuple((id(t),) + scene._Cache.key_of(t) for t in srcs) + _params_sig(model) + \
        (getattr(model, "proj_combination", "avg"),)
    per_key = _pass_cache.get(model, sig, dict)
    split = _state["split_density"] and precision == NVSR_F16
    key = (scene_id, precision, split)
    hit = per_key.get(key)
    # the source planes are held weakly: a recycled id()/data_ptr() of a dead tensor must not match
    if hit is None or any(r() is not t for r, t in zip(hit[0], srcs)):
        hit = ([weakref.ref(t) for t in srcs], _PlanesPass(model, scene_id, precision, split))
        per_key[key] = hit
    return hit[1]


def _mip_pass(model, precision):
    per_key = _pass_cache.get(model, _params_sig(model), dict)
    if precision not in per_key:
        per_key[precision] = _MipPass(model, precision)
    return per_key[precision]


class _PlanesPass:
    """Everything one (model, scene) pair needs on the device, packed once and cached."""

    def __init__(self, model, scene_id, precision, split=False):
        scene.check_supported_planes_model(model)
        self.precision = precision
        self.layout = ops.FEAT_LAYOUT[precision]
        self.rows = ops.LAYOUT_ROWS[self.layout]
        self.dec = scene.pack_planes_decoder(model, precision)
        self.split = None
        if split:   # 'fp16-split': the planes as fp16 hi + lo halves (density features) + split weights of the density chain
            self.planes, self.planes_lo = scene.pack_scene_planes_hilo(model, scene_id)
            self.split = self.dec.density_split(model)
        else:
            self.planes = scene.pack_scene_planes(model, scene_id, precision)

    # The sparse colour path pays while few samples are lit (15 % on the bench scene); on a volume that is dense
    # almost everywhere the second gather would cost more than the skipped rgb rows save.  The lit fraction of the
    # last sparse pass comes back asynchronously (pinned 4-byte copy + event, never a host sync); above
    # `_DENSE_ABOVE` the pass runs dense, and every 16th pass probes the sparse path again.  Both paths give
    # bit-identical maps, so the switch is invisible in the results.
    _DENSE_ABOVE = 0.6

    def _sparse_pays(self):
        st = self.__dict__.setdefault("_lit", {"frac": None, "pending": None, "skipped": 0})
        if st["pending"] is not None and st["pending"][1].query():
            buf, _, total = st["pending"]
            st["frac"], st["pending"] = float(buf.item()) / max(total, 1), None
        if st["frac"] is not None and st["frac"] > self._DENSE_ABOVE:
            st["skipped"] += 1
            if st["skipped"] % 16:
                return False
        return True

    def _note_lit(self, count, total):
        st = self.__dict__.setdefault("_lit", {"frac": None, "pending": None, "skipped": 0})
        if st["pending"] is None:
            buf = torch.empty((1,), dtype=torch.int32, pin_memory=True)
            buf.copy_(count, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            st["pending"] = (buf, ev, total)

    def radiance(self, ro, rd, vfeat, near, far, lindisp, S, t_vals=None, z_in=None, t_rand=None, noise=None):
        """-> (raw planar [4,stride], z [n,S]).  `noise`: what the compositing will add to sigma (decides which
        samples can contribute on the sparse colour path)."""
        n = ro.shape[0]
        rows = ops.rows_padded(n, S, self.rows)
        sparse = _state["sparse_rgb"] and self.precision != NVSR_F32 and self._sparse_pays()
        rbias = ops.row_bias(vfeat, self.dec.view_w, self.dec.view_b)
        raw = ops.raw_buffer(n, S, self.rows, ro.device)
        if self.split is not None:
            # one gather for both chains: fp16 3-plane features from the planes' high halves, fp32 combined features from
            # hi + lo (an fp32 tile image the split chain reads directly)
            fp, fm32, z = ops.sample_gather_hilo(ro, rd, near, far, self.planes, self.planes_lo, t_vals=t_vals, z_in=z_in,
                                                 t_rand=t_rand, lindisp=lindisp, density_only=sparse)
            ops.mlp_chain_split(fm32, *self.split, 3, n, S, raw)
        else:
            fp, fm, z = ops.sample_gather(ro, rd, near, far, self.planes, self.layout, t_vals=t_vals, z_in=z_in,
                                          t_rand=t_rand, lindisp=lindisp, density_only=sparse)
            ops.mlp_chain(fm, self.dec.density, rows, raw, self.precision, S, n, self.rows)
        if sparse:
            keep, count = ops.keep_rows(raw, n, S, noise)
            fp = ops.sample_gather_rows(ro, rd, self.planes, self.layout, z, keep, count)
            ops.mlp_chain(fp, self.dec.rgb_chain(rbias), rows, raw, self.precision, S, n, self.rows, row_ids=keep,
                          row_count=count)
            self._note_lit(count, n * S)
        else:
            ops.mlp_chain(fp, self.dec.rgb_chain(rbias), rows, raw, self.precision, S, n, self.rows)
        return raw, z


def _render_planes_chunk(pc, pf, ro, rd, vd, near, far, cfg, randoms, trace, coarse_only=False):
    """`coarse_only`: stop after the coarse pass (maps + merged depths in the returned dict) — the frozen-decoder
    training path of nvsr_b200.autograd renders its gradient-free coarse pass through these kernels."""
    n = ro.shape[0]
    dev = ro.device
    Nc, Nf = cfg.num_coarse, cfg.num_fine
    t_rand = randoms.get("t_rand") if cfg.perturb else None
    if cfg.perturb and t_rand is None:
        t_rand = torch.rand([n, Nc]).to(dev)  # CPU RNG like train_utils.py:108
    noise_c = _noise(randoms.get("noise_c"), cfg, n, Nc, dev)
    sparse_on = _state["sparse_rgb"] and pc.precision != NVSR_F32
    if (_state["one_call"] and trace is None and not coarse_only and not sparse_on and "z_fine" not in randoms
            and ops.PROFILE is None and (Nf == 0 or pf is not None) and pc.split is None):
        u = None
        if Nf > 0:
            u = randoms.get("u")
            if u is None:
                u = _t_vals(Nf, dev) if cfg.perturb == 0.0 else torch.rand([n, Nf]).to(dev)
        noise_f = _noise(randoms.get("noise_f"), cfg, n, Nc + Nf, dev) if Nf > 0 else None
        return ops.render_rays(ro, rd, vd, near, far, pc.planes, pc.dec, pf.planes if pf else None, pf.dec if pf else None,
                               pc.precision, Nc, Nf, _t_vals(Nc, dev), u=u, t_rand=t_rand, noise_c=noise_c, noise_f=noise_f,
                               lindisp=cfg.lindisp, white_background=cfg.white_background)
    vfeat = ops.viewdir_gather(vd, pc.planes)
    raw, z = pc.radiance(ro, rd, vfeat, near, far, cfg.lindisp, Nc, t_vals=_t_vals(Nc, dev), t_rand=t_rand, noise=noise_c)
    u = None
    if Nf > 0:
        u = randoms.get("u")
        if u is None:
            u = _t_vals(Nf, dev) if cfg.perturb == 0.0 else torch.rand([n, Nf]).to(dev)
    co = ops.composite(raw, z, rd, Nc, noise=noise_c, white_background=cfg.white_background, n_fine=Nf, u=u,
                       want_weights=trace is not None, want_inds=trace is not None, want_samples=trace is not None,
                       row_order=pc.rows)
    if trace is not None:
        trace.update(z_coarse=z, raw_coarse=ops.raw_to_nsc(raw, n, Nc, pc.rows), weights_coarse=co["weights"],
                     depth_coarse=co["depth"])
    fo = None
    if Nf > 0 and not coarse_only:
        zf = randoms["z_fine"] if "z_fine" in randoms else co["z_merged"]   # test hook: teacher-forced depths
        # fine model may read different (super-resolved) planes but shares the view-direction plane
        vfeat_f = vfeat if pf.planes.vplane is pc.planes.vplane else ops.viewdir_gather(vd, pf.planes)
        noise_f = _noise(randoms.get("noise_f"), cfg, n, Nc + Nf, dev)
        raw_f, _ = pf.radiance(ro, rd, vfeat_f, near, far, cfg.lindisp, Nc + Nf, z_in=zf, noise=noise_f)
        fo = ops.composite(raw_f, zf, rd, Nc + Nf, noise=noise_f, white_background=cfg.white_background,
                           row_order=pf.rows)
        if trace is not None:
            trace.update(inds=co["inds"], z_samples=co["z_samples"], z_fine=zf,
                         raw_fine=ops.raw_to_nsc(raw_f, n, Nc + Nf, pf.rows), depth_fine=fo["depth"])
    return co, fo


def _noise(given, cfg, n, S, dev):
    std = cfg.radiance_field_noise_std
    if not std or std <= 0.0:
        return None
    if given is None:
        given = torch.randn([n, S])  # CPU RNG like volume_rendering_utils.py:32
    return (given * std).to(device=dev, dtype=torch.float32)


class _MipPass:
    def __init__(self, model, precision):
        self.precision = precision
        self.layout = ops.FEAT_LAYOUT[precision]
        self.dec = scene.pack_mip_decoder(model, precision)

    def radiance(self, z_edges, ro, rd, denc, radius, n_freqs):
        n, s1 = z_edges.shape
        S = s1 - 1
        rows = n * S
        enc = ops.ipe(z_edges, ro, rd, radius, n_freqs, self.layout, self.dec.k0 if self.precision != NVSR_F32 else None)
        rbias = ops.row_bias(denc, self.dec.dir_w, self.dec.dir_b)
        raw = ops.raw_buffer(n, S, ops.ROWS_RAY_MAJOR, ro.device)   # IPE rows are ray-major in every precision
        ops.mlp_chain(enc, self.dec.chain(rbias), rows, raw, self.precision, S, n)
        return raw


def _render_mip_chunk(mc, mf, ro, rd, vd, near, far, cfg, radius, n_freqs, n_dir_freqs, randoms, trace):
    n, dev = ro.shape[0], ro.device
    Nc, Nf = cfg.num_coarse, cfg.num_fine
    # z edges [n, Nc+1] (train_utils.py:95-109).  Elementwise on [n,65]: plumbing, kept in torch so the
    # rounding is torch's own.
    t = _t_vals(Nc + 1, dev)
    nr = torch.full((n, 1), float(near), device=dev)
    fr = torch.full((n, 1), float(far), device=dev)
    z = nr * (1.0 - t) + fr * t if not cfg.lindisp else 1.0 / (1.0 / nr * (1.0 - t) + 1.0 / fr * t)
    if cfg.perturb:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper, lower = torch.cat((mids, z[..., -1:]), -1), torch.cat((z[..., :1], mids), -1)
        t_rand = randoms.get("t_rand")
        t_rand = torch.rand(z.shape).to(dev) if t_rand is None else t_rand
        z = lower + (upper - lower) * t_rand
    z = z.contiguous()
    denc = ops.dir_encoding(vd, n_dir_freqs, True)
    raw = mc.radiance(z, ro, rd, denc, radius, n_freqs)
    u = None
    if Nf > 0:
        u = randoms.get("u")
        if u is None:
            u = _t_vals(Nf + 1, dev) if cfg.perturb == 0.0 else torch.rand([n, Nf + 1]).to(dev)
    co = ops.composite(raw, z, rd, Nc, noise=_noise(randoms.get("noise_c"), cfg, n, Nc, dev),
                       white_background=cfg.white_background, mip=True, n_fine=(Nf + 1 if Nf > 0 else 0), u=u,
                       want_weights=trace is not None, want_inds=trace is not None, want_samples=trace is not None)
    if trace is not None:
        trace.update(z_coarse=z, raw_coarse=raw[:, :n * Nc].t().reshape(n, Nc, 4), weights_coarse=co["weights"],
                     depth_coarse=co["depth"])
    fo = None
    if Nf > 0:
        zf = co["z_merged"]
        Sf = zf.shape[1] - 1
        raw_f = mf.radiance(zf, ro, rd, denc, radius, n_freqs)
        fo = ops.composite(raw_f, zf, rd, Sf, noise=_noise(randoms.get("noise_f"), cfg, n, Sf, dev),
                           white_background=cfg.white_background, mip=True)
        if trace is not None:
            trace.update(inds=co["inds"], z_samples=co["z_samples"], z_fine=zf,
                         raw_fine=raw_f[:, :n * Sf].t().reshape(n, Sf, 4), depth_fine=fo["depth"])
    return co, fo


def run_one_iter_of_nerf(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode="train",
                         encode_position_fn=None, encode_direction_fn=None, scene_config={}, randoms=None,
                         trace=None):
    """Drop-in for train_utils.run_one_iter_of_nerf (train_utils.py:185-282): same arguments, same
    9-tuple (rgb_coarse, disp_coarse, acc_coarse, rgb_fine, disp_fine, acc_fine, None, None, None).

    Extra keyword-only extensions (not in the reference): `randoms` supplies the uniform/normal draws
    the reference takes from the CPU RNG (keys t_rand [N,Nc], u [N,Nf], noise_c, noise_f), `trace`
    (a dict) receives per-stage tensors for parity tests."""
    if torch.is_grad_enabled():
        raise RuntimeError("nvsr_b200.run_one_iter_of_nerf is forward-only: call it under torch.no_grad() "
                           "(training keeps the reference's autograd path; see install())")
    precision = _state["precision"]
    cfg = getattr(options.nerf, mode)
    mip = getattr(options.nerf, "encode_position_fn", None) == "mip"
    if not options.nerf.use_viewdirs:
        raise NotImplementedError("nvsr_b200: use_viewdirs=False is not supported")
    ro_in, rd_in = batch_rays[0], batch_rays[1]
    if not ro_in.is_cuda:
        raise _lib.NvsrError("batch_rays must be CUDA tensors: nvsr_b200 has no CPU path")
    use_ndc = scene_config.no_ndc is False
    ro, rd, vd = ops.prepare_rays(ro_in, rd_in, use_ndc, H, W, focal if use_ndc else 1.0, 1.0)
    n_total = ro.shape[0]
    near, far = float(scene_config.near), float(scene_config.far)
    randoms = randoms or {}

    if _is_planes_model(model_coarse):
        if mip:
            raise NotImplementedError("nvsr_b200: IPE feeds FlexibleNeRFModel only (train_nerf.py:290-291)")
        model_coarse.set_cur_scene_id(scene_id)
        model_fine.set_cur_scene_id(scene_id)
        pc = _planes_pass(model_coarse, scene_id, precision)
        pf = _planes_pass(model_fine, scene_id, precision) if cfg.num_fine > 0 else None
        runner = lambda a, b, c, rnd, tr: _render_planes_chunk(pc, pf, a, b, c, near, far, cfg, rnd, tr)
    else:
        if not mip:
            raise NotImplementedError("nvsr_b200: FlexibleNeRFModel is supported with the mip/IPE encoding only")
        n_freqs = getattr(encode_position_fn, "max_freq", None)
        if n_freqs is None:
            raise NotImplementedError("nvsr_b200: encode_position_fn must be an IntegratedPositionalEncoding")
        if model_coarse.dim_xyz != 6 * n_freqs:
            raise _lib.NvsrError("IPE width does not match the model's dim_xyz")
        n_dir = (model_coarse.dim_dir - 3) // 6
        radius = ops.mip_radius(scene_id)
        mc = _mip_pass(model_coarse, precision)
        mf = _mip_pass(model_fine, precision) if cfg.num_fine > 0 else None
        runner = lambda a, b, c, rnd, tr: _render_mip_chunk(mc, mf, a, b, c, near, far, cfg, radius, n_freqs, n_dir,
                                                            rnd, tr)

    # equal-sized chunks (multiples of the 8-ray block), bounded in rays and in rays x samples
    chunk = max(8, min(_state["ray_chunk"], _state["max_chunk_rows"] // max(1, cfg.num_coarse + cfg.num_fine)))
    n_chunks = max(1, -(-n_total // chunk))
    chunk = max(8, -(-(-(-n_total // n_chunks)) // 8) * 8)
    outs_c, outs_f, traces = [], [], []
    for i0 in range(0, n_total, chunk):
        i1 = min(n_total, i0 + chunk)
        rnd = {k: (v[i0:i1] if (torch.is_tensor(v) and v.dim() == 2 and v.shape[0] == n_total) else v)
               for k, v in randoms.items()}
        rnd = {k: (v.to(ro.device) if torch.is_tensor(v) else v) for k, v in rnd.items()}
        tr = {} if trace is not None else None
        co, fo = runner(ro[i0:i1], rd[i0:i1], vd[i0:i1], rnd, tr)
        outs_c.append(co)
        outs_f.append(fo)
        if tr is not None:
            traces.append(tr)
    if trace is not None and traces:
        for k in traces[0]:
            trace[k] = torch.cat([t[k] for t in traces], 0)

    def cat(outs, key):
        if outs[0] is None:
            return None
        return outs[0][key] if len(outs) == 1 else torch.cat([o[key] for o in outs], 0)

    return (cat(outs_c, "rgb"), cat(outs_c, "disp"), cat(outs_c, "acc"),
            cat(outs_f, "rgb"), cat(outs_f, "disp"), cat(outs_f, "acc"), None, None, None)


def planes_model_forward(model, x, scene_id=None):
    """Drop-in for `TwoDimPlanesModel.forward(x[n,6]) -> [n,4]` (models.py:381-421; what `run_network` calls as
    `network_fn(batch)`, train_utils.py:56): x = (point xyz, view direction) per row, output `cat(rgb, alpha)` raw
    (sigmoid / relu are applied by the compositing).  Same stage kernels as the fused path — gather, per-row view bias,
    both decoder chains — in the precision of `set_precision()`; each point is evaluated as a one-sample ray
    (origin = the point, direction 0), so the 16-bit tile layout carries 15 padding rows per point: a parity /
    integration surface, not the fast path (`run_one_iter_of_nerf` never materialises x)."""
    if torch.is_grad_enabled():
        raise RuntimeError("nvsr_b200.planes_model_forward is forward-only: call it under torch.no_grad()")
    if not x.is_cuda:
        raise _lib.NvsrError("x must be a CUDA tensor: nvsr_b200 has no CPU path")
    sid = scene_id if scene_id is not None else model.cur_id
    pp = _planes_pass(model, sid, _state["precision"])
    x = x.float()
    n = x.shape[0]
    pts, dirs = x[:, :3].contiguous(), x[:, 3:6].contiguous()
    vfeat = ops.viewdir_gather(dirs, pp.planes)
    zero_d = torch.zeros_like(pts)
    z0 = torch.zeros((n, 1), dtype=torch.float32, device=x.device)
    sparse, _state["sparse_rgb"] = _state["sparse_rgb"], False       # every row's colour is wanted here
    try:
        raw, _ = pp.radiance(pts, zero_d, vfeat, 0.0, 1.0, False, 1, z_in=z0)
    finally:
        _state["sparse_rgb"] = sparse
    return ops.raw_to_nsc(raw, n, 1, pp.rows).reshape(n, 4).contiguous()


def eval_nerf(height, width, focal_length, model_coarse, model_fine, ray_origins, ray_directions, options, scene_id,
              mode="validation", encode_position_fn=None, encode_direction_fn=None, scene_config={}):
    """Drop-in for train_utils.eval_nerf (train_utils.py:285-331): full-image synthesis; like the
    reference it returns only the rgb images (disp/acc slots are None)."""
    batch = torch.stack((ray_origins.reshape(-1, 3), ray_directions.reshape(-1, 3)), 0)
    rgb_c, _, _, rgb_f, _, _, _, _, _ = run_one_iter_of_nerf(
        height, width, focal_length, model_coarse, model_fine, batch, options, scene_id, mode="validation",
        encode_position_fn=encode_position_fn, encode_direction_fn=encode_direction_fn, scene_config=scene_config)
    rgb_c = rgb_c.reshape([height, width, -1])
    if rgb_f is not None:
        rgb_f = rgb_f.reshape([height, width, -1])
    return rgb_c, None, None, rgb_f, None, None, None, None, None


def render_frame(height, width, focal, pose, model_coarse, model_fine, options, scene_id, scene_config,
                 downsampling_offset=0.0, row_range=None, encode_position_fn=None, encode_direction_fn=None):
    """get_ray_bundle + run_one_iter_of_nerf for image rows `row_range` (default: all).  Returns the
    9-tuple for that row band (ray order = row-major within the band)."""
    ro, rd = ops.get_ray_bundle(height, width, focal, pose, 0, downsampling_offset, row_range=row_range)
    batch = torch.stack((ro.reshape(-1, 3), rd.reshape(-1, 3)), 0)
    return run_one_iter_of_nerf(height, width, focal, model_coarse, model_fine, batch, options, scene_id,
                                mode="validation", encode_position_fn=encode_position_fn,
                                encode_direction_fn=encode_direction_fn, scene_config=scene_config)


_installed = {}   # module -> {attribute name: original object} for everything install() rebound


def _rebind(module, name, new):
    _installed.setdefault(module, {}).setdefault(name, getattr(module, name))
    setattr(module, name, new)


def install(train_utils_module, nerf_helpers_module=None, train_nerf_module=None, differentiable=False, precision=None):
    """Rebind the reference's seams: `train_utils.run_one_iter_of_nerf` (which
    `eval_nerf` resolves through module globals at call time, train_utils.py:311) and, optionally,
    `nerf_helpers.get_ray_bundle`.  Under autograd the original functions keep running.

    `train_nerf_module`: also rebind the name `train()` itself uses (`from train_utils import run_one_iter_of_nerf`,
    train_nerf.py:13, resolved in train_nerf's globals at train_nerf.py:860).  With `differentiable=True` a
    grad-enabled call then goes to `nvsr_b200.autograd.run_one_iter_of_nerf` (hand-written backward of the gather
    and compositing stages) instead of the reference's function; unsupported configurations raise, there is no
    silent fallback.  Default: training stays on the reference's own autograd path.

    `precision`: 'fp32' | 'fp16' | 'bf16' selects the arithmetic of every later call (same as set_precision); the
    mode in force is logged once here, because the 16-bit modes are a stated-tolerance contract, not the 1e-3 one
    (DESIGN.md §2).  Calling install() again updates the options of the existing wrapper (and rebinds whatever
    modules the second call names); `uninstall()` restores every name any install() call rebound."""
    if precision is not None:
        set_precision(precision)
    cur = train_utils_module.run_one_iter_of_nerf
    if getattr(cur, "_nvsr_b200", False):
        wrapper = cur
        wrapper._options["differentiable"] = bool(differentiable)
    else:
        orig = cur
        options = {"differentiable": bool(differentiable)}

        def run_one_iter_of_nerf_b200(*args, **kwargs):
            if torch.is_grad_enabled():
                if options["differentiable"]:
                    from . import autograd
                    return autograd.run_one_iter_of_nerf(*args, **kwargs)
                return orig(*args, **kwargs)
            return run_one_iter_of_nerf(*args, **kwargs)

        run_one_iter_of_nerf_b200._nvsr_b200 = True
        run_one_iter_of_nerf_b200._options = options
        run_one_iter_of_nerf_b200._orig = orig
        wrapper = run_one_iter_of_nerf_b200
        _rebind(train_utils_module, "run_one_iter_of_nerf", wrapper)
    if train_nerf_module is not None:
        bound = getattr(train_nerf_module, "run_one_iter_of_nerf", None)
        if bound is wrapper._orig:
            _rebind(train_nerf_module, "run_one_iter_of_nerf", wrapper)
    if nerf_helpers_module is not None and not getattr(nerf_helpers_module.get_ray_bundle, "_nvsr_b200", False):
        orig_grb = nerf_helpers_module.get_ray_bundle

        def get_ray_bundle_b200(height, width, focal_length, tform_cam2world, padding_size=0, downsampling_offset=0):
            if not tform_cam2world.is_cuda:
                return orig_grb(height, width, focal_length, tform_cam2world, padding_size, downsampling_offset)
            return ops.get_ray_bundle(height, width, focal_length, tform_cam2world, padding_size, downsampling_offset)

        get_ray_bundle_b200._nvsr_b200 = True
        _rebind(nerf_helpers_module, "get_ray_bundle", get_ray_bundle_b200)
    import logging
    logging.getLogger("nvsr_b200").info(
        "nvsr_b200 installed: precision=%s (%s), differentiable=%s", get_precision(),
        "1e-3 parity contract" if get_precision() == "fp32" else "16-bit operands, stated tolerance - see DESIGN.md section 2",
        wrapper._options["differentiable"])
    return wrapper


def uninstall(train_utils_module=None):
    """Restore every name install() rebound — in `train_utils_module` and in every other module (train_nerf,
    nerf_helpers) an install() call touched — and drop the packed-scene caches."""
    for module, names in list(_installed.items()):
        for name, orig in names.items():
            setattr(module, name, orig)
        del _installed[module]
    clear_caches()


def clear_caches():
    """Drop every packed image this module and `scene` hold (planes, decoder weights, per-(model, scene) passes).
    Call it after writing a plane or weight through `.data` (which does not bump the tensor's version counter, so the
    identity+version cache keys cannot see it) and to release the device memory of scenes no longer rendered."""
    _pass_cache.store.clear()
    _t_vals_cache.clear()
    scene._plane_cache.store.clear()
    scene._decoder_cache.store.clear()
