"""Differentiable wrappers of the gather and compositing stages (SURVEY.md §8f rank 1 — first version).

The reference trains through `run_one_iter_of_nerf` with autograd (train_nerf.py:860-916).  This module gives the two
memory-bound stages of that path hand-written forward AND backward kernels as `torch.autograd.Function`s; the decoder
MLP between them stays an ordinary `nn.Module` under torch autograd in this version (its backward is two plain GEMMs
per layer).  z_samples are detached in the reference (train_utils.py:153): nothing flows through sample_pdf.

    feats  = TriPlaneGather.apply(plane0, plane1, plane2, ro, rd, z, geometry)      # a5 (xyz half), fp32
    vfeat  = ViewdirGather.apply(view_plane, viewdirs, geometry)                     # a5 (view half)
    raw    = decoder(feats, vfeat)                                                   # torch (models.py:393-421)
    rgb, disp, acc, weights, depth = volume_render_radiance_field(raw, z, rd, ...)   # a7, differentiable

There is no CPU path: every tensor must live on a CUDA device (`NvsrError` otherwise).
Verified on the CPU against autograd of the oracle (tests/test_backward_bodies.py compiles the kernels' own
per-element source for the host) and on a B200 through the C-ABI (tests/test_gpu_next_rows.py).
"""
import torch

from . import _lib, ops
from .ops import NVSR_F32, FEAT_ROWMAJOR_F32


_state = {"fast_frozen_coarse": False, "decoder": "tc", "loss_scale": 1024.0, "sparse_backward": True, "sparse_forward": True}
_state["device_rng"] = False
_state["precise_density"] = False
_DENSE = "dense"     # planes_model_forward's default for sigma_noise: the caller does not say what the compositing will add


def set_sparse_backward(flag):
    """True (default): the 'tc' decoder's backward runs over the samples whose raw gradient is not identically zero
    only (row list built on the device from d_raw: csrc/train_tc.cu nonzero_rows / compact_rows) — the others have
    alpha = 0 or transmittance 0 and add nothing to any gradient.  False: every sample goes through the backward chains.
    The two give the same gradients up to the order of the fp32 sums."""
    _state["sparse_backward"] = bool(flag)


def set_device_rng(flag):
    """False (default): the stratified offsets, the inverse-CDF u and the density noise are drawn like the reference
    draws them — `torch.rand` / `torch.randn` on the CPU (train_utils.py:108, nerf_helpers.py:683,
    volume_rendering_utils.py:32), so the same torch seed gives the reference's numbers — and uploaded (a pageable copy
    that synchronises the host with the device, ~3 MB per 4 096-ray step).  True: drawn on the device from torch's CUDA
    generator: same distributions, another stream of numbers, no upload, and the step stays capturable into a CUDA graph
    (`GraphedStep`) without the caller providing `randoms`."""
    _state["device_rng"] = bool(flag)


def _rand(shape, device, normal=False):
    dev = device if _state["device_rng"] else None
    return (torch.randn if normal else torch.rand)(tuple(shape), device=dev)


def set_precise_density(flag):
    """False (default).  True: the density of the 'tc' training forward — what decides which samples are lit and what
    every compositing weight is — comes from the split-operand chain on features interpolated from the planes' fp16 hi + lo
    halves (the 'fp16-split' inference mode's kernels, fp32-grade sigma) instead of the fp16 chain; activations and
    gradients stay on the fp16 training kernels.  The fp16 forward moves sigma by up to ~0.1, which flips
    relu(sigma + noise) on ~0.5 % of the samples and is the whole 6-8 % plane-gradient difference to the fp32 mode
    (DESIGN 2.3); this removes the flips for ~0.4 ms per 4 096-ray step.  Needs the sparse forward."""
    _state["precise_density"] = bool(flag)


def _packed16_lo(plane_nchw):
    """fp16 x-pair image of p - fp16(p), cached like `_packed16`"""
    from . import scene
    per = scene._plane_cache.get(plane_nchw, scene._Cache.key_of(plane_nchw), dict)
    if "autograd_f16_lo" not in per:
        p = plane_nchw.detach().float()
        per["autograd_f16_lo"] = ops.pack_plane((p - p.half().float()).contiguous(), ops.NVSR_F16, range_check=_range_check)
    return per["autograd_f16_lo"]


def set_sparse_forward(flag):
    """True (default): when the caller of `planes_model_forward` tells it the density noise the compositing will add
    (`run_one_iter_of_nerf` does), the 'tc' decoder's rgb chain runs — forward and backward — over the samples with
    relu(sigma + noise) > 0 only, exactly as the inference path's sparse colour path does; the other samples have weight
    exactly 0 (volume_rendering_utils.py:29-44), their rgb is returned as 0.  Needs the row-list backward."""
    _state["sparse_forward"] = bool(flag)


def set_decoder(mode):
    """'tc' (default): the tri-plane decoder of the differentiable path — forward, data gradient and weight gradients —
    runs on this package's tcgen05 kernels (fp16 operands, fp32 accumulation, loss-scaled deltas: csrc/train_tc.cu) and
    the gather writes its 16-bit tile images.  'fp32': the model's own nn.Linear layers under torch autograd on fp32
    features — the parity mode of the training path (gradients within 1e-3 of the reference's, tests/golden)."""
    if mode not in ("tc", "fp32"):
        raise ValueError("decoder mode must be 'tc' or 'fp32'")
    _state["decoder"] = mode


def set_loss_scale(scale):
    """Power-of-two factor folded into the fp16 deltas of the 'tc' decoder backward (default 2**10; the weight and plane
    gradients are returned unscaled).  scripts/studies/backward_precision.py: without it the deltas of an mse over
    thousands of rays underflow fp16 (18 % gradient error); 2**10 .. 2**16 give 6e-4."""
    _state["loss_scale"] = float(scale)


def set_fast_frozen_coarse(on):
    """Opt-in: while the decoder is frozen (`model_coarse.optional_no_grad is torch.no_grad`, train_nerf.py:560 — the
    phase in which only the SR model trains) the gradient-free coarse pass runs on the forward kernels in the
    precision of `set_precision()` instead of the fp32 gather + torch decoder.  Off by default: in the 16-bit modes
    the coarse maps and the resampled depths then carry the forward path's stated tolerance."""
    _state["fast_frozen_coarse"] = bool(on)


class Geometry:
    """What the gather needs besides the plane values: box, projection matrices, view-angle box (models.py:261-268,
    :495-497) — `ops.PackedPlanes` without plane images."""

    def __init__(self, box_lo, box_rng, proj, view_lo_rng, combine="avg"):
        self.box_lo, self.box_rng, self.proj, self.view_lo_rng, self.combine = box_lo, box_rng, proj, view_lo_rng, combine

    _cache = {}

    @classmethod
    def of_model(cls, model, scene_id):
        """Cached per (model, scene, box tensor, its version): reading the box and the projection matrices is a
        device -> host copy, i.e. a host synchronisation — once per scene, not once per training step."""
        import weakref
        box_t = model.box_coords[scene_id]
        key = (id(model), scene_id)
        hit = cls._cache.get(key)
        sig = (box_t.data_ptr(), box_t._version, getattr(model, "proj_combination", "avg"))
        if hit is not None and hit[0]() is model and hit[1] == sig:
            return hit[2]
        geom = cls._build(model, scene_id)
        if len(cls._cache) > 256:
            cls._cache.clear()
        cls._cache[key] = (weakref.ref(model), sig, geom)
        return geom

    @classmethod
    def _build(cls, model, scene_id):
        box = model.box_coords[scene_id].detach().double().cpu()
        lo, rng = box[0].float(), (box[1] - box[0]).float()
        rots = model.coord_projector.rot_mats_NON_LEARNED
        proj = [rots[d].detach().float().cpu()[:, 1:].tolist() for d in range(3)]
        return cls(lo[:3].tolist(), rng[:3].tolist(), proj, (float(lo[3]), float(rng[3]), float(lo[4]), float(rng[4])),
                   getattr(model, "proj_combination", "avg"))


def _cl_image(plane_nchw):
    """fp32 channels-last image of a plane, cached per (tensor, version) like the forward path's packed planes: the
    coarse and the fine pass of a step (and every ray chunk) share one transpose; an optimizer step bumps the version."""
    from . import scene
    per = scene._plane_cache.get(plane_nchw, scene._Cache.key_of(plane_nchw), dict)
    if "autograd_cl" not in per:
        per["autograd_cl"] = ops.pack_plane(plane_nchw, NVSR_F32)
    return per["autograd_cl"]


def _packed(planes_nchw, geom, vplane=None):
    imgs = [_cl_image(p) for p in planes_nchw]
    return ops.PackedPlanes(imgs, NVSR_F32, geom.box_lo, geom.box_rng, geom.proj, vplane, geom.view_lo_rng,
                            combine=geom.combine)


class TriPlaneGather(torch.autograd.Function):
    """(plane0, plane1, plane2 [1,C,R,R] fp32, ro [n,3], rd [n,3], z [n,S]) -> featP [n*S,3C], featM [n*S,C]
    (project_xyz + combine_pos_planes('avg'), models.py:289-310,355-361).  Gradients: the three planes only."""

    @staticmethod
    def forward(ctx, p0, p1, p2, ro, rd, z, geom):
        packed = _packed((p0, p1, p2), geom)
        feat_p, feat_m, _ = ops.sample_gather(ro, rd, 0.0, 1.0, packed, FEAT_ROWMAJOR_F32, z_in=z)
        ctx.save_for_backward(ro, rd, z)
        ctx.geom, ctx.shapes = geom, [tuple(p.shape) for p in (p0, p1, p2)]
        ctx.set_materialize_grads(False)
        return feat_p, feat_m

    @staticmethod
    def backward(ctx, g_p, g_m):
        ro, rd, z = ctx.saved_tensors
        dev = ro.device
        acc = [torch.zeros((s[-2], s[-1], s[-3]), dtype=torch.float32, device=dev) for s in ctx.shapes]
        shell = ops.PackedPlanes(acc, NVSR_F32, ctx.geom.box_lo, ctx.geom.box_rng, ctx.geom.proj, None, ctx.geom.view_lo_rng)
        if g_p is not None or g_m is not None:
            # the kernel scatters w * (dP + dM / 3) ('avg'); for 'sum' the combined features' gradient counts fully
            if g_m is not None and ctx.geom.combine == "sum":
                g_m = g_m * 3.0
            ops.sample_gather_bwd(ro, rd, z, shell, g_p, g_m, acc)
        grads = [a.permute(2, 0, 1).reshape(s) for a, s in zip(acc, ctx.shapes)]   # channels-last -> the parameter's NCHW
        return grads[0], grads[1], grads[2], None, None, None, None


class ViewdirGather(torch.autograd.Function):
    """(view plane [1,C,Rv,Rv], viewdirs [n,3]) -> per-ray view features [n,C] (cart2az_el + project_viewdir,
    nerf_helpers.py:492-496, models.py:312-326).  The decoder broadcasts them over a ray's samples; autograd sums."""

    @staticmethod
    def forward(ctx, vplane, viewdirs, geom):
        img = _cl_image(vplane)
        packed = ops.PackedPlanes([img, img, img], NVSR_F32, geom.box_lo, geom.box_rng, geom.proj, img, geom.view_lo_rng)
        ctx.save_for_backward(viewdirs)
        ctx.geom, ctx.shape = geom, tuple(vplane.shape)
        ctx.set_materialize_grads(False)
        return ops.viewdir_gather(viewdirs, packed)

    @staticmethod
    def backward(ctx, g):
        (viewdirs,) = ctx.saved_tensors
        s = ctx.shape
        acc = torch.zeros((s[-2], s[-1], s[-3]), dtype=torch.float32, device=viewdirs.device)
        shell = ops.PackedPlanes([acc, acc, acc], NVSR_F32, ctx.geom.box_lo, ctx.geom.box_rng, ctx.geom.proj, acc,
                                 ctx.geom.view_lo_rng)
        if g is not None:
            ops.viewdir_gather_bwd(viewdirs, shell, g, acc)
        return acc.permute(2, 0, 1).reshape(s), None, None


class _VolumeRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, radiance_field, depth_values, ray_directions, noise_scaled, white_background, mip):
        n, S, _ = radiance_field.shape
        o = ops.composite(ops.raw_to_planar(radiance_field), depth_values, ray_directions, S, noise=noise_scaled,
                          white_background=white_background, mip=mip, want_weights=True)
        ctx.save_for_backward(radiance_field, depth_values, ray_directions, noise_scaled)
        ctx.white, ctx.mip = bool(white_background), bool(mip)
        ctx.mark_non_differentiable(o["disp"])
        ctx.set_materialize_grads(False)
        return o["rgb"], o["disp"], o["acc"], o["weights"], o["depth"]

    @staticmethod
    def backward(ctx, g_rgb, _g_disp, g_acc, g_w, g_depth):
        rf, z, rd, nz = ctx.saved_tensors
        if g_rgb is None:
            g_rgb = torch.zeros((rf.shape[0], 3), dtype=torch.float32, device=rf.device)
        d_rf = ops.composite_bwd(rf, z, rd, g_rgb, d_acc=g_acc, d_depth=g_depth, d_weights=g_w, noise=nz,
                                 white_background=ctx.white, mip=ctx.mip)
        return d_rf, None, None, None, None, None


# fp16 range of the planes / weights packed by the training forward: checked without a host synchronisation
_range_check = ops.DeferredRangeCheck()


def _packed16(plane_nchw):
    """fp16 x-pair image of a plane (the forward path's gather layout), cached per (tensor, version)"""
    from . import scene
    per = scene._plane_cache.get(plane_nchw, scene._Cache.key_of(plane_nchw), dict)
    if "autograd_f16" not in per:
        per["autograd_f16"] = ops.pack_plane(plane_nchw, ops.NVSR_F16, range_check=_range_check)
    return per["autograd_f16"]


def _with_head_ch(layer, ch):
    """a copy of a ChainLayer whose head (if any) writes channel `ch`"""
    import copy
    l = copy.copy(layer)
    if getattr(l, "head_w", None) is not None:
        l.head_ch = ch
    return l


class PlanesRadianceTC(torch.autograd.Function):
    """TwoDimPlanesModel.forward (models.py:381-421) for n rays x S samples entirely on this package's kernels, forward
    AND backward: 16-bit tile-image gather -> training forward of both decoder chains on tcgen05 (activation images
    kept) | backward: data-gradient chains and weight gradients on tcgen05 (forward operand images read MN-major),
    plane gradients by the scatter kernels.  Inputs: 4 planes, then (weight, bias) of density_dec x4, fc_alpha,
    rgb_dec x4, fc_rgb (20 tensors), then ro, rd, z, viewdirs, geometry.  Output radiance_field [n, S, 4]."""

    @staticmethod
    def forward(ctx, p0, p1, p2, pv, *rest):
        params, (ro, rd, z, vd, geom, sigma_noise) = rest[:20], rest[20:]
        sparse_fwd = sigma_noise is not _DENSE and _state["sparse_forward"] and _state["sparse_backward"]
        dW, dB = params[0:8:2], params[1:8:2]
        aW, aB = params[8], params[9]
        cW, cB = params[10:18:2], params[11:18:2]
        rW, rB = params[18], params[19]
        n, S = z.shape
        C3 = 3 * p0.shape[1]
        F16 = ops.NVSR_F16
        packed = ops.PackedPlanes([_packed16(p) for p in (p0, p1, p2)], F16, geom.box_lo, geom.box_rng, geom.proj,
                                  _cl_image(pv), geom.view_lo_rng, combine=geom.combine)
        precise = sparse_fwd and _state["precise_density"]
        fm32 = None
        if precise:
            _, fm32, feat_m, _ = ops.sample_gather_hilo(ro, rd, 0.0, 1.0, packed, [_packed16_lo(p) for p in (p0, p1, p2)], z_in=z,
                                                        density_only=True, want_m16=True)
            feat_p = None
        else:
            feat_p, feat_m, _ = ops.sample_gather(ro, rd, 0.0, 1.0, packed, ops.FEAT_TILE_F16, z_in=z, density_only=sparse_fwd)
        vfeat = ops.viewdir_gather(vd, packed)
        rb = ops.row_bias(vfeat, cW[0].detach()[:, C3:], cB[0])
        f = lambda t: t.detach().float().contiguous()
        packed_w = ops.pack_weights16(list(dW) + [cW[0].detach()[:, :C3]] + list(cW[1:]), F16, range_check=_range_check)
        wd, wc = packed_w[:4], packed_w[4:]
        _range_check.commit()
        Ld = [ops.ChainLayer(wd[i], f(dB[i]), dW[i].shape[1], 128, True, head_w=f(aW) if i == 3 else None,
                             head_b=f(aB) if i == 3 else None, head_ch=3) for i in range(4)]
        Lc = [ops.ChainLayer(wc[i], None if i == 0 else f(cB[i]), C3 if i == 0 else 128, 128, True,
                             row_bias=rb if i == 0 else None, head_w=f(rW) if i == 3 else None,
                             head_b=f(rB) if i == 3 else None, head_ch=0) for i in range(4)]
        rows = ops.rows_padded(n, S, ops.ROWS_BLOCKED)
        raw = ops.raw_buffer(n, S, ops.ROWS_BLOCKED, ro.device)
        keep = count = None
        if sparse_fwd:
            # sparse path: the density chain runs on every sample WITHOUT storing activations (the inference kernel); only
            # the samples whose density (+ the noise the compositing adds) is positive can reach the maps or carry a
            # gradient, so only they are decoded by the training kernels — density again (its activations, ~1/6 of the
            # rows instead of 1 KB per row for all of them) and colour — in LIST order; the others' rgb stays 0
            if precise:
                # sigma at fp32 grade: split-operand chain (3 MMA passes per layer) on the hi + lo features
                lo_w = ops.pack_weights16([w.detach().float() - w.detach().half().float() for w in dW], F16, range_check=_range_check)
                ops.mlp_chain_split(fm32, wd, lo_w, [f(b) for b in dB], f(aW), f(aB), 3, n, S, raw)
            else:
                ops.mlp_chain(feat_m, Ld, rows, raw, F16, S, n, ops.ROWS_BLOCKED)
            raw[:3].zero_()
            keep, count = ops.keep_rows(raw, n, S, sigma_noise)
            (feat_m,), _ = ops.compact_rows([feat_m], None, keep, count)
            scratch = torch.empty_like(raw[3:4])      # the list pass recomputes the listed sigmas: kept out of `raw`
            acts_d = ops.mlp_chain_train(feat_m, [_with_head_ch(l, 0) for l in Ld], rows, scratch, S, n, row_ids=keep, row_count=count)
            feat_p = ops.sample_gather_rows(ro, rd, packed, ops.FEAT_TILE_F16, z, keep, count)
            acts_c = ops.mlp_chain_train(feat_p, Lc, rows, raw, S, n, row_ids=keep, row_count=count)
        else:
            acts_d = ops.mlp_chain_train(feat_m, Ld, rows, raw, S, n)
            acts_c = ops.mlp_chain_train(feat_p, Lc, rows, raw, S, n)
        ctx.fwd_list = sparse_fwd
        extra = (keep, count) if sparse_fwd else ()
        ctx.save_for_backward(ro, rd, z, vd, vfeat, feat_p, feat_m, *acts_d, *acts_c, *wd, *wc, aW, rW, cW[0], *extra)
        ctx.geom, ctx.shapes = geom, [tuple(p.shape) for p in (p0, p1, p2, pv)]
        ctx.scale = _state["loss_scale"]
        ctx.set_materialize_grads(False)
        return ops.raw_to_nsc(raw, n, S, ops.ROWS_BLOCKED).contiguous()

    @staticmethod
    def backward(ctx, d_rf):
        t = ctx.saved_tensors
        ro, rd, z, vd, vfeat, feat_p, feat_m = t[:7]
        acts_d, acts_c, wd, wc = t[7:11], t[11:15], t[15:19], t[19:23]
        aW, rW, cW0 = t[23:26]
        n, S = z.shape
        dev = ro.device
        Cc = ctx.shapes[0][1]
        C3 = 3 * Cc
        if d_rf is None:
            return (None,) * (4 + 20 + 6)
        scale, inv = ctx.scale, 1.0 / ctx.scale
        d_rf = d_rf.float()
        d_raw = ops.nsc_to_planar_blocked(d_rf, n, S)
        rows = ids = count = None
        if ctx.fwd_list:
            # the forward's own list (sigma + noise > 0): a superset of the rows with a non-zero raw gradient, and the
            # order the rgb chain's images are already in
            ids, count = t[26], t[27]
            rows = (ids, count)
        elif _state["sparse_backward"]:
            # the rows that carry a gradient: the data-gradient chains gather them through the list and leave
            # LIST-ordered copies of the forward's images for the weight gradients
            ids, count = ops.nonzero_rows(d_raw)
            rows = (ids, count)
        # every accumulator of the pass — the two chains' weight gradients, the plane gradients, the per-ray sums — comes
        # out of ONE zero-filled buffer (one memset per pass instead of ~15 fill kernels)
        acc_numel = sum((s_[-1] * s_[-2] * s_[-3] + 3) // 4 * 4 for s_ in ctx.shapes) + (n * 128 + 3) // 4 * 4
        pool = torch.zeros((2 * (128 * (Cc + C3 + 6 * 128 + 2 * 16) + 8 * 128) + acc_numel,), dtype=torch.float32, device=dev)
        cursor = [0]

        def z0(*shape):
            k = 1
            for d in shape:
                k *= d
            if cursor[0] + k > pool.numel():
                return torch.zeros(shape, dtype=torch.float32, device=dev)
            v = pool[cursor[0]:cursor[0] + k].view(shape)
            cursor[0] += (k + 3) // 4 * 4       # keep every view 16-byte aligned
            return v
        grads = []
        needs = ctx.needs_input_grad
        # decoder frozen (the phase in which only the SR model / the planes train, train_nerf.py:560): no weight gradients
        want_w = any(needs[4:24])
        # ---- density chain: data gradient, then the weight gradients on the same images
        if rows is None:
            g, dout, d_fm = ops.mlp_dgrad(wd, Cc, aW, 3, d_raw, scale, acts_d, n, S)
        elif ctx.fwd_list:
            g, dout, d_fm, _, _ = ops.mlp_dgrad(wd, Cc, aW, 3, d_raw, scale, acts_d, n, S, row_count=count, row_ids=ids,
                                                acts_listed=True)
        else:
            g, dout, d_fm, acts_d, feat_m = ops.mlp_dgrad(wd, Cc, aW, 3, d_raw, scale, acts_d, n, S, row_count=count, row_ids=ids,
                                                          x0_img=feat_m if want_w else None)
        dens = [None] * 10
        if want_w:
            dws, dbs, dwh = [z0(128, Cc if l == 0 else 128) for l in range(4)], [z0(128) for _ in range(4)], z0(128, 16)
            ops.mlp_wgrad_chain(g, feat_m, Cc, acts_d, dout, inv, dws, dbs, dwh, row_count=count)
            dens = [t for pair in zip(dws, dbs) for t in pair] + [dwh[:, :1].t().contiguous(), d_rf[..., 3].sum().reshape(1)]
        # ---- rgb chain (its first layer's view-feature columns are a per-ray bias in the forward)
        if rows is None:
            g, dout, d_fp = ops.mlp_dgrad(wc, C3, rW, 0, d_raw, scale, acts_c, n, S)
        elif ctx.fwd_list:
            g, dout, d_fp, _, _ = ops.mlp_dgrad(wc, C3, rW, 0, d_raw, scale, acts_c, n, S, row_count=count, row_ids=ids,
                                                acts_listed=True)
        else:
            g, dout, d_fp, acts_c, feat_p = ops.mlp_dgrad(wc, C3, rW, 0, d_raw, scale, acts_c, n, S, row_count=count, row_ids=ids,
                                                          x0_img=feat_p if want_w else None)
        col = [None] * 10
        g0_ray = None                                                              # [n, 128]: per-ray sum of g_0
        if want_w or needs[3]:
            g0_ray = ops.ray_sum(g[0], n, S, inv) if rows is None else ops.ray_sum_rows(g[0], ids, count, n, S, inv, out=z0(n, 128))
        if want_w:
            dws, dbs, dwh = [z0(128, C3 if l == 0 else 128) for l in range(4)], [z0(128) for _ in range(4)], z0(128, 16)
            ops.mlp_wgrad_chain(g, feat_p, C3, acts_c, dout, inv, dws, dbs, dwh, row_count=count)
            dws[0] = torch.cat([dws[0], g0_ray.t() @ vfeat], 1)         # [128, 3C + C]
            col = [t for pair in zip(dws, dbs) for t in pair] + [dwh[:, :3].t().contiguous(), d_rf[..., :3].sum((0, 1))]
        # ---- planes
        acc = [z0(s[-2], s[-1], s[-3]) for s in ctx.shapes[:3]]
        shell = ops.PackedPlanes(acc, NVSR_F32, ctx.geom.box_lo, ctx.geom.box_rng, ctx.geom.proj, None, ctx.geom.view_lo_rng)
        if ctx.geom.combine == "sum":
            d_fm = d_fm * 3.0
        ops.sample_gather_bwd(ro, rd, z, shell, d_fp, d_fm, acc, rows=rows)
        sv = ctx.shapes[3]
        vgrad = None
        if needs[3]:
            d_v = g0_ray @ cW0.detach()[:, C3:].float()                  # [n, C]
            vacc = z0(sv[-2], sv[-1], sv[-3])
            vshell = ops.PackedPlanes([vacc, vacc, vacc], NVSR_F32, ctx.geom.box_lo, ctx.geom.box_rng, ctx.geom.proj, vacc,
                                      ctx.geom.view_lo_rng)
            ops.viewdir_gather_bwd(vd, vshell, d_v, vacc)
            vgrad = vacc.permute(2, 0, 1).reshape(sv)
        pg = [a.permute(2, 0, 1).reshape(s) for a, s in zip(acc, ctx.shapes[:3])] + [vgrad]
        out = pg + dens + col + [None] * 6
        return tuple(o if (o is None or needs[i]) else None for i, o in enumerate(out))


class GraphedStep:
    """One whole training step — forward, backward and (if `step_fn` does it) the optimizer update — captured into a CUDA
    graph and replayed: the step is ~40 kernel launches of a few microseconds each plus torch glue, so eagerly it is
    bound by the host's enqueue time, not by the GPU (scripts/bench_train_step.py).  The path has no host
    synchronisation and takes every size from the tensors' shapes (the sparse backward reads its row count on the
    device), which is what makes it capturable.

    step_fn(): runs the step reading its inputs (ray batch, targets, random draws — CUDA tensors; draw them inside with
    `.uniform_()` / `.normal_()` or `copy_` into them between replays) from tensors the caller keeps alive, and
    returns whatever should be readable after a replay (e.g. the loss tensor).  Gradients: step_fn must start with
    `optimizer.zero_grad(set_to_none=True)` (or `p.grad = None`), so that the capture allocates them in the graph's pool
    and a replay overwrites instead of accumulating on top of the warm-up steps' gradients.
    Use `capturable=True` optimizers.  `check_every`: replays between host reads of the fp16 range flag."""

    def __init__(self, step_fn, warmup=3, check_every=256):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        _range_check.flush()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = step_fn()
        self.replays, self.check_every = 0, check_every

    def __call__(self):
        self.graph.replay()
        self.replays += 1
        # the replay may have updated parameters (an optimizer step inside the graph) without any `_version` moving:
        # every packed-plane / packed-weight cache entry made before it is stale from here on
        from . import scene
        scene.bump_generation()
        if self.check_every and self.replays % self.check_every == 0:
            _range_check.check_graph()
        return self.result


def _tc_supported(model):
    try:
        d, c = list(model.density_dec["0"]), list(model.rgb_dec["0"])
    except (KeyError, AttributeError, TypeError):
        return False
    C = next(iter(model.planes_.values())).shape[1] if len(model.planes_) else 0
    ok = len(d) == 4 and len(c) == 4 and all(l.out_features == 128 and l.bias is not None for l in d + c)
    return ok and C % 16 == 0 and d[0].in_features == C and c[0].in_features == 4 * C and 3 * C <= 144 \
        and model.fc_alpha["0"].out_features == 1 and model.fc_rgb["0"].out_features == 3


def _render(radiance_field, depth_values, ray_directions, noise_std, white_background, noise, mip=False, nz=None):
    """nz: the density noise ALREADY scaled by the std (what `planes_model_forward(sigma_noise=)` was given)"""
    if nz is None and noise_std > 0.0:
        if noise is None:
            noise = _rand(radiance_field[..., 3].shape, radiance_field.device, normal=True)
        nz = (noise * noise_std).to(radiance_field).contiguous()
    return _VolumeRender.apply(radiance_field.float().contiguous(), depth_values.float().contiguous(),
                               ray_directions.float().contiguous(), nz, white_background, mip)


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, mip_nerf=False, noise=None):
    """Differentiable drop-in for volume_rendering_utils.volume_render_radiance_field (:6-51): same signature and
    5-tuple as `ops.volume_render_radiance_field`, gradient w.r.t. `radiance_field` (disp_map is not differentiated:
    no loss of the reference uses it).  `noise`: the CPU randn draw of :32 ([N,S], unscaled) for parity."""
    if not radiance_field.is_cuda:
        raise _lib.NvsrError("radiance_field must be a CUDA tensor: nvsr_b200 has no CPU path")
    return _render(radiance_field, depth_values, ray_directions, radiance_field_noise_std, white_background, noise, mip_nerf)


def planes_model_forward(model, scene_id, ro, rd, z, viewdirs, sigma_noise=_DENSE):
    """TwoDimPlanesModel.forward (models.py:381-421) for the points ro + rd*z of n rays x S samples with the gather
    done by the Functions above and the decoder by the model's own nn.Linear layers under torch autograd.
    Returns radiance_field [n,S,4].  Gradients reach the model's planes_ parameters and decoder weights.
    sigma_noise (the 'tc' decoder only): what the compositing will add to the density before its relu — None (nothing) or
    the [n,S] noise already scaled by radiance_field_noise_std.  When given, the rgb chain runs over the samples with
    relu(sigma + noise) > 0 only (`set_sparse_forward`) and the rgb of the others — weight exactly 0 in
    volume_render_radiance_field — is returned as 0 instead of the decoder's value.  Default: every sample is decoded.
    Not reproduced: the region-of-interest SR of a TRAINING SR model (models.py:277-280 crops the SR input to the
    footprint of the batch; here the whole plane is super-resolved, which differs near the crop borders)."""
    from . import scene
    scene.check_supported_planes_model(model)
    if getattr(model, "plane_stats", False) and model.training:
        raise NotImplementedError("nvsr_b200.autograd: plane_stats coverage bookkeeping (models.py:304-305) is not reproduced")
    model.set_cur_scene_id(scene_id)
    geom = Geometry.of_model(model, scene_id)
    # models.py:296-310: a position plane is read through the SR model when the scene is an SR scene of this model —
    # the gather's gradient then flows on into the SR network and the LR planes through torch autograd; the
    # view-direction plane is never super-resolved (models.py:312-326)
    planes = [model.planes(d, super_resolve=scene._should_sr(model, d)) for d in range(3)] + [model.planes(3, super_resolve=False)]
    n, S = z.shape
    if _state["decoder"] == "tc" and _tc_supported(model) and geom.combine in ("avg", "sum"):
        params = []
        for seq in (model.density_dec["0"], [model.fc_alpha["0"]], model.rgb_dec["0"], [model.fc_rgb["0"]]):
            for lin in seq:
                params += [lin.weight, lin.bias]
        if sigma_noise is not _DENSE and sigma_noise is not None:
            sigma_noise = sigma_noise.to(device=ro.device, dtype=torch.float32).contiguous()
        return PlanesRadianceTC.apply(planes[0], planes[1], planes[2], planes[3], *params, ro.contiguous(), rd.contiguous(),
                                      z.contiguous(), viewdirs.contiguous(), geom, sigma_noise)
    feat_p, feat_m = TriPlaneGather.apply(planes[0], planes[1], planes[2], ro, rd, z, geom)
    vfeat = ViewdirGather.apply(planes[3], viewdirs, geom)
    h = feat_m
    for lin in model.density_dec["0"]:
        h = torch.relu(lin(h))
    alpha = model.fc_alpha["0"](h)
    h = torch.cat([feat_p, vfeat[:, None, :].expand(n, S, vfeat.shape[-1]).reshape(n * S, -1)], 1)
    for lin in model.rgb_dec["0"]:
        h = torch.relu(lin(h))
    rgb = model.fc_rgb["0"](h)
    return torch.cat([rgb, alpha], -1).reshape(n, S, 4)


def run_one_iter_of_nerf(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode="train",
                         encode_position_fn=None, encode_direction_fn=None, scene_config=None, randoms=None):
    """Differentiable `run_one_iter_of_nerf` (train_utils.py:185-282 -> predict_and_render_radiance :71-182) for the
    tri-plane model and for the mip/IPE model: same signature and 9-tuple as the reference; gradients reach `planes_` and
    the decoder weights of both models.  Gather / IPE and compositing run on this package's kernels (forward and
    backward), the decoder on torch.
    `randoms` (optional dict: 't_rand' [n,Nc], 'u' [n,Nf], 'noise_c' [n,Nc], 'noise_f' [n,Nc+Nf], unscaled) replaces
    the reference's CPU RNG draws (train_utils.py:108, nerf_helpers.py:683, volume_rendering_utils.py:32)."""
    if not options.nerf.use_viewdirs:
        raise NotImplementedError("nvsr_b200: use_viewdirs=False is not supported")
    if not batch_rays.is_cuda:
        raise _lib.NvsrError("batch_rays must be CUDA tensors: nvsr_b200 has no CPU path")
    # ray batches of `chunksize` like train_utils.py:228-247 (get_minibatches over the rays; with an SR model / mip
    # encoding the reference divides the chunk, :231-234): bounds the live activations of one decoder call
    n = batch_rays.shape[1]
    chunk = int(getattr(getattr(options.nerf, mode), "chunksize", 0) or n)
    if hasattr(model_fine, "SR_model"):
        chunk //= 10
    if getattr(options.nerf, "encode_position_fn", None) == "mip":
        chunk //= 4
    chunk = max(1, chunk)
    if n <= chunk:
        return _run_one_iter(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode, scene_config,
                             randoms, encode_position_fn)
    outs = []
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        rnd = {k: (v[i0:i1] if (torch.is_tensor(v) and v.dim() == 2 and v.shape[0] == n) else v)
               for k, v in (randoms or {}).items()}
        outs.append(_run_one_iter(H, W, focal, model_coarse, model_fine, batch_rays[:, i0:i1], options, scene_id, mode,
                                  scene_config, rnd, encode_position_fn))
    return tuple(None if outs[0][k] is None else torch.cat([o[k] for o in outs], 0) for k in range(9))


def _coarse_context(model_coarse):
    """train_utils.py:88: the whole coarse pass runs inside `model_coarse.optional_no_grad()` — `torch.no_grad` when
    the decoder is not being trained (train_nerf.py:560), a null context otherwise (train_nerf.py:349)."""
    import contextlib
    ctx = getattr(model_coarse, "optional_no_grad", None)
    return ctx() if ctx is not None else contextlib.nullcontext()


def mip_model_forward(model, xyz_feat, dir_feat):
    """FlexibleNeRFModel.forward (models.py:85-108, use_viewdirs, xyz_input_2_dir=False) on torch autograd:
    xyz_feat [rows, dim_xyz] (IPE), dir_feat [rows, dim_dir] -> [rows, 4].  No ReLU after layer1 (models.py:88)."""
    h = model.layer1(xyz_feat)
    for i, lin in enumerate(model.layers_xyz):
        if i % model.skip_connect_every == 0 and i > 0 and i != len(model.layers_xyz):
            h = torch.cat((h, xyz_feat), dim=-1)
        h = torch.relu(lin(h))
    feat = torch.relu(model.fc_feat(h))
    alpha = model.fc_alpha(h)
    h = torch.cat((feat, dir_feat), dim=-1)
    for lin in model.layers_dir:
        h = torch.relu(lin(h))
    return torch.cat((model.fc_rgb(h), alpha), dim=-1)


def _run_one_iter_mip(model_coarse, model_fine, ro, rd, vd, near, far, cfg, scene_id, n_freqs, randoms):
    """mip/IPE branch (train_utils.py:15-64 run_network with cast_rays, mip.py:9-43,154-199): the encodings are data
    (no parameters), so only the decoder (torch autograd) and the compositing (nvsr_composite_bwd, interval edges)
    carry gradients."""
    n, dev = ro.shape[0], ro.device
    Nc, Nf = int(cfg.num_coarse), int(cfg.num_fine)
    radius = ops.mip_radius(scene_id)
    n_dir = (model_coarse.dim_dir - 3) // 6
    if model_coarse.dim_xyz != 6 * n_freqs:
        raise _lib.NvsrError("IPE width does not match the model's dim_xyz")
    from .render import _t_vals
    t = _t_vals(Nc + 1, dev)       # cached per (n, device): an upload from pageable memory synchronises the host
    z = near * (1.0 - t) + far * t if not cfg.lindisp else 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    z = z.expand(n, Nc + 1)
    if cfg.perturb:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper, lower = torch.cat((mids, z[..., -1:]), -1), torch.cat((z[..., :1], mids), -1)
        t_rand = randoms["t_rand"] if "t_rand" in randoms else _rand((n, Nc + 1), dev)
        z = lower + (upper - lower) * t_rand.to(device=dev, dtype=torch.float32)
    z = z.contiguous()
    denc = ops.dir_encoding(vd, n_dir, True)
    std = float(cfg.radiance_field_noise_std)

    def noise_of(name, S):
        """the density noise of one pass, scaled by the std, on the device (volume_rendering_utils.py:30-35)"""
        if std <= 0.0:
            return None
        nse = randoms[name] if name in randoms else _rand((n, S), dev, normal=True)     # the reference draws on the CPU
        return (nse * std).to(device=dev, dtype=torch.float32).contiguous()

    def radiance(model, z_edges):
        S = z_edges.shape[1] - 1
        with torch.no_grad():
            enc = ops.ipe(z_edges, ro, rd, radius, n_freqs)
            dirs = denc[:, None, :].expand(n, S, denc.shape[-1]).reshape(n * S, -1)
        return mip_model_forward(model, enc, dirs).reshape(n, S, 4)

    with _coarse_context(model_coarse):
        rf = radiance(model_coarse, z)
        rgb_c, disp_c, acc_c, weights, _ = _render(rf, z, rd, 0.0, cfg.white_background, None, mip=True, nz=noise_of("noise_c", Nc))
    rgb_f = disp_f = acc_f = None
    if Nf > 0:
        with torch.no_grad():
            mid = 0.5 * (z[..., 1:] + z[..., :-1])
            mid = 0.5 * (mid[..., 1:] + mid[..., :-1])
            u = randoms.get("u")
            if u is None and cfg.perturb != 0.0:
                u = _rand([n, Nf + 1], dev)
            z_samples = ops.sample_pdf(mid, weights[..., 1:-1], Nf + 1, det=(cfg.perturb == 0.0), u=u)
            z_f = ops.sort_cat(z, z_samples)
            if "z_fine" in randoms:     # test hook: teacher-forced merged depths
                z_f = randoms["z_fine"].to(device=dev, dtype=torch.float32).contiguous()
        rf_f = radiance(model_fine, z_f)
        rgb_f, disp_f, acc_f, _, _ = _render(rf_f, z_f, rd, 0.0, cfg.white_background, None, mip=True,
                                             nz=noise_of("noise_f", z_f.shape[1] - 1))
    return rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, None, None, None


def _run_one_iter(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode, scene_config, randoms,
                  encode_position_fn=None):
    cfg = getattr(options.nerf, mode)
    randoms = randoms or {}
    ro_in, rd_in = batch_rays[0], batch_rays[1]
    dev = ro_in.device
    use_ndc = scene_config.no_ndc is False
    ro, rd, vd = ops.prepare_rays(ro_in, rd_in, use_ndc, H, W, focal if use_ndc else 1.0, 1.0)
    n = ro.shape[0]
    near, far = float(scene_config.near), float(scene_config.far)
    Nc, Nf = int(cfg.num_coarse), int(cfg.num_fine)
    if getattr(options.nerf, "encode_position_fn", None) == "mip":
        n_freqs = getattr(encode_position_fn, "max_freq", None)
        if n_freqs is None or hasattr(model_coarse, "planes_"):
            raise NotImplementedError("nvsr_b200: the mip path needs an IntegratedPositionalEncoding and a FlexibleNeRFModel")
        return _run_one_iter_mip(model_coarse, model_fine, ro, rd, vd, near, far, cfg, scene_id, n_freqs, randoms)

    def draw(name, shape):
        t = randoms[name] if name in randoms else _rand(shape, dev)     # the reference draws on the CPU
        return t.to(device=dev, dtype=torch.float32)

    def coarse_depths():
        # stratified depths (train_utils.py:95-109); data, not parameters: no gradient
        from .render import _t_vals
        t_vals = _t_vals(Nc, dev)  # cached per (n, device): an upload from pageable memory synchronises the host
        if not cfg.lindisp:
            z = near * (1.0 - t_vals) + far * t_vals
        else:
            z = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
        z = z.expand(n, Nc)
        if cfg.perturb:
            mids = 0.5 * (z[..., 1:] + z[..., :-1])
            upper, lower = torch.cat((mids, z[..., -1:]), -1), torch.cat((z[..., :1], mids), -1)
            z = lower + (upper - lower) * draw("t_rand", (n, Nc))
        z = z.contiguous()
        return z

    std = float(cfg.radiance_field_noise_std)

    def noise_of(name, S):
        """the density noise of one pass, scaled by the std, on the device (volume_rendering_utils.py:30-35)"""
        if std <= 0.0:
            return None
        nse = randoms[name] if name in randoms else _rand((n, S), dev, normal=True)     # the reference draws on the CPU
        return (nse * std).to(device=dev, dtype=torch.float32).contiguous()

    z_f = None
    if _state["fast_frozen_coarse"] and getattr(model_coarse, "optional_no_grad", None) is torch.no_grad:
        # decoder frozen (train_nerf.py:560): the coarse pass carries no gradient at all, so it can run on the forward
        # kernels (16-bit gather + tcgen05 decoder + compositing fused with sample_pdf and the sort-merge)
        from . import render
        with torch.no_grad():
            pc = render._planes_pass(model_coarse, scene_id, render._state["precision"])
            rnd = {k: v.to(dev) for k, v in randoms.items() if k in ("t_rand", "u", "noise_c") and torch.is_tensor(v)}
            co, _ = render._render_planes_chunk(pc, None, ro, rd, vd, near, far, cfg, rnd, None, coarse_only=True)
        rgb_c, disp_c, acc_c = co["rgb"], co["disp"], co["acc"]
        z_f = co.get("z_merged")
    else:
        z = coarse_depths()
        with _coarse_context(model_coarse):
            nz = noise_of("noise_c", Nc)
            rf = planes_model_forward(model_coarse, scene_id, ro, rd, z, vd, sigma_noise=nz)
            rgb_c, disp_c, acc_c, weights, _ = _render(rf, z, rd, 0.0, cfg.white_background, None, nz=nz)
    rgb_f = disp_f = acc_f = None
    if Nf > 0 and z_f is not None:
        nz = noise_of("noise_f", Nc + Nf)
        rf_f = planes_model_forward(model_fine, scene_id, ro, rd, z_f, vd, sigma_noise=nz)
        rgb_f, disp_f, acc_f, _, _ = _render(rf_f, z_f, rd, 0.0, cfg.white_background, None, nz=nz)
    elif Nf > 0:
        with torch.no_grad():    # z_samples.detach() (train_utils.py:153)
            mid = 0.5 * (z[..., 1:] + z[..., :-1])
            u = randoms.get("u")
            if u is None and cfg.perturb != 0.0:
                u = _rand([n, Nf], dev)
            z_samples = ops.sample_pdf(mid, weights[..., 1:-1], Nf, det=(cfg.perturb == 0.0), u=u)
            z_f = ops.sort_cat(z, z_samples)
            if "z_fine" in randoms:     # test hook: teacher-forced merged depths
                z_f = randoms["z_fine"].to(device=dev, dtype=torch.float32).contiguous()
            if isinstance(randoms.get("trace"), dict):
                randoms["trace"]["z_fine"] = z_f
        nz = noise_of("noise_f", Nc + Nf)
        rf_f = planes_model_forward(model_fine, scene_id, ro, rd, z_f, vd, sigma_noise=nz)
        rgb_f, disp_f, acc_f, _, _ = _render(rf_f, z_f, rd, 0.0, cfg.white_background, None, nz=nz)
    return rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, None, None, None
