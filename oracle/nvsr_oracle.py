"""CPU ORACLE for the ray-rendering hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module; the product path (neural-volume-super-resolution_b200/) never does and has no CPU path.

What it is: a restatement, in plain torch-on-CPU fp32 ops, of the reference's algorithm for the path
SURVEY.md §8(a) lists.  The reference is pure Python over ATen ops, so the faithful restatement is
the same ATen op sequence (grid_sample, linear, cumsum, cumprod, searchsorted, sort ...) — on one
machine it is bit-identical to the reference, which is how it was pinned:

  PINNING: the reference has NO tests / golden vectors for this path ("parity
  unpinned" by the reference's own tests).  This oracle is instead pinned against OUTPUTS OF THE
  REFERENCE ITSELF, run in the build container by tests/golden/make_golden.py (committed) through
  train_utils.run_one_iter_of_nerf / eval_nerf and the stage functions; the resulting vectors are
  committed under tests/golden/*.npz and tests/test_oracle_golden.py checks this file against them.

Each function cites the reference file:line it follows (paths relative to the reference checkout).
Models are duck-typed: anything exposing the attributes of the reference's TwoDimPlanesModel /
FlexibleNeRFModel works (the real classes in the build container, the stand-ins of
neural-volume-super-resolution_b200/scene.py on the GPU box).
"""
import math
import re

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# a1  nerf_helpers.py:507-549 (get_ray_bundle), :396-406 (meshgrid_xy), :432-437 (get_focal)
def _focal(f, dim):
    if isinstance(f, (list, tuple)):
        return f[1] if dim == "H" else f[0]
    return f


def get_ray_bundle(height, width, focal_length, tform_cam2world, padding_size=0, downsampling_offset=0):
    xs = (torch.arange(width + 2 * padding_size) + downsampling_offset).to(tform_cam2world)
    ys = (torch.arange(height + 2 * padding_size) + downsampling_offset).to(tform_cam2world)
    gi, gj = torch.meshgrid(xs, ys, indexing="ij")
    ii, jj = gi.transpose(-1, -2), gj.transpose(-1, -2)
    if padding_size > 0:
        ii = ii - padding_size
        jj = jj - padding_size
    cam_dirs = torch.stack(
        [(ii - width * 0.5) / _focal(focal_length, "H"), -(jj - height * 0.5) / _focal(focal_length, "W"),
         -torch.ones_like(ii)], dim=-1)
    rd = torch.sum(cam_dirs[..., None, :] * tform_cam2world[:3, :3], dim=-1)
    ro = tform_cam2world[:3, -1].expand(rd.shape)
    return ro, rd


# a3  nerf_helpers.py:578-605
def ndc_rays(H, W, focal, near, rays_o, rays_d):
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
    oy = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
    oz = 1.0 + 2.0 * near / rays_o[..., 2]
    dx = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    dy = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    dz = -2.0 * near / rays_o[..., 2]
    return torch.stack([ox, oy, oz], -1), torch.stack([dx, dy, dz], -1)


# nerf_helpers.py:492-496
def cart2az_el(dirs):
    el = torch.atan2(dirs[..., 2], torch.sqrt(torch.sum(dirs[..., :2] ** 2, -1)))
    az = torch.atan2(dirs[..., 1], dirs[..., 0])
    return torch.stack([az, el], -1)


# nerf_helpers.py:552-575
def positional_encoding(tensor, num_encoding_functions=6, include_input=True):
    parts = [tensor] if include_input else []
    for i in range(num_encoding_functions):
        for fn in (torch.sin, torch.cos):
            parts.append(fn(2.0 ** i * tensor))
    return torch.cat(parts, dim=-1)


# --------------------------------------------------------------------------------------------------
# a5/a6  models.py:261-421 — TwoDimPlanesModel.forward
def plane_name(scene_id, d):  # models.py:110-113
    return "_D%d" % d if scene_id is None else "sc%s_D%d" % (scene_id, d)


# f2  models.py:773-822 (EDSR with padding=0) and :884-926 (PlanesSR.forward, full plane, eval mode)
def edsr_forward(net, x):
    out = F.conv2d(x, net.conv_input.weight)
    for blk in net.residual:
        m = 2 * (blk.conv1.kernel_size[0] // 2)               # _Residual_Block.margins (models.py:776)
        ident = out[..., m:-m, m:-m] if m else out
        y = F.conv2d(F.relu(F.conv2d(out, blk.conv1.weight)), blk.conv2.weight)
        y = y * 0.1                                           # output *= 0.1
        out = torch.add(y, ident)
    out = F.conv2d(out, net.conv_mid.weight)
    for m in net.upscale:
        out = F.pixel_shuffle(out, 2) if isinstance(m, torch.nn.PixelShuffle) else F.conv2d(out, m.weight)
    return F.conv2d(out, net.conv_output.weight)


def planes_sr_forward(sr, name):
    """PlanesSR.forward(plane_name) (models.py:884-926) for a full plane in eval mode, on the device of the LR plane."""
    lr = sr.LR_planes[name].detach()
    x = 1 * lr
    if hasattr(sr, "planes_mean_NON_LEARNED"):
        x = (x - sr.planes_mean_NON_LEARNED.to(x)) / sr.planes_std_NON_LEARNED.to(x)
    pad = int(sr.inner_model.required_padding)
    x = F.pad(x, pad=(pad, pad, pad, pad), mode="replicate")
    crop = int(sr.HR_overpadding)
    diff = edsr_forward(sr.inner_model, x)
    if crop > 0:
        diff = diff[..., crop:-crop, crop:-crop]
    resid = F.interpolate(lr, scale_factor=sr.scale_factor, mode=sr.plane_interp, align_corners=sr.align_corners)
    return torch.add(diff, resid)


def _plane_tensor(model, d, super_resolve):
    """models.py:270-284 (`planes`) without the hard-coded .cuda()."""
    name = plane_name(model.cur_id, d)
    coupler = model.scene_coupler
    saved = coupler.scene_with_saved_plane(name, plane_not_scene=True)
    if super_resolve:
        sr = model.SR_model
        if hasattr(sr, "inner_model") and hasattr(sr.inner_model, "conv_input"):   # an EDSR-based PlanesSR: restated above
            if saved not in sr.LR_planes:
                sr.set_LR_plane(model.raw_plane(saved, False, detach=True), id=saved, save_interpolated=False)
            cache = sr.__dict__.setdefault("_oracle_sr_cache", {})
            lrp = sr.LR_planes[saved]
            key = (saved, lrp.data_ptr(), lrp._version, str(lrp.device))
            if key not in cache:
                with torch.no_grad():
                    cache[key] = planes_sr_forward(sr, saved)
            return cache[key]
        return model.SR_model(saved)
    return model.raw_plane(saved, coupler.should_downsample(name), detach=False)


def _should_sr(model, d):
    """models.py:296-300"""
    name = plane_name(model.cur_id, d)
    sr = hasattr(model, "SR_model") and (not hasattr(model, "scene_coupler") or
                                         model.scene_coupler.should_SR(name, plane_not_scene=True))
    return bool(sr and not model.skip_SR_)


def planes_gather(model, x6):
    """Gather half of forward (models.py:381-391): returns (list of 3 [n,C] projections, [n,C] view proj)."""
    c = torch.cat([x6[..., :3], cart2az_el(x6[..., 3:])], -1)
    box = model.box_coords[model.cur_id + ""]
    cn = 2 * (c - box[:1].type(c.type())) / (box[1:] - box[:1]).type(c.type()) - 1  # models.py:264-265
    rots = model.coord_projector.rot_mats_NON_LEARNED
    pos = []
    for d in range(model.num_density_planes):
        grid = torch.matmul(cn[..., :3], rots[d][:, 1:].type(cn.type())).reshape([1, cn.shape[0], 1, 2])
        plane = _plane_tensor(model, d, _should_sr(model, d))
        smp = F.grid_sample(input=plane, grid=grid, mode=model.plane_interp, align_corners=model.align_corners,
                            padding_mode="border")
        pos.append(smp.squeeze(0).squeeze(-1).permute(1, 0))
    vgrid = cn[..., 3:].reshape([1, cn.shape[0], 1, 2])
    vplane = _plane_tensor(model, model.num_density_planes, False)
    view = F.grid_sample(input=vplane, grid=vgrid, mode=model.plane_interp, align_corners=model.align_corners,
                         padding_mode="border").squeeze(0).squeeze(-1).permute(1, 0)
    return pos, view


def _skip(model, layer_num):  # models.py:203-207
    s = model.skip_connect_every
    return False if s is None else (layer_num % s == 0 and layer_num > 0)


def planes_decode(model, pos, view):
    """Decoder half (models.py:393-421), eval mode => ensemble member '0'."""
    assert model.proj_combination in ("avg", "sum") and model.viewdir_proj_combination == "concat_pos"
    assert model.rgb_dec_input == "projections" and model.use_viewdirs
    # combine_pos_planes (models.py:355-361)
    mean = torch.stack(pos, 0).mean(0) if model.proj_combination == "avg" else torch.stack(pos, 0).sum(0)
    h = 1 * mean
    for i, lin in enumerate(model.density_dec["0"]):
        if _skip(model, i - 1):
            h = torch.cat((h, mean), dim=-1)
        h = F.relu(lin(h))
    alpha = model.fc_alpha["0"](h)
    x_rgb = torch.cat(list(pos) + [view], 1)
    h = x_rgb
    for i, lin in enumerate(model.rgb_dec["0"]):
        if _skip(model, i - 1):
            h = torch.cat((h, x_rgb), dim=-1)
        h = F.relu(lin(h))
    rgb = model.fc_rgb["0"](h)
    return torch.cat((rgb, alpha), dim=-1)


def planes_model_forward(model, x6):
    pos, view = planes_gather(model, x6)
    return planes_decode(model, pos, view)


# a6'  models.py:85-108 — FlexibleNeRFModel.forward (use_viewdirs=True, xyz_input_2_dir=False)
def flexible_model_forward(model, x):
    xyz, view = x[..., : model.dim_xyz], x[..., model.dim_xyz:]
    h = model.layer1(xyz)  # NOTE: no ReLU after layer1 (models.py:88)
    for i in range(len(model.layers_xyz)):
        if i % model.skip_connect_every == 0 and i > 0 and i != len(model.layers_xyz):
            h = torch.cat((h, xyz), dim=-1)
        h = F.relu(model.layers_xyz[i](h))
    feat = F.relu(model.fc_feat(h))
    alpha = model.fc_alpha(h)
    h = torch.cat((feat, view), dim=-1)
    for lin in model.layers_dir:
        h = F.relu(lin(h))
    return torch.cat((model.fc_rgb(h), alpha), dim=-1)


def is_planes_model(model):
    return hasattr(model, "planes_") or hasattr(model, "coord_projector")


def model_forward(model, x):
    return planes_model_forward(model, x) if is_planes_model(model) else flexible_model_forward(model, x)


# --------------------------------------------------------------------------------------------------
# a7  volume_rendering_utils.py:6-51 + cumprod_exclusive nerf_helpers.py:409-430
def cumprod_exclusive(t):
    cp = torch.roll(torch.cumprod(t, -1), 1, -1)
    cp[..., 0] = 1.0
    return cp


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, mip_nerf=False, noise=None):
    big = torch.tensor([1e10], dtype=ray_directions.dtype, device=ray_directions.device)
    dists = depth_values[..., 1:] - depth_values[..., :-1]
    if not mip_nerf:
        dists = torch.cat((dists, big.expand(depth_values[..., :1].shape)), dim=-1)
    dists = dists * ray_directions[..., None, :].norm(p=2, dim=-1)
    rgb = torch.sigmoid(radiance_field[..., :3])
    nz = 0.0
    if radiance_field_noise_std > 0.0:
        if noise is None:
            noise = torch.randn(radiance_field[..., 3].shape)
        nz = (noise * radiance_field_noise_std).to(radiance_field)
    sigma = F.relu(radiance_field[..., 3] + nz)
    alpha = 1.0 - torch.exp(-sigma * dists)
    weights = alpha * cumprod_exclusive(1.0 - alpha + 1e-10)
    rgb_map = (weights[..., None] * rgb).sum(dim=-2)
    if mip_nerf:
        depth_values = 0.5 * (depth_values[:, :-1] + depth_values[:, 1:])
    depth_map = (weights * depth_values).sum(dim=-1)
    acc_map = weights.sum(dim=-1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_background:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


# a8  nerf_helpers.py:668-702 (sample_pdf_2, bound as sample_pdf at train_utils.py:4)
def sample_pdf(bins, weights, num_samples, det=False, u=None, return_all=False):
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=num_samples)
            u = u.expand(list(cdf.shape[:-1]) + [num_samples]).to(weights)
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [num_samples]).to(weights)
    else:
        u = u.expand(list(cdf.shape[:-1]) + [num_samples]).to(weights)
    u = u.contiguous()
    cdf = cdf.contiguous()
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min(cdf.shape[-1] - 1 * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], dim=-1)
    shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    if return_all:
        return samples, inds, cdf
    return samples


def searchsorted_lerp(cdf, bins, u):
    """The search/gather/lerp tail of sample_pdf on a GIVEN cdf (stage test of SURVEY §7)."""
    u = u.expand(list(cdf.shape[:-1]) + [u.shape[-1]]).contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, side="right")
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bb + (u - cb) / denom * (ba - bb), inds


# --------------------------------------------------------------------------------------------------
# a9  mip.py:9-43 and :154-199
def cast_rays(t_vals, origins, directions, radii):
    t0, t1 = t_vals[..., :-1], t_vals[..., 1:]
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    t_mean = mu + (2 * mu * hw ** 2) / (3 * mu ** 2 + hw ** 2)
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / (3 * mu ** 2 + hw ** 2) ** 2)
    r_var = radii ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / (3 * mu ** 2 + hw ** 2))
    d = directions
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.maximum(torch.tensor(1e-10).type(d.type()), torch.sum(d ** 2, axis=-1, keepdims=True))
    d_sq = d ** 2
    null_diag = 1 - d_sq / d_mag_sq
    cov = t_var[..., None] * d_sq[..., None, :] + r_var[..., None] * null_diag[..., None, :]
    return mean + origins[..., None, :], cov


def integrated_pos_enc(means, covs, multires):
    """IntegratedPositionalEncoding(3, multires).forward((means, covs)) — mip.py:170-199."""
    scales = torch.tensor([2 ** i for i in range(0, multires - 1)], device=means.device)
    shape = list(means.shape[:-1]) + [-1]
    y = torch.reshape(means[..., None, :] * scales[:, None], shape)
    y_var = torch.reshape(covs[..., None, :] * scales[:, None] ** 2, shape)
    xx = torch.cat([y, y + 0.5 * np.pi], dim=-1)
    vv = torch.cat([y_var] * 2, dim=-1)
    return torch.exp(-0.5 * vv) * torch.sin(xx)


def mip_radius(scene_id):  # train_utils.py:21-23
    dx = int(re.search(r"(?<=_DS)(\d)+(?=$)", scene_id).group(0)) * 0.00135
    return dx * 2 / np.sqrt(12.0)


# --------------------------------------------------------------------------------------------------
# a4/a2  train_utils.py:15-64 (run_network), :71-182 (predict_and_render_radiance), :185-282, :285-331
def _chunks(t, n):
    return [t] if n is None else [t[i:i + n] for i in range(0, t.shape[0], n)]


def run_network(model, pts, ray_batch, chunksize, embed_fn, embeddirs_fn, scene_id, mip_nerf=False, z_vals=None):
    shape = list(pts.shape)
    if mip_nerf:
        ro, rd, _, _, _ = torch.split(ray_batch, [3, 3, 1, 1, 3], dim=-1)
        means, covs = cast_rays(z_vals, ro, rd, mip_radius(scene_id))
        flat = embed_fn((means, covs))
        shape[1] = flat.shape[1]
        emb = flat.reshape((-1, flat.shape[-1]))
    else:
        emb = embed_fn(pts.reshape((-1, shape[-1])))
    if embeddirs_fn is not None:
        dirs = ray_batch[..., None, -3:].expand(shape)
        emb = torch.cat((emb, embeddirs_fn(dirs.reshape((-1, dirs.shape[-1])))), dim=-1)
    out = torch.cat([model_forward(model, b) for b in _chunks(emb, chunksize)], dim=0)
    return out.reshape(shape[:-1] + [out.shape[-1]])


def _identity(x):
    return x


def predict_and_render_radiance(ray_batch, model_coarse, model_fine, options, scene_id, mode="train",
                                encode_position_fn=None, encode_direction_fn=None, randoms=None, trace=None):
    """`randoms` (dict with optional 't_rand' [n,Nc], 'u' [n,Nf], 'noise_c', 'noise_f') replaces the
    reference's CPU RNG draws (train_utils.py:108, nerf_helpers.py:683, volume_rendering_utils.py:32)
    so that CPU oracle and GPU path see identical draws.  `trace` (dict) receives stage tensors."""
    mip = getattr(options.nerf, "encode_position_fn", None) == "mip"
    cfg = getattr(options.nerf, mode)
    embed = encode_position_fn or _identity
    embeddirs = encode_direction_fn or _identity
    randoms = randoms or {}
    trace = trace if trace is not None else {}
    ro, rd = ray_batch[..., :3], ray_batch[..., 3:6]
    bounds = ray_batch[..., 6:8].reshape(list(ray_batch.shape)[:-1] + [1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0.0, 1.0, cfg.num_coarse + mip).to(ro)
    if not cfg.lindisp:
        z_vals = near * (1.0 - t_vals) + far * t_vals
    else:
        z_vals = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
    z_vals = z_vals.expand(list(ray_batch.shape)[:-1] + [cfg.num_coarse + mip])
    if cfg.perturb:
        mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat((mids, z_vals[..., -1:]), dim=-1)
        lower = torch.cat((z_vals[..., :1], mids), dim=-1)
        t_rand = randoms["t_rand"] if "t_rand" in randoms else torch.rand(z_vals.shape)
        z_vals = lower + (upper - lower) * t_rand.to(ro)
    pts = ro[..., None, :] + rd[..., None, :] * z_vals[..., :, None]
    rf = run_network(model_coarse, pts, ray_batch, cfg.chunksize, embed, embeddirs, scene_id, mip, z_vals)
    rgb_c, disp_c, acc_c, weights, depth_c = volume_render_radiance_field(
        rf, z_vals, rd, cfg.radiance_field_noise_std, cfg.white_background, mip, noise=randoms.get("noise_c"))
    trace.update(z_coarse=z_vals, raw_coarse=rf, weights_coarse=weights, depth_coarse=depth_c)
    rgb_f = disp_f = acc_f = None
    if cfg.num_fine > 0:
        mid = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        if mip:
            mid = 0.5 * (mid[..., 1:] + mid[..., :-1])
        z_samples, inds, cdf = sample_pdf(mid, weights[..., 1:-1], cfg.num_fine + mip, det=(cfg.perturb == 0.0),
                                          u=randoms.get("u"), return_all=True)
        z_samples = z_samples.detach()
        z_vals, _ = torch.sort(torch.cat((z_vals, z_samples), dim=-1), dim=-1)
        if "z_fine" in randoms:   # test hook (not in the reference): teacher-forced fine depths, e.g. the CUDA path's
            z_vals = randoms["z_fine"].to(ro)
        pts = ro[..., None, :] + rd[..., None, :] * z_vals[..., :, None]
        rf = run_network(model_fine, pts, ray_batch, cfg.chunksize, embed, embeddirs, scene_id, mip, z_vals)
        rgb_f, disp_f, acc_f, w_f, depth_f = volume_render_radiance_field(
            rf, z_vals, rd, cfg.radiance_field_noise_std, cfg.white_background, mip, noise=randoms.get("noise_f"))
        trace.update(inds=inds, cdf=cdf, z_samples=z_samples, z_fine=z_vals, raw_fine=rf, depth_fine=depth_f)
    return rgb_c, disp_c, acc_c, rgb_f, disp_f, acc_f, None, None, None


def run_one_iter_of_nerf(H, W, focal, model_coarse, model_fine, batch_rays, options, scene_id, mode="train",
                         encode_position_fn=None, encode_direction_fn=None, scene_config=None, randoms=None,
                         trace=None):
    if is_planes_model(model_coarse):
        model_coarse.set_cur_scene_id(scene_id)
        model_fine.set_cur_scene_id(scene_id)
    ro_in, rd_in = batch_rays[0], batch_rays[1]
    viewdirs = None
    if options.nerf.use_viewdirs:
        viewdirs = rd_in / rd_in.norm(p=2, dim=-1).unsqueeze(-1)
        viewdirs = viewdirs.reshape((-1, 3))
    if scene_config.no_ndc is False:
        ro, rd = ndc_rays(H, W, focal, 1.0, ro_in, rd_in)
        ro, rd = ro.reshape((-1, 3)), rd.reshape((-1, 3))
    else:
        ro, rd = ro_in.reshape((-1, 3)), rd_in.reshape((-1, 3))
    near = scene_config.near * torch.ones_like(rd[..., :1])
    far = scene_config.far * torch.ones_like(rd[..., :1])
    rays = torch.cat((ro, rd, near, far), dim=-1)
    if options.nerf.use_viewdirs:
        rays = torch.cat((rays, viewdirs), dim=-1)
    # ray-batch chunking (train_utils.py:228-235) does not change results (nothing couples rays);
    # the oracle keeps it only to bound memory, and slices `randoms` consistently.
    chunk = getattr(options.nerf, mode).chunksize
    outs = []
    traces = []
    for i in range(0, rays.shape[0], chunk):
        sub = None
        if randoms:
            sub = {k: v[i:i + chunk] if (torch.is_tensor(v) and v.dim() == 2 and v.shape[0] == rays.shape[0]) else v
                   for k, v in randoms.items()}
        tr = {} if trace is not None else None
        outs.append(predict_and_render_radiance(rays[i:i + chunk], model_coarse, model_fine, options, scene_id,
                                                mode=mode, encode_position_fn=encode_position_fn,
                                                encode_direction_fn=encode_direction_fn, randoms=sub, trace=tr))
        if tr is not None:
            traces.append(tr)
    if trace is not None and traces:
        for k in traces[0]:
            trace[k] = torch.cat([t[k] for t in traces], 0)
    cat = lambda j: None if outs[0][j] is None else torch.cat([o[j] for o in outs], dim=0)
    return tuple(cat(j) for j in range(9))


def eval_nerf(height, width, focal_length, model_coarse, model_fine, ray_origins, ray_directions, options, scene_id,
              mode="validation", encode_position_fn=None, encode_direction_fn=None, scene_config=None):
    batch = torch.cat((ray_origins.reshape((1, -1, 3)), ray_directions.reshape((1, -1, 3))), dim=0)
    rgb_c, _, _, rgb_f, _, _, _, _, _ = run_one_iter_of_nerf(
        height, width, focal_length, model_coarse, model_fine, batch, options, scene_id, mode="validation",
        encode_position_fn=encode_position_fn, encode_direction_fn=encode_direction_fn, scene_config=scene_config)
    rgb_c = rgb_c.reshape([height, width, -1])
    if rgb_fine_present := (rgb_f is not None):
        rgb_f = rgb_f.reshape([height, width, -1])
    return rgb_c, None, None, rgb_f, None, None, None, None, None


# load_blender.py:15-39 (pose_spherical): camera on a sphere looking at the origin
def pose_spherical(theta, phi, radius):
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]],
                        dtype=np.float32)

    def rot_theta(th):
        return np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]],
                        dtype=np.float32)

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180.0 * np.pi) @ c2w
    c2w = rot_theta(theta / 180.0 * np.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    return c2w
