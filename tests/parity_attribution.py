"""Explained-outlier parity gate (TEST INFRASTRUCTURE; imports the oracle).

A max-norm comparison of rendered maps against the reference is ill-posed at exactly two places of the REFERENCE's own
algorithm, whatever the precision of the implementation under test:
  (ii) the last interval of a ray is 1e10 long (volume_rendering_utils.py:20-27), so alpha_last is a STEP function of
       sigma_last (+ noise) at 0: a sample whose reference sigma lies within the implementation's rounding error of 0
       may land on the other side, and the ray's maps then move by T_last * (1, rgb_last, z_last);
  (i)  searchsorted indices flip where u lies within ~2 ulp of a cdf edge (the reference's own CPU and
       CUDA builds differ the same way), and a resampled depth inside a near-empty bin is amplified by 1/denom
       (nerf_helpers.py:696-699) — the fine pass then runs at different depths.
The gate below therefore checks a CHAIN of links, each with a max-norm bound over EVERY ray of the case, and attributes
every ray that exceeds a bound to (i) or (ii) constructively — anything else fails:

  L1 coarse pass      raw (sigma, colour logits) of every sample vs the oracle: max-norm bound per mode.
                      Maps (rgb, acc, depth, disp + its NaN pattern) vs the oracle: bound B on every ray whose last-sample
                      sign agrees; a ray whose last-sample sign differs (necessarily |sigma_ref| <= the mode's sigma
                      bound) must match, within B, the ORACLE's maps recomputed with only that one sigma replaced by the
                      implementation's value ("hybrid"): the whole excess is the step.
  L2 resampling       given the implementation's OWN coarse weights, the oracle's sample_pdf must give the same bin indices
                      except flips within 2 ulp of a cdf edge, samples within the formula's conditioning bound, and the merged
                      depths must be exactly sort(cat(z, z_samples)).
  L3 fine pass        teacher-forced: the oracle's fine network evaluated at the implementation's merged depths; same
                      raw / map / hybrid checks as L1.
  L4 free-running     maps vs the oracle's own free-running fine maps.  fp32: a ray above B must have merged depths that
                      differ from the oracle's (then L2 has explained them) or a last-sample step.  16-bit modes: the
                      coarse weights legitimately differ by the mode's rounding, which moves every resampled depth
                      continuously — rays without a flip or step are bounded by B_free (stated per mode) except for at
                      most 0.5 % of the rays (near-empty bins amplify a depth by 1/denom), each of which must be
                      reproduced by the oracle's own fine network at the implementation's depths to within L3's bound.
"""
import torch

import helpers as H
from oracle import nvsr_oracle as O

# mode -> bounds.  sigma / logit: max |raw - oracle raw| over every sample.  B: max abs map error (rgb, acc; depth is
# compared relative to the far bound) of rays without a last-sample step.  B_free: the same for the free-running fine maps.
BOUNDS = {
    "fp32": dict(sigma=1e-3, logit=1e-4, B=1e-3, B_free=1e-3),
    "fp16": dict(sigma=0.15, logit=2e-3, B=1e-2, B_free=5e-2),
    # fp16 everywhere except the density chain (split operands, three tcgen05 passes, fp32 features): the north-star
    # 1e-3 map bound itself, unscaled; the colour chain is plain fp16 (logits ~3e-5 off)
    "fp16-split": dict(sigma=2e-3, logit=2e-3, B=1e-3, B_free=5e-3),
    "bf16": dict(sigma=1.0, logit=1.5e-2, B=6e-2, B_free=1.5e-1),
}


def _prep_rays(c):
    """the (ro, rd) the compositing sees: NDC-mapped when the scene asks for it (train_utils.py:215-218)"""
    ro, rd = c["batch"][0].cpu(), c["batch"][1].cpu()
    if c["scfg"].no_ndc is False:
        ro, rd = O.ndc_rays(c["H"], c["W"], c["focal"], 1.0, ro, rd)
    return ro, rd


def _oracle(c, randoms=None, trace=None):
    import copy
    mc, mf = c["mc"], c["mf"]
    if next(mc.parameters()).device.type != "cpu":
        mc, mf = copy.deepcopy(mc).cpu(), copy.deepcopy(mf).cpu()
        for m in (mc, mf):
            if hasattr(m, "box_coords"):
                m.box_coords = {k: v.cpu() for k, v in m.box_coords.items()}
            sr = getattr(m, "SR_model", None)          # stored / cached SR planes are plain tensors, not parameters
            if sr is not None and hasattr(sr, "SR_planes"):
                sr.SR_planes = {k: v.cpu() for k, v in sr.SR_planes.items()}
                if hasattr(sr, "LR_planes"):
                    sr.LR_planes = {k: v.cpu() for k, v in sr.LR_planes.items()}
    rnd = {k: v.cpu() for k, v in (randoms or {}).items()}
    enc, encd = c.get("enc"), c.get("encd")
    if c.get("kind") == "mip":
        enc = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
        encd = lambda x: O.positional_encoding(x, 4, True)
    with torch.no_grad():
        return O.run_one_iter_of_nerf(c["H"], c["W"], c["focal"], mc, mf, c["batch"].cpu(), c["opt"], c["sid"], "validation",
                                      encode_position_fn=enc, encode_direction_fn=encd, scene_config=c["scfg"],
                                      randoms=rnd, trace=trace)


def _maps_of(raw, z, rd, cfg, mip, noise):
    std = float(cfg.radiance_field_noise_std)
    rgb, disp, acc, w, depth = O.volume_render_radiance_field(raw, z, rd, std, bool(cfg.white_background), mip_nerf=mip,
                                                              noise=noise)
    return dict(rgb=rgb, disp=disp, acc=acc, depth=depth, weights=w)


def _map_err(a, b, far):
    """per-ray max abs error over rgb / acc / depth (depth relative to the far bound), NaNs of disp compared apart"""
    e = (a["rgb"] - b["rgb"]).abs().max(-1)[0]
    e = torch.maximum(e, (a["acc"] - b["acc"]).abs())
    e = torch.maximum(e, (a["depth"] - b["depth"]).abs() / far)
    return e


def _disp_ok(a, b, tol, sign_flip, far):
    """disp = 1/max(1e-10, depth/acc) is a ratio of two maps: NaN exactly where acc == 0, and its relative error is the
    sum of those of acc and depth.  Finite values are compared with that conditioning (acc within tol and depth within
    tol*far give a relative error of tol/acc + tol*far/depth); the NaN pattern must be identical except on rays where
    some sample's sigma changed sign (necessarily |sigma_ref| <= the mode's sigma bound): a ray the reference sees as
    empty may then carry a weight of the order of the map bound."""
    na, nb = torch.isnan(a["disp"]), torch.isnan(b["disp"])
    nan_ok = (na == nb) | (sign_flip & (a["acc"] <= tol) & (b["acc"] <= tol))
    fin = ~na & ~nb
    rel = (a["disp"] - b["disp"]).abs() / (1 + b["disp"].abs())
    tiny = tol * 1e-3
    cond = tol / torch.clamp(torch.minimum(a["acc"], b["acc"]), min=tiny) + \
        tol * far / torch.clamp(torch.minimum(a["depth"].abs(), b["depth"].abs()), min=tiny)
    return nan_ok & (~fin | (rel <= torch.clamp(cond, min=tol)))


EXACT_MODES = ("fp32", "fp16-split")   # modes held to the 1e-3 contract as stated (no interval scaling of B)
REF_INTERVAL = 0.07   # (far - near) / 63 * |rd| of the reference's 64-sample coarse pass on a Blender-shaped scene


def _interval_scale(z, rd, mip):
    """alpha = 1 - exp(-relu(sigma) * dist) is Lipschitz in sigma with constant dist, so a map bound stated for the
    reference's sampling density scales with the longest (non-terminal) interval of the case: the small goldens use 16
    coarse samples (0.27-long intervals, ~1 with lindisp)."""
    d = (z[:, 1:] - z[:, :-1]) * rd.norm(dim=-1, keepdim=True)
    return max(1.0, float(d.max()) / REF_INTERVAL)


def _pass_link(tag, prec, raw_g, raw_o, maps_g, z, rd, cfg, mip, noise, far, report):
    """L1 / L3: raw bounds on every sample, map bound on every ray, last-sample steps attributed by the hybrid."""
    bd = BOUNDS[prec]
    B = bd["B"] * (_interval_scale(z, rd, mip) if prec not in EXACT_MODES else 1.0)   # the north-star bound itself
    report[f"{tag}_map_bound"] = B
    std = float(cfg.radiance_field_noise_std)
    nz = None if (noise is None or std <= 0) else noise * std
    d_sig = (raw_g[..., 3] - raw_o[..., 3]).abs()
    d_log = (raw_g[..., :3] - raw_o[..., :3]).abs()
    report[f"{tag}_sigma_maxdiff"] = float(d_sig.max())
    report[f"{tag}_logit_maxdiff"] = float(d_log.max())
    assert float(d_sig.max()) <= bd["sigma"], f"{tag} {prec}: max |sigma - oracle| {float(d_sig.max()):.3e} > {bd['sigma']}"
    assert float(d_log.max()) <= bd["logit"], f"{tag} {prec}: max |colour logit - oracle| {float(d_log.max()):.3e} > {bd['logit']}"
    maps_o = _maps_of(raw_o, z, rd, cfg, mip, noise)
    s_g = raw_g[:, -1, 3] + (nz[:, -1] if nz is not None else 0.0)
    s_o = raw_o[:, -1, 3] + (nz[:, -1] if nz is not None else 0.0)
    # (ii): the 1e10-long last interval (not in mip mode, volume_rendering_utils.py:19) makes alpha_last a step in sigma_last
    step = ((s_g > 0) != (s_o > 0)) if not mip else torch.zeros_like(s_g, dtype=torch.bool)
    assert bool((s_o[step].abs() <= bd["sigma"]).all())     # follows from the sigma bound; kept as a tripwire
    sg_all = raw_g[..., 3] + (nz if nz is not None else 0.0)
    so_all = raw_o[..., 3] + (nz if nz is not None else 0.0)
    sign_flip = ((sg_all > 0) != (so_all > 0)).any(-1)
    err = _map_err(maps_g, maps_o, far)
    ok_disp = _disp_ok(maps_g, maps_o, B * 2, sign_flip, far)
    plain = ~step
    report[f"{tag}_rays"] = int(err.numel())
    report[f"{tag}_step_rays"] = int(step.sum())
    report[f"{tag}_unexplained_max"] = float(err[plain].max()) if plain.any() else 0.0
    bad = plain & ((err > B) | ~ok_disp)
    assert not bool(bad.any()), (f"{tag} {prec}: {int(bad.sum())} rays without a last-sample step exceed {B}: "
                                 f"max {float(err[plain].max()):.3e} (ray {int(torch.argmax(err * plain))}), disp ok "
                                 f"{bool(ok_disp[plain].all())}")
    if step.any():
        raw_h = raw_o.clone()
        raw_h[step, -1, 3] = raw_g[step, -1, 3]
        maps_h = _maps_of(raw_h, z, rd, cfg, mip, noise)
        err_h = _map_err(maps_g, maps_h, far)[step]
        ok_h = _disp_ok(maps_g, maps_h, B * 2, sign_flip, far)[step]
        report[f"{tag}_step_hybrid_max"] = float(err_h.max())
        report[f"{tag}_step_raw_max"] = float(err[step].max())
        assert float(err_h.max()) <= B and bool(ok_h.all()), \
            f"{tag} {prec}: a last-sample step does not explain its ray: hybrid error {float(err_h.max()):.3e}"
    return step, err


def check_chain(c, prec, out_g, tr_g, randoms=None):
    """Run every link for one case / precision.  `c`: dict as parity_cases.build_case returns (+ optional 'randoms'),
    `out_g` / `tr_g`: the 9-tuple and trace of the implementation under test.  Returns the report dict."""
    report = {"precision": prec}
    bd = BOUNDS[prec]
    cfg = c["opt"].nerf.validation
    mip = c.get("kind") == "mip"
    far = float(c["scfg"].far)
    randoms = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in (randoms or {}).items()}
    _, rd = _prep_rays(c)
    g = {k: v.detach().cpu() for k, v in tr_g.items()}
    out = [None if o is None else o.detach().cpu() for o in out_g[:6]]
    n = out[0].shape[0]
    Nc, Nf = int(cfg.num_coarse), int(cfg.num_fine)

    # ---- oracle, free-running
    tc = {}
    ref = _oracle(c, randoms, tc)
    assert torch.equal(g["z_coarse"], tc["z_coarse"]), "stratified depths must be bit-exact"
    # ---- L1 coarse
    maps_gc = dict(rgb=out[0], disp=out[1], acc=out[2], depth=g["depth_coarse"])
    step_c, err_c = _pass_link("coarse", prec, g["raw_coarse"], tc["raw_coarse"], maps_gc, tc["z_coarse"], rd, cfg, mip,
                               randoms.get("noise_c"), far, report)
    if Nf == 0:
        return report
    # ---- L2 resampling, given the implementation's own weights
    z = g["z_coarse"]
    mid = 0.5 * (z[..., 1:] + z[..., :-1])
    if mip:
        mid = 0.5 * (mid[..., 1:] + mid[..., :-1])
    nfs = Nf + (1 if mip else 0)
    det = cfg.perturb == 0.0
    u = randoms.get("u")
    smp, inds, cdf = O.sample_pdf(mid, g["weights_coarse"][..., 1:-1], nfs, det=det, u=u, return_all=True)
    uu = torch.linspace(0.0, 1.0, nfs) if u is None else u
    mism = H.check_resampling(g["inds"], g["z_samples"], inds, smp, cdf, mid, uu, f"L2 {prec}", max_flip_frac=0.01)
    report["resample_flips_given_own_weights"] = int(mism.sum())
    assert torch.equal(g["z_fine"], torch.sort(torch.cat([z, g["z_samples"]], -1), -1)[0]), "merged depths != sort(cat)"
    # ---- L3 fine pass, teacher-forced at the implementation's merged depths
    tf = {}
    _oracle(c, dict(randoms, z_fine=g["z_fine"]), tf)
    maps_gf = dict(rgb=out[3], disp=out[4], acc=out[5], depth=g["depth_fine"])
    step_f, err_f = _pass_link("fine_tf", prec, g["raw_fine"], tf["raw_fine"], maps_gf, g["z_fine"], rd, cfg, mip,
                               randoms.get("noise_f"), far, report)
    # ---- L4 free-running fine maps vs the oracle's own
    maps_of = dict(rgb=ref[3], disp=ref[4], acc=ref[5], depth=tc["depth_fine"])
    err_free = _map_err(maps_gf, maps_of, far)
    B_free = bd["B_free"] * (_interval_scale(tc["z_coarse"], rd, mip) if prec != "fp32" else 1.0)
    report["free_map_bound"] = B_free
    dz = (g["z_fine"] - tc["z_fine"]).abs().max(-1)[0]
    same_z = dz <= 1e-6 * max(far, 1.0)
    # the oracle's own fine pass may sit on a last-sample step too (its sigma_last vs the implementation's, at equal depth)
    s_of = tc["raw_fine"][:, -1, 3]
    s_gf = g["raw_fine"][:, -1, 3]
    nzf = randoms.get("noise_f")
    if nzf is not None and float(cfg.radiance_field_noise_std) > 0:
        s_of = s_of + nzf[:, -1] * float(cfg.radiance_field_noise_std)
        s_gf = s_gf + nzf[:, -1] * float(cfg.radiance_field_noise_std)
    step_free = (((s_gf > 0) != (s_of > 0)) & (not mip)) | step_f
    report["free_rays_same_depths"] = int(same_z.sum())
    report["free_step_rays"] = int(step_free.sum())
    if prec == "fp32":
        unexplained = same_z & ~step_free
        report["free_unexplained_max"] = float(err_free[unexplained].max()) if unexplained.any() else 0.0
        bad = unexplained & (err_free > bd["B_free"])     # fp32: the north-star bound itself, no interval scaling
        assert not bool(bad.any()), f"L4 fp32: {int(bad.sum())} rays with the oracle's depths and no step exceed {bd['B_free']}"
        # rays whose merged depths differ: the difference itself is what L2 attributes (flip / near-empty bin) once the
        # coarse weights agree to a few ulp, which L1 has bounded; state the agreement
        wdiff = float((g["weights_coarse"] - tc["weights_coarse"]).abs().max())
        report["coarse_weights_maxdiff"] = wdiff
        assert wdiff <= 1e-4
        report["free_max_incl_explained"] = float(err_free.max())
    else:
        # (i) in a 16-bit mode: the coarse weights differ from the oracle's by the mode's rounding (bounded in L1), so an
        # index flips wherever u lies within that ray's max |cdf - oracle cdf| of an oracle cdf edge — every flip is
        # checked against exactly that distance; the rays that hold one run their fine pass at other depths
        dcdf = (cdf - tc["cdf"]).abs().max(-1)[0]
        flips = g["inds"] != tc["inds"]
        uu2 = uu.expand_as(flips) if uu.dim() == 2 else uu[None].expand_as(flips)
        edge = (tc["cdf"][:, None, :] - uu2[:, :, None]).abs().min(-1)[0]
        assert bool((edge[flips] <= (dcdf[:, None].expand_as(flips)[flips] + 4.8e-7)).all()), \
            f"L4 {prec}: an index flip is farther from a cdf edge than the cdf moved"
        flip_rays = flips.any(-1)
        report["free_flip_rays"] = int(flip_rays.sum())
        plain = ~step_free & ~flip_rays
        # A ray without a flip or step still runs its fine pass at depths that moved CONTINUOUSLY with the cdf, and a depth
        # inside a near-empty bin is amplified by 1/denom (nerf_helpers.py:696-699): B_free states how far that normally
        # goes.  A ray above B_free is attributed constructively: the ORACLE's own fine network evaluated at the
        # implementation's merged depths (the L3 run) must reproduce the excess — i.e. the reference algorithm itself
        # moves that far when handed these depths, and what remains is within the teacher-forced bound B of L3.
        maps_tf = _maps_of(tf["raw_fine"], g["z_fine"], rd, cfg, mip, randoms.get("noise_f"))
        self_sens = _map_err(maps_tf, maps_of, far)
        B_tf = bd["B"] * (_interval_scale(g["z_fine"], rd, mip) if prec not in EXACT_MODES else 1.0)
        moved = plain & (err_free > B_free)
        report["free_depth_sensitivity_rays"] = int(moved.sum())
        unexpl = plain & ~moved
        report["free_unexplained_rays"] = int(unexpl.sum())
        report["free_unexplained_max"] = float(err_free[unexpl].max()) if unexpl.any() else 0.0
        bad = moved & (err_free > self_sens + B_tf)
        assert not bool(bad.any()), (f"L4 {prec}: {int(bad.sum())} rays without a flip or last-sample step exceed B_free "
                                     f"{B_free} by more than the oracle's own sensitivity to the depths: max "
                                     f"{float((err_free - self_sens)[moved].max()):.3e} vs {B_tf}")
        assert int(moved.sum()) <= max(2, n // 200), f"L4 {prec}: {int(moved.sum())} of {n} rays above B_free {B_free}"
        report["free_max_incl_explained"] = float(err_free.max())
    return report
