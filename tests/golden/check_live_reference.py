"""Oracle vs the LIVE reference on fresh seeded inputs (only where the reference checkout exists: the build
container).  Run as a subprocess by tests/test_oracle_live_reference.py — importing the reference needs the three
shims of SURVEY.md §8c, one of which patches torch.Tensor.cuda, so it stays out of the pytest process.

Prints one JSON object {check: max abs difference}; the stage functions are the same ATen op sequences, so the
differences must be exactly 0 on one machine.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402  (applies the shims and imports the reference modules)
from oracle import nvsr_oracle as O  # noqa: E402

out = {}


def mad(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    both_nan = torch.isnan(a) & torch.isnan(b)
    return float(torch.where(both_nan, torch.zeros_like(a), (a - b).abs()).max())


with torch.no_grad():
    for seed in (101, 102, 103):
        g = torch.Generator().manual_seed(seed)
        # a1 get_ray_bundle (nerf_helpers.py:507-549), with padding and downsampling offset
        pose = torch.from_numpy(G.pose_spherical(17.0 * seed % 360, -30.0, 4.0)).float()
        for (h, w, f, pad, off) in ((37, 53, 61.5, 0, 0.0), (16, 24, [40.0, 44.0], 2, 0.25)):
            ro_r, rd_r = G.nerf_helpers.get_ray_bundle(h, w, f, pose, pad, off)
            ro_o, rd_o = O.get_ray_bundle(h, w, f, pose, pad, off)
            out[f"ray_bundle_{seed}_{h}"] = max(mad(ro_r, ro_o), mad(rd_r, rd_o))
        # a3 ndc_rays (nerf_helpers.py:578-605)
        ro = torch.randn(257, 3, generator=g); rd = torch.randn(257, 3, generator=g); rd[:, 2] -= 2.0
        a, b = G.nerf_helpers.ndc_rays(48, 64, 50.0, 1.0, ro, rd)
        c, d = O.ndc_rays(48, 64, 50.0, 1.0, ro, rd)
        out[f"ndc_{seed}"] = max(mad(a, c), mad(b, d))
        # a7 volume_render_radiance_field (volume_rendering_utils.py:6-51): plain / white background / mip
        n, S = 193, 64
        raw = torch.randn(n, S, 4, generator=g) * 3
        raw[: n // 4, :, 3] = -5.0                                  # empty rays: acc = 0, disp = NaN
        z = torch.sort(torch.rand(n, S, generator=g) * 4 + 2, -1)[0]
        dirs = torch.randn(n, 3, generator=g)
        for kw in ({}, {"white_background": True}):
            r = G.volume_rendering_utils.volume_render_radiance_field(raw, z, dirs, **kw)
            o = O.volume_render_radiance_field(raw, z, dirs, **kw)
            out[f"composite_{seed}_{'white' if kw else 'plain'}"] = max(mad(x, y) for x, y in zip(r, o))
            assert bool(torch.equal(torch.isnan(r[1]), torch.isnan(o[1])))
        ze = torch.sort(torch.rand(n, S + 1, generator=g) * 4 + 2, -1)[0]
        r = G.volume_rendering_utils.volume_render_radiance_field(raw, ze, dirs, mip_nerf=True)
        o = O.volume_render_radiance_field(raw, ze, dirs, mip_nerf=True)
        out[f"composite_{seed}_mip"] = max(mad(x, y) for x, y in zip(r, o))
        # a8 sample_pdf_2 (nerf_helpers.py:668-702), deterministic u: samples must be bit-identical
        bins = torch.sort(torch.rand(n, S - 1, generator=g) * 4 + 2, -1)[0]
        wts = torch.rand(n, S - 2, generator=g) ** 4
        out[f"sample_pdf_{seed}"] = mad(G.nerf_helpers.sample_pdf_2(bins, wts, 128, det=True), O.sample_pdf(bins, wts, 128, det=True))
        # a9 cast_rays + IntegratedPositionalEncoding (mip.py:9-43,154-199)
        enc = G.mip.IntegratedPositionalEncoding(3, 7)
        mc_r = G.mip.cast_rays(ze, ro[:n], rd[:n], torch.full((n, 1), 0.00156), None)
        mc_o = O.cast_rays(ze, ro[:n], rd[:n], torch.full((n, 1), 0.00156))
        out[f"cast_rays_{seed}"] = max(mad(mc_r[0], mc_o[0]), mad(mc_r[1], mc_o[1]))
        out[f"ipe_{seed}"] = mad(enc(mc_r), O.integrated_pos_enc(mc_o[0], mc_o[1], 7))
        out[f"dir_enc_{seed}"] = mad(G.nerf_helpers.positional_encoding(dirs, 4, True), O.positional_encoding(dirs, 4, True))
    # a5 + a6: TwoDimPlanesModel.forward (models.py:381-421) on the reference's own model objects
    sid = "live_DS2_PlRes24_8"
    mc, mf, _ = G.build_planes_models(sid, 24, 8, seed=5)
    x6 = G.random_points(1500, 7)
    mc.set_cur_scene_id(sid)
    out["planes_model_forward"] = mad(mc(x6), O.planes_model_forward(mc, x6))
    # the whole path (train_utils.py:185-282): coarse + fine, 9-tuple
    opt = G.options(32, 48)
    scfg = G.CfgNode(dict(near=2.0, far=6.0, no_ndc=True))
    pose = torch.from_numpy(G.pose_spherical(40.0, -30.0, 4.0)).float()
    ro, rd = G.nerf_helpers.get_ray_bundle(12, 12, 15.0, pose)
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    r = G.train_utils.run_one_iter_of_nerf(12, 12, 15.0, mc, mf, batch, opt, sid, mode="validation", scene_config=scfg)
    o = O.run_one_iter_of_nerf(12, 12, 15.0, mc, mf, batch, opt, sid, mode="validation", scene_config=scfg)
    out["run_one_iter_of_nerf"] = max(mad(x, y) for x, y in zip(r[:6], o[:6]))

print("LIVE_REFERENCE_JSON " + json.dumps(out))
