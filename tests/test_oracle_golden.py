"""CPU: the oracle (oracle/nvsr_oracle.py) against vectors produced by the reference itself
(tests/golden/make_golden.py).  On the machine that generated them the match is bit-exact; the
tolerances below only absorb CPU-to-CPU differences in ATen's vectorised reductions / libm."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import T, golden
from oracle import nvsr_oracle as O

TOL = 2e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_get_ray_bundle(tag):
    g = golden(f"stage_raybundle_{tag}.npz")
    f = g["focal"].tolist()
    ro, rd = O.get_ray_bundle(int(g["H"]), int(g["W"]), f, T(g["pose"]), int(g["pad"]), float(g["offset"]))
    H.assert_close(ro, g["ro"], 0, what="ro")
    H.assert_close(rd, g["rd"], TOL, what="rd")


def test_ndc_rays():
    g = golden("stage_ndc.npz")
    o, d = O.ndc_rays(int(g["H"]), int(g["W"]), float(g["focal"]), 1.0, T(g["ro"]), T(g["rd"]))
    H.assert_close(o, g["ro_ndc"], TOL, what="ro_ndc")
    H.assert_close(d, g["rd_ndc"], TOL, what="rd_ndc")


def test_planes_forward():
    g = golden("stage_planes_forward.npz")
    sid = str(g["scene_id"])
    mc, _ = H.load_planes_scene("scene_planes_small.npz", sid)
    mc.set_cur_scene_id(sid)
    with torch.no_grad():
        pos, view = O.planes_gather(mc, T(g["x6"]))
        out = O.planes_decode(mc, pos, view)
    for d in range(3):
        H.assert_close(pos[d], g[f"pos{d}"], TOL, what=f"pos{d}")
    H.assert_close(view, g["view"], TOL, what="view")
    H.assert_close(out, g["out"], 1e-4, 1e-5, what="decoder out")


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("white", dict(white_background=True)), ("mip", dict(mip_nerf=True))])
def test_volume_render(tag, kw):
    g = golden(f"stage_composite_{tag}.npz")
    o = O.volume_render_radiance_field(T(g["raw"]), T(g["z"]), T(g["rd"]), **kw)
    for name, v in zip(("rgb", "disp", "acc", "weights", "depth"), o):
        H.assert_close(v, g[name], 2e-6, 2e-6, what=name)
    assert np.isnan(g["disp"][0]) and not np.isnan(g["disp"][1:]).any()  # acc == 0 ray -> NaN disp


@pytest.mark.parametrize("tag", ["rand", "dyadic"])
def test_sample_pdf_det(tag):
    g = golden(f"stage_samplepdf_{tag}.npz")
    smp, inds, cdf = O.sample_pdf(T(g["bins"]), T(g["weights"]), 48, det=True, return_all=True)
    H.assert_close(cdf, g["cdf"], 2e-7, what="cdf")
    H.assert_close(smp, g["samples"], 2e-6, what="samples")
    # stage test: searching the reference's own cdf must give its indices bit-exactly
    s2, i2 = O.searchsorted_lerp(T(g["cdf"]), T(g["bins"]), T(g["u"])[None])
    assert torch.equal(i2, T(g["inds"]))
    if tag == "dyadic":
        assert torch.equal(inds, T(g["inds"]))


def test_sample_pdf_random_u():
    g = golden("stage_samplepdf_random_u.npz")
    smp = O.sample_pdf(T(g["bins"]), T(g["weights"]), 31, det=False, u=T(g["u"]))
    H.assert_close(smp, g["samples"], 2e-6, what="samples")


def test_ipe_and_mip_mlp():
    g = golden("stage_ipe.npz")
    means, covs = O.cast_rays(T(g["z"]), T(g["ro"]), T(g["rd"]), float(g["radius"]))
    H.assert_close(means, g["means"], TOL, what="means")
    H.assert_close(covs, g["covs"], 1e-7, 1e-5, what="covs")
    enc = O.integrated_pos_enc(means, covs, 7)
    H.assert_close(enc, g["enc"], TOL, what="ipe")
    H.assert_close(O.positional_encoding(T(g["viewdirs"]), 4, True), g["dir_enc"], TOL, what="dir enc")
    mc, _ = H.load_mip_scene(str(g["scene_file"]))
    with torch.no_grad():
        out = O.flexible_model_forward(mc, T(g["mlp_in"]))
    H.assert_close(out, g["mlp_out"], 1e-4, 1e-5, what="mip mlp")


E2E = ["e2e_planes_det.npz", "e2e_planes_perturb.npz", "e2e_planes_ndc.npz", "e2e_planes_coarse_only.npz",
       "e2e_planes_sr.npz", "e2e_mip_det.npz"]


def run_oracle_e2e(name, device="cpu", runner=None, z_fine=None, **kw):
    g = golden(name)
    sid = str(g["scene_id"])
    mip = name.startswith("e2e_mip")
    enc = encd = None
    if mip:
        mc, mf = H.load_mip_scene(str(g["scene_file"]), device)
    else:
        lr = str(g["lr_scene_id"]) if "lr_scene_id" in g else None
        mc, mf = H.load_planes_scene(str(g["scene_file"]), sid, device, lr_scene_id=lr)
    opt, scfg = H.options_from(g, mip), H.scene_cfg_from(g)
    batch = torch.stack([T(g["ro"], device).reshape(-1, 3), T(g["rd"], device).reshape(-1, 3)], 0)
    rnd = H.randoms_from(g, device)
    if z_fine is not None:
        rnd["z_fine"] = z_fine
    if runner is None:
        if mip:
            enc = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
            encd = lambda x: O.positional_encoding(x, 4, True)
        runner = O.run_one_iter_of_nerf
    else:
        if mip:
            import nvsr_b200
            enc = nvsr_b200.IntegratedPositionalEncoding(3, 7)
            encd = object()
    with torch.no_grad():
        out = runner(int(g["H"]), int(g["W"]), float(g["focal"]), mc, mf, batch, opt, sid, mode="validation",
                     encode_position_fn=enc, encode_direction_fn=encd, scene_config=scfg, randoms=rnd, **kw)
    return g, out


NAMES = ["rgb_coarse", "disp_coarse", "acc_coarse", "rgb_fine", "disp_fine", "acc_fine"]


@pytest.mark.parametrize("name", E2E)
def test_e2e(name):
    g, out = run_oracle_e2e(name)
    for k, v in zip(NAMES, out[:6]):
        if k not in g:
            assert v is None
            continue
        # disp = 1/(depth/acc) amplifies last-ulp differences when depth/acc is tiny: relative tolerance
        H.assert_close(v, g[k], 5e-6, 1e-4 if "disp" in k else 1e-5, what=f"{name}:{k}")
    assert out[6] is None and out[7] is None and out[8] is None


@pytest.mark.parametrize("name", ["sr_small.npz", "sr_x2_norm.npz", "sr_ragged.npz"])
def test_planes_sr_forward_vs_reference(name):
    """The oracle's restatement of PlanesSR.forward + EDSR (models.py:773-822, 884-926) against planes super-resolved by
    the reference's own classes (tests/golden/make_golden_sr.py)."""
    sr, g, planes = H.load_sr_model(name)
    assert int(sr.inner_model.required_padding) == int(g["required_padding"]) and int(sr.HR_overpadding) == int(g["hr_overpadding"])
    with torch.no_grad():
        for pname, want in planes.items():
            got = O.planes_sr_forward(sr, pname)
            H.assert_close(got, want, 2e-6, 1e-5, what=pname)
