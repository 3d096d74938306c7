"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (CPU, this container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference) is pure Python; it is imported here with the three shims of
SURVEY.md §8c.  It cannot travel to the GPU box, so its inputs/outputs are committed as small .npz
fixtures.  Models are built with the reference's OWN classes (models.TwoDimPlanesModel /
FlexibleNeRFModel / SceneCoupler / PlanesSR+EDSR, cfgnode.CfgNode, mip.IntegratedPositionalEncoding);
their weights are saved in the fixture so the tests can load them into the stand-ins of
neural-volume-super-resolution_b200/scene.py.
"""
import os
import sys
import types

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

REF = os.environ.get("NVSR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

# ---- shims (before importing the reference) ----
scipy.signal.gaussian = scipy.signal.windows.gaussian      # imresize.py:4
sys.modules.setdefault("imageio", types.ModuleType("imageio"))  # nerf_helpers.py:17
if not torch.cuda.is_available():
    torch.Tensor.cuda = lambda self, *a, **k: self         # models.py:284 hard-codes .cuda()
sys.path.insert(0, REF)
import mip  # noqa: E402
import models  # noqa: E402
import nerf_helpers  # noqa: E402
import train_utils  # noqa: E402
import volume_rendering_utils  # noqa: E402
from cfgnode import CfgNode  # noqa: E402

BOX = [[-1.5, -1.5, -1.5, -np.pi, -np.pi / 2], [1.5, 1.5, 1.5, np.pi, np.pi / 2]]
KW = dict(use_viewdirs=True, dec_density_layers=4, dec_rgb_layers=4, dec_channels=128, skip_connect_every=3,
          num_plane_channels=48, rgb_dec_input="projections", proj_combination="avg",
          viewdir_proj_combination="concat_pos", align_corners=True)  # config/TrainModels.yml:66-93


def pose_spherical(theta, phi, radius):  # load_blender.py:34-39 restated (load_blender needs `magic`)
    def tz(t):
        m = np.eye(4, dtype=np.float32); m[2, 3] = t; return m

    def rx(p):
        m = np.eye(4, dtype=np.float32); m[1, 1] = m[2, 2] = np.cos(p); m[1, 2] = -np.sin(p); m[2, 1] = np.sin(p); return m

    def ry(t):
        m = np.eye(4, dtype=np.float32); m[0, 0] = m[2, 2] = np.cos(t); m[0, 2] = -np.sin(t); m[2, 0] = np.sin(t); return m

    c2w = ry(theta / 180 * np.pi) @ (rx(phi / 180.0 * np.pi) @ tz(radius))
    return np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w


def shape_density(m, x, target_std=10.0, shift=-8.0):
    """Standardise the density head on the model's own outputs at inputs `x` so that the volume is
    not degenerate: raw sigma ~ N(shift, target_std^2)."""
    with torch.no_grad():
        y = m(x)[..., 3]
        a = target_std / float(y.std())
        head = m.fc_alpha["0"] if isinstance(m.fc_alpha, torch.nn.ModuleDict) else m.fc_alpha
        head.weight.mul_(a)
        head.bias.copy_(a * (head.bias - float(y.mean())) + shift)


def random_points(n, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * 3.0 - 1.5
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return torch.cat([pts, dirs], -1)


def build_planes_models(sid, res, vres, seed, lr_sid=None, lr_res=None):
    torch.manual_seed(seed); np.random.seed(seed)
    ids = [sid] if lr_sid is None else [lr_sid, sid]
    coupler = models.SceneCoupler(ids, planes_res="LR", num_pos_planes=3, training_scenes=ids)
    mc = models.TwoDimPlanesModel(num_planes_or_rot_mats=3, scene_coupler=coupler, **KW)
    mf = models.TwoDimPlanesModel(num_planes_or_rot_mats=mc.rot_mats(), scene_coupler=coupler, **KW)
    mc.optional_no_grad = nerf_helpers.null_with
    stored = sid if lr_sid is None else lr_sid
    r = res if lr_sid is None else lr_res
    planes = torch.nn.ParameterDict([(models.get_plane_name(stored, d), models.create_plane(r if d < 3 else vres, 48, 0.5))
                                     for d in range(4)])
    box = torch.tensor(BOX, dtype=torch.float64)
    for m in (mc, mf):
        m.planes_, m.plane_rank, m.generated_planes, m.downsampled_planes, m.coverages = planes, None, {}, {}, {}
        m.box_coords = {i: box for i in ids}
        m.eval()
    if lr_sid is None:
        for m in (mc, mf):
            m.set_cur_scene_id(sid)
            shape_density(m, random_points(4096, 99))
    return mc, mf, coupler


def options(num_coarse, num_fine, perturb=False, lindisp=False, white=False, noise=0.0, mip_enc=False, chunk=131072):
    sub = dict(chunksize=chunk, perturb=perturb, num_coarse=num_coarse, num_fine=num_fine, white_background=white,
               radiance_field_noise_std=noise, lindisp=lindisp)
    nerf = dict(use_viewdirs=True, train=dict(sub), validation=dict(sub))
    if mip_enc:
        nerf["encode_position_fn"] = "mip"
    return CfgNode(dict(nerf=nerf))


class Recorder:
    """Record the CPU RNG draws the reference makes (train_utils.py:108, nerf_helpers.py:683,
    volume_rendering_utils.py:32) so tests can feed the SAME numbers to the oracle and the kernels."""

    def __enter__(self):
        self.rand, self.randn = [], []
        self._r, self._n = torch.rand, torch.randn

        def rand(*a, **k):
            v = self._r(*a, **k); self.rand.append(v.clone()); return v

        def randn(*a, **k):
            v = self._n(*a, **k); self.randn.append(v.clone()); return v

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn = self._r, self._n


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if v is None:
            continue
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, len(out), "arrays")


def state_of(m, prefix):
    return {prefix + k.replace(".", "__"): v for k, v in m.state_dict().items()
            if "planes_" not in k and "rot_mats" not in k and "SR_model" not in k}


def save_scene(name, mc, mf, planes, extra=None):
    """planes + decoder weights, stored ONCE per scene and shared by the e2e cases that name it"""
    d = {}
    if planes is not None:
        for k, p in planes.items():
            d["plane__" + k] = p
    d.update(state_of(mc, "coarse__"))
    d.update(state_of(mf, "fine__"))
    d.update(extra or {})
    npz(name, **d)


def e2e_case(name, mc, mf, sid, scene_file, opt, scfg, H, W, focal, pose, offset=0.0, enc=None, encd=None, extra=None):
    with torch.no_grad():
        ro, rd = nerf_helpers.get_ray_bundle(H, W, focal, pose, downsampling_offset=offset)
        batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
        with Recorder() as rec:
            out = train_utils.run_one_iter_of_nerf(H, W, focal, mc, mf, batch, opt, sid, mode="validation",
                                                   encode_position_fn=enc, encode_direction_fn=encd, scene_config=scfg)
    d = dict(H=H, W=W, focal=focal, pose=pose, offset=offset, ro=ro, rd=rd, scene_id=np.array(sid),
             rgb_coarse=out[0], disp_coarse=out[1], acc_coarse=out[2], rgb_fine=out[3], disp_fine=out[4], acc_fine=out[5],
             near=scfg.near, far=scfg.far, no_ndc=scfg.no_ndc)
    v = opt.nerf.validation
    d.update(num_coarse=v.num_coarse, num_fine=v.num_fine, perturb=v.perturb, lindisp=v.lindisp,
             white_background=v.white_background, noise_std=v.radiance_field_noise_std)
    # RNG draw order inside predict_and_render_radiance (single ray batch): t_rand, [noise_c], u, [noise_f]
    if v.perturb:
        d["t_rand"] = rec.rand[0]
        if v.num_fine > 0:
            d["u"] = rec.rand[1]
    if v.radiance_field_noise_std > 0:
        d["noise_c"] = rec.randn[0]
        if v.num_fine > 0:
            d["noise_f"] = rec.randn[1]
    d["scene_file"] = np.array(scene_file)
    d.update(extra or {})
    npz(name, **d)


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    # ------------------------------------------------------------------ planes model, small scene
    sid = "synth_DS2_PlRes24_16"
    mc, mf, _ = build_planes_models(sid, 24, 16, seed=0)
    save_scene("scene_planes_small.npz", mc, mf, dict(mc.planes_.items()))
    planes = "scene_planes_small.npz"
    H = W = 12
    pose = torch.from_numpy(pose_spherical(30.0, -30.0, 4.0)).float()
    focal = float(0.5 * W / np.tan(0.5 * 0.6911112070083618))
    scfg = CfgNode(dict(near=2.0, far=6.0, no_ndc=True))
    e2e_case("e2e_planes_det.npz", mc, mf, sid, planes, options(16, 24), scfg, H, W, focal, pose)
    torch.manual_seed(123)
    e2e_case("e2e_planes_perturb.npz", mc, mf, sid, planes,
             options(16, 24, perturb=True, lindisp=True, white=True, noise=0.5), scfg, H, W, focal, pose, offset=0.25)
    scfg_ndc = CfgNode(dict(near=0.0, far=1.0, no_ndc=False))
    pose_ff = torch.eye(4); pose_ff[2, 3] = 0.3; pose_ff[0, 3] = 0.05
    e2e_case("e2e_planes_ndc.npz", mc, mf, sid, planes, options(16, 24), scfg_ndc, H, W, 0.8 * W, pose_ff)
    e2e_case("e2e_planes_coarse_only.npz", mc, mf, sid, planes, options(16, 0), scfg, H, W, focal, pose)

    # ------------------------------------------------------------------ stage vectors (same scene)
    with torch.no_grad():
        g = torch.Generator().manual_seed(7)
        # decoder forward: [n,6] points incl. out-of-box points (border clamp) and exact corners
        n = 257
        pts = (torch.rand(n, 3, generator=g) * 3.6 - 1.8)
        pts[:4] = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5], [0.0, 0.0, 0.0], [1.5, -1.5, 0.3]])
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        x6 = torch.cat([pts, dirs], -1)
        mc.set_cur_scene_id(sid)
        x5 = torch.cat([x6[..., :3], nerf_helpers.cart2az_el(x6[..., 3:])], -1)
        xn = mc.normalize_coords(x5)
        pos = mc.project_xyz(xn[..., :3])
        view = mc.project_viewdir(xn[..., 3:])
        y = mc(x6)
        npz("stage_planes_forward.npz", x6=x6, out=y, pos0=pos[0], pos1=pos[1], pos2=pos[2], view=view, scene_id=np.array(sid))

        # get_ray_bundle variants
        for tag, (h, w, f, off, pad) in dict(a=(7, 9, 11.5, 0.0, 0), b=(5, 6, [13.0, 12.0], 0.375, 0), c=(4, 5, 9.0, 0.0, 2)).items():
            ro, rd = nerf_helpers.get_ray_bundle(h, w, f, pose, padding_size=pad, downsampling_offset=off)
            npz("stage_raybundle_%s.npz" % tag, H=h, W=w, focal=np.array(f, dtype=np.float64), offset=off, pad=pad, pose=pose,
                ro=ro.contiguous(), rd=rd)
        # ndc_rays
        ro = torch.rand(50, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 0.5])
        rd = torch.cat([torch.rand(50, 2, generator=g) - 0.5, -torch.ones(50, 1)], -1)
        o2, d2 = nerf_helpers.ndc_rays(378, 504, 400.0, 1.0, ro, rd)
        npz("stage_ndc.npz", ro=ro, rd=rd, ro_ndc=o2, rd_ndc=d2, H=378, W=504, focal=400.0)

        # volume_render_radiance_field: incl. sigma<=0 rays (acc=0 -> disp NaN) and saturated rays
        N, S = 37, 40
        raw = torch.randn(N, S, 4, generator=g) * 2.0
        raw[0, :, 3] = -1.0          # empty ray: acc == 0, disp == NaN
        raw[1, :, 3] = 50.0          # opaque at first sample
        z = torch.sort(torch.rand(N, S, generator=g) * 4 + 2, -1)[0]
        rdd = torch.randn(N, 3, generator=g)
        for tag, kw in dict(plain={}, white=dict(white_background=True)).items():
            o = volume_rendering_utils.volume_render_radiance_field(raw, z, rdd, **kw)
            npz("stage_composite_%s.npz" % tag, raw=raw, z=z, rd=rdd, rgb=o[0], disp=o[1], acc=o[2], weights=o[3], depth=o[4])
        zz = torch.sort(torch.rand(N, S + 1, generator=g) * 4 + 2, -1)[0]
        o = volume_rendering_utils.volume_render_radiance_field(raw, zz, rdd, mip_nerf=True)
        npz("stage_composite_mip.npz", raw=raw, z=zz, rd=rdd, rgb=o[0], disp=o[1], acc=o[2], weights=o[3], depth=o[4])

        # sample_pdf_2: random weights, dyadic weights (order-independent sums), delta and flat pdfs
        B = 33
        bins = torch.sort(torch.rand(N, B, generator=g) * 4 + 2, -1)[0]
        w = torch.rand(N, B - 1, generator=g)
        w[0] = 0.0                                   # flat pdf from the 1e-5 floor
        w[1] = 0.0; w[1, 5] = 1.0                    # delta
        # dyadic: (k/64 - 1e-5) is not exactly representable, so use exact multiples and let +1e-5 round
        wd = torch.randint(0, 64, (N, B - 1), generator=g).float() / 64.0
        for tag, ww in dict(rand=w, dyadic=wd).items():
            for det in (True,):
                smp = nerf_helpers.sample_pdf_2(bins, ww, 48, det=det)
                # recompute the intermediates exactly as the function does, for the stage test
                w2 = ww + 1e-5
                pdf = w2 / torch.sum(w2, -1, keepdim=True)
                cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
                u = torch.linspace(0.0, 1.0, steps=48).expand(N, 48).contiguous()
                inds = torch.searchsorted(cdf.contiguous(), u, side="right")
                npz("stage_samplepdf_%s.npz" % tag, bins=bins, weights=ww, samples=smp, cdf=cdf, inds=inds, u=u[0])
        torch.manual_seed(5)
        with Recorder() as rec:
            smp = nerf_helpers.sample_pdf_2(bins, w, 31, det=False)
        npz("stage_samplepdf_random_u.npz", bins=bins, weights=w, samples=smp, u=rec.rand[0])

    # ------------------------------------------------------------------ mip / IPE + FlexibleNeRFModel
    torch.manual_seed(1)
    pe = lambda x: nerf_helpers.positional_encoding(x, 4, True)
    fc = models.FlexibleNeRFModel(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, include_input_xyz=False,
                                  include_input_dir=True, use_viewdirs=True)
    ff = models.FlexibleNeRFModel(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, include_input_xyz=False,
                                  include_input_dir=True, use_viewdirs=True)
    for m in (fc, ff):
        m.eval()
        shape_density(m, torch.cat([torch.sin(torch.rand(4096, 36) * 6.2831853), pe(random_points(4096, 98)[:, 3:])], -1))
    fc.optional_no_grad = nerf_helpers.null_with
    enc = mip.IntegratedPositionalEncoding(3, multires=7)
    save_scene("scene_mip_small.npz", fc, ff, None)
    e2e_case("e2e_mip_det.npz", fc, ff, "lego_DS2", "scene_mip_small.npz", options(16, 24, mip_enc=True), scfg, H, W, focal,
             pose, enc=enc, encd=pe)
    with torch.no_grad():
        g = torch.Generator().manual_seed(11)
        zed = torch.sort(torch.rand(9, 21, generator=g) * 4 + 2, -1)[0]
        o = torch.randn(9, 3, generator=g); d = torch.randn(9, 3, generator=g)
        radius = 2 * 0.00135 * 2 / np.sqrt(12.0)
        means, covs = mip.cast_rays(zed, o, d, radius, None)
        e = enc((means, covs))
        vd = torch.nn.functional.normalize(d, dim=-1)
        x = torch.cat([e.reshape(-1, 36), pe(vd[:, None, :].expand(9, 20, 3).reshape(-1, 3))], -1)
        npz("stage_ipe.npz", z=zed, ro=o, rd=d, radius=radius, means=means, covs=covs, enc=e, dir_enc=pe(vd), viewdirs=vd,
            mlp_in=x, mlp_out=fc(x), scene_file=np.array("scene_mip_small.npz"))

    # ------------------------------------------------------------------ SR planes (config 3a): real PlanesSR + EDSR
    lr, hr = "synth_DS8_PlRes8_16", "synth_DS2_PlRes32_16"
    mc2, mf2, coupler = build_planes_models(hr, 32, 16, seed=2, lr_sid=lr, lr_res=8)
    torch.manual_seed(3)
    sr = models.PlanesSR(model_arch=models.EDSR, scale_factor=4, in_channels=48, out_channels=48,
                         sr_config=CfgNode({"model": {"hidden_size": 16, "n_blocks": 2}}), plane_interp="bilinear").eval()
    mf2.assign_SR_model(sr, SR_viewdir=False)
    mf2.assign_LR_planes()
    for m in (mc2, mf2):   # calibrate after the SR model is attached: the fine model reads SR planes
        m.set_cur_scene_id(hr)
        shape_density(m, random_points(4096, 97))
    sr.clear_SR_planes()
    with torch.no_grad():
        e2e_case("e2e_planes_sr.npz", mc2, mf2, hr, "scene_planes_sr.npz", options(16, 24), scfg, H, W, focal, pose,
                 offset=(2 - 1) / (2 * 2), extra={"lr_scene_id": np.array(lr)})
        # SR_planes is filled by the render above (PlanesSR.forward caches its output, models.py:925)
        save_scene("scene_planes_sr.npz", mc2, mf2, dict(mc2.planes_.items()),
                   extra={"sr_plane__" + k: v for k, v in sr.SR_planes.items()})


if __name__ == "__main__":
    main()
