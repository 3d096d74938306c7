"""Backward kernels' arithmetic, checked on the CPU (no GPU needed).

csrc/backward_bodies.h holds the per-element bodies of the backward CUDA kernels (csrc/backward.cu) as
host/device-neutral source.  tests/hostcheck/hostcheck.cpp compiles that same source with g++ and loops over the
kernels' own index decomposition; here its output is compared with autograd through the oracle's restatement of the
reference forward (volume_rendering_utils.py:15-51, models.py:289-326,355-361) — the gradients the reference's
`loss.backward()` (train_nerf.py:905) produces.  Tolerances: 2e-5 relative to the largest gradient of the tensor
(fp32 accumulation order differs from ATen's).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import nvsr_b200  # noqa: F401  (puts the package on the path)
from nvsr_b200 import scene
from oracle import nvsr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "neural-volume-super-resolution_b200", "csrc")


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", out,
                    os.path.join(HERE, "hostcheck", "hostcheck.cpp")], check=True)
    return C.CDLL(out)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _close(got, want, rel=2e-5):
    scale = float(want.abs().max()) + 1e-30
    err = float((got - want).abs().max())
    assert err <= rel * scale, f"max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("white,noise_std,mip,extras", [(False, 0.0, False, False), (True, 0.0, False, True),
                                                          (False, 0.7, False, True), (True, 0.3, True, True)])
def test_composite_bwd_body_matches_autograd(hc, white, noise_std, mip, extras):
    g = torch.Generator().manual_seed(5 + int(white) + 2 * int(mip))
    n, S = 37, 29
    raw = torch.randn(n, S, 4, generator=g) * 1.5
    raw[..., 3] = raw[..., 3] * 4.0 - 1.0            # lit and unlit samples, some nearly opaque
    raw[3, 5:9, 3] = 60.0                            # opaque run: q_i = 1e-10 factors
    Z = S + (1 if mip else 0)
    z = torch.sort(2.0 + 4.0 * torch.rand(n, Z, generator=g), -1).values
    rd = torch.randn(n, 3, generator=g)
    noise = torch.randn(n, S, generator=g) if noise_std > 0 else None
    raw_req = raw.clone().requires_grad_(True)
    rgb, _, acc, w, depth = O.volume_render_radiance_field(raw_req, z, rd, noise_std, white, mip_nerf=mip, noise=noise)
    g_rgb = torch.randn(n, 3, generator=g)
    g_acc = torch.randn(n, generator=g) if extras else None
    g_depth = torch.randn(n, generator=g) if extras else None
    g_w = torch.randn(n, S, generator=g) if extras else None
    loss = (rgb * g_rgb).sum()
    if extras:
        loss = loss + (acc * g_acc).sum() + (depth * g_depth).sum() + (w * g_w).sum()
    loss.backward()
    d_raw = torch.full((n, S, 4), float("nan"))
    nz = None if noise is None else (noise * noise_std).contiguous()
    hc.hc_composite_bwd(_p(raw), _p(z), _p(rd), _p(nz), C.c_int64(n), S, int(white), int(mip), _p(g_rgb), _p(g_acc),
                        _p(g_depth), _p(g_w), _p(d_raw))
    assert torch.isfinite(d_raw).all()
    _close(d_raw[..., :3], raw_req.grad[..., :3])
    _close(d_raw[..., 3], raw_req.grad[..., 3], rel=1e-4)
    # a sample with sigma + noise <= 0 receives exactly zero density gradient (relu)
    pre = raw[..., 3] + (0 if nz is None else nz)
    assert (d_raw[..., 3][pre <= 0] == 0).all()


def _geometry(model, sid):
    box = model.box_coords[sid].detach().double().cpu()
    lo, rng = box[0].float(), (box[1] - box[0]).float()
    rots = model.coord_projector.rot_mats_NON_LEARNED
    proj = torch.stack([rots[d].detach().float().cpu()[:, 1:] for d in range(3)]).contiguous()   # [3][3][2]
    return lo, rng, proj


def test_gather_bwd_body_matches_autograd(hc):
    mc, mf, sid = scene.make_synthetic_scene(plane_res=12, view_res=6, channels=8, seed=3)
    model = mf
    model.set_cur_scene_id(sid)
    g = torch.Generator().manual_seed(11)
    n, S = 19, 7
    ro = torch.randn(n, 3, generator=g) * 0.3
    rd = torch.randn(n, 3, generator=g)
    rd[0] = torch.tensor([0.0, 0.0, 1.0])
    z = torch.sort(0.2 + 2.5 * torch.rand(n, S, generator=g), -1).values       # some points leave the box: border clamp
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
    x6 = torch.cat([pts, vd[:, None, :].expand(pts.shape)], -1).reshape(-1, 6)
    planes = [model.planes_[scene.get_plane_name(sid, d)] for d in range(4)]
    for p in planes:
        p.grad = None
    pos, view = O.planes_gather(model, x6)
    Cc = pos[0].shape[1]
    gp = torch.randn(n * S, 3 * Cc, generator=g)
    gp[5:9] = 0.0                                      # rows without gradient are skipped
    gm = torch.randn(n * S, Cc, generator=g)
    gv_row = torch.randn(n * S, Cc, generator=g)
    mean = torch.stack(pos, 0).mean(0)
    loss = (torch.cat(pos, 1) * gp).sum() + (mean * gm).sum() + (view * gv_row).sum()
    loss.backward()
    lo, rng, proj = _geometry(model, sid)
    lo3, rng3 = lo[:3].clone(), rng[:3].clone()
    R = planes[0].shape[-1]
    rh = (C.c_int * 3)(R, R, R)
    acc = [torch.zeros(R, R, Cc) for _ in range(3)]
    hc.hc_gather_bwd(rh, rh, Cc, _p(lo3), _p(rng3), _p(proj), _p(ro), _p(rd), _p(z),
                     C.c_int64(n), S, _p(gp), _p(gm), _p(acc[0]), _p(acc[1]), _p(acc[2]))
    for d in range(3):
        _close(acc[d].permute(2, 0, 1), planes[d].grad[0])
    # NULL halves: only d_feat_m / only d_feat_p
    only_m = [torch.zeros(R, R, Cc) for _ in range(3)]
    only_p = [torch.zeros(R, R, Cc) for _ in range(3)]
    hc.hc_gather_bwd(rh, rh, Cc, _p(lo3), _p(rng3), _p(proj), _p(ro), _p(rd), _p(z),
                     C.c_int64(n), S, None, _p(gm), _p(only_m[0]), _p(only_m[1]), _p(only_m[2]))
    hc.hc_gather_bwd(rh, rh, Cc, _p(lo3), _p(rng3), _p(proj), _p(ro), _p(rd), _p(z),
                     C.c_int64(n), S, _p(gp), None, _p(only_p[0]), _p(only_p[1]), _p(only_p[2]))
    for d in range(3):
        _close(only_m[d] + only_p[d], acc[d], rel=1e-5)
    # view plane: the per-sample gradient of a ray's view feature is summed per ray (the decoder broadcasts it)
    Rv = planes[3].shape[-1]
    gv = gv_row.reshape(n, S, Cc).sum(1).contiguous()
    vacc = torch.zeros(Rv, Rv, Cc)
    vd = vd.contiguous()
    hc.hc_viewdir_gather_bwd(_p(vd), C.c_int64(n), Rv, Rv, Cc, C.c_float(float(lo[3])), C.c_float(float(rng[3])),
                             C.c_float(float(lo[4])), C.c_float(float(rng[4])), _p(gv), _p(vacc))
    _close(vacc.permute(2, 0, 1), planes[3].grad[0])


def test_gather_bwd_mass_conservation(hc):
    """Size-independent property: the bilinear weights of a footprint sum to 1, so the total mass scattered into a plane
    equals the total feature gradient (exactly representable inputs: power-of-two gradients)."""
    g = torch.Generator().manual_seed(2)
    n, S, Cc, R = 64, 16, 4, 9
    ro = torch.zeros(n, 3)
    rd = torch.randn(n, 3, generator=g)
    z = torch.sort(torch.rand(n, S, generator=g), -1).values
    lo = torch.tensor([-1.5, -1.5, -1.5])
    rng = torch.tensor([3.0, 3.0, 3.0])
    proj = torch.tensor([[[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]], [[1.0, 0.0], [0.0, 0.0], [0.0, 1.0]],
                         [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]])
    gp = torch.ones(n * S, 3 * Cc) * 0.5
    rh = (C.c_int * 3)(R, R, R)
    acc = [torch.zeros(R, R, Cc) for _ in range(3)]
    hc.hc_gather_bwd(rh, rh, Cc, _p(lo), _p(rng), _p(proj), _p(ro), _p(rd), _p(z), C.c_int64(n), S, _p(gp), None,
                     _p(acc[0]), _p(acc[1]), _p(acc[2]))
    for d in range(3):
        np.testing.assert_allclose(float(acc[d].sum()), 0.5 * n * S * Cc, rtol=1e-5)


@pytest.fixture
def host_ops(hc, monkeypatch):
    """nvsr_b200.ops with the C-ABI calls replaced by the host stand-ins of tests/host_ops.py"""
    import host_ops as HO
    from nvsr_b200 import ops
    for name, fn in HO.standins(hc).items():
        monkeypatch.setattr(ops, name, fn)
    return ops


def test_autograd_module_plumbing_with_host_standins(host_ops):
    """The gradients of an mse loss on rgb_map through nvsr_b200.autograd must equal those of the oracle's pure-autograd
    pipeline: one decoder call + render, then the whole train-mode step."""
    from nvsr_b200 import autograd as A
    mc, model, sid = scene.make_synthetic_scene(plane_res=10, view_res=5, channels=8, seed=4)
    g = torch.Generator().manual_seed(8)
    n, S = 23, 9
    ro = torch.randn(n, 3, generator=g) * 0.2
    rd = torch.randn(n, 3, generator=g)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(0.3 + 2.0 * torch.rand(n, S, generator=g), -1).values
    target = torch.rand(n, 3, generator=g)
    noise = torch.randn(n, S, generator=g)
    params = [p for p in model.parameters() if p.requires_grad]

    def grads_of(loss):
        for p in params:
            p.grad = None
        loss.backward()
        return [None if p.grad is None else p.grad.clone() for p in params]

    # product plumbing
    rf = A.planes_model_forward(model, sid, ro, rd, z, vd)
    rgb, disp, acc, w, depth = A._VolumeRender.apply(rf, z, rd, (noise * 0.5).contiguous(), True, False)
    assert not disp.requires_grad
    got = grads_of(((rgb - target) ** 2).mean() + 0.1 * acc.mean())
    # oracle pipeline (pure torch autograd)
    model.set_cur_scene_id(sid)
    pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
    x6 = torch.cat([pts, vd[:, None, :].expand(pts.shape)], -1).reshape(-1, 6)
    rf_o = O.planes_model_forward(model, x6).reshape(n, S, 4)
    rgb_o, _, acc_o, _, _ = O.volume_render_radiance_field(rf_o, z, rd, 0.5, True, noise=noise)
    want = grads_of(((rgb_o - target) ** 2).mean() + 0.1 * acc_o.mean())
    assert torch.allclose(rf, rf_o, atol=1e-5)
    n_checked = 0
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if b is not None and float(b.abs().max()) > 0:
            _close(a, b, rel=2e-4)
            n_checked += 1
    assert n_checked >= 10   # 4 planes + decoder weights and biases

    # the whole train-mode step (train_nerf.py:860-905): coarse + fine with perturbation and density noise, loss on both
    # rgb maps; product composition (autograd._run_one_iter) against the oracle's run_one_iter_of_nerf
    opt = scene.render_options(9, 6, perturb=True, white_background=True, noise_std=0.4)
    scfg = scene.scene_cfg(0.3, 2.3, True)
    batch = torch.stack([ro, rd], 0)
    rnd = {"t_rand": torch.rand(n, 9, generator=g), "u": torch.rand(n, 6, generator=g),
           "noise_c": torch.randn(n, 9, generator=g), "noise_f": torch.randn(n, 15, generator=g)}
    all_params = list({id(p): p for m in (mc, model) for p in m.parameters() if p.requires_grad}.values())

    def step(fn):
        for p in all_params:
            p.grad = None
        out = fn()
        (((out[0] - target) ** 2).mean() + ((out[3] - target) ** 2).mean()).backward()
        return out, [None if p.grad is None else p.grad.clone() for p in all_params]

    out_p, got = step(lambda: A._run_one_iter(4, 4, 5.0, mc, model, batch, opt, sid, "train", scfg, rnd))
    out_o, want = step(lambda: O.run_one_iter_of_nerf(4, 4, 5.0, mc, model, batch, opt, sid, "train", scene_config=scfg, randoms=rnd))
    for j in (0, 2, 3, 5):
        assert torch.allclose(out_p[j], out_o[j], atol=2e-5), j
    assert out_p[6:] == (None, None, None)
    n_checked = 0
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if b is not None and float(b.abs().max()) > 0:
            _close(a, b, rel=5e-4)
            n_checked += 1
    assert n_checked >= 20


def _golden_step(run):
    """loss and gradients of the golden training step (tests/golden/backward_planes_train.npz, made by the reference)"""
    import helpers as H
    g = H.golden("backward_planes_train.npz")
    sid = str(g["scene_id"])
    mc, mf = H.load_planes_scene(str(g["scene_file"]), sid)
    opt, scfg, rnd = H.options_from(g), H.scene_cfg_from(g), H.randoms_from(g)
    batch = torch.stack([H.T(g["ro"]).reshape(-1, 3), H.T(g["rd"]).reshape(-1, 3)], 0)
    target = H.T(g["target"])
    named = {"plane__" + k: p for k, p in mc.planes_.items()}
    for prefix, m in (("coarse__", mc), ("fine__", mf)):
        for k, p in m.named_parameters():
            if "planes_" not in k and "rot_mats" not in k:
                named[prefix + k.replace(".", "__")] = p
    for p in named.values():
        p.grad = None
    out = run(int(g["H"]), int(g["W"]), float(g["focal"]), mc, mf, batch, opt, sid, scfg, rnd)
    loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
    loss.backward()
    return g, out, loss, named


def test_oracle_backward_matches_reference_gradients():
    """Pins the oracle's BACKWARD: autograd through the oracle's forward reproduces the gradients the reference's own
    loss.backward() produced (train_nerf.py:860-905; same ATen ops, so equal to rounding of the summation order)."""
    g, out, loss, named = _golden_step(lambda H_, W_, f, mc, mf, b, opt, sid, scfg, rnd: O.run_one_iter_of_nerf(
        H_, W_, f, mc, mf, b, opt, sid, "train", scene_config=scfg, randoms=rnd))
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-7
    assert torch.equal(out[0].detach(), torch.from_numpy(g["rgb_coarse"])) and torch.equal(out[3].detach(), torch.from_numpy(g["rgb_fine"]))
    keys = [k[len("grad__"):] for k in g if k.startswith("grad__")]
    assert len(keys) >= 20
    for k in keys:
        assert named[k].grad is not None, k
        _close(named[k].grad, torch.from_numpy(g["grad__" + k]), rel=1e-6)


def test_product_backward_composition_matches_reference_gradients(host_ops):
    """nvsr_b200.autograd's train-mode composition with the kernels' host-built bodies against the reference's gradients."""
    from nvsr_b200 import autograd as A
    g, out, loss, named = _golden_step(lambda H_, W_, f, mc, mf, b, opt, sid, scfg, rnd: A._run_one_iter(
        H_, W_, f, mc, mf, b, opt, sid, "train", scfg, rnd))
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5
    for k in [k[len("grad__"):] for k in g if k.startswith("grad__")]:
        assert named[k].grad is not None, k
        _close(named[k].grad, torch.from_numpy(g["grad__" + k]), rel=5e-4)


def test_mip_train_step_matches_oracle_autograd(host_ops):
    """mip/IPE family (FlexibleNeRFModel + IntegratedPositionalEncoding): train-mode step through
    nvsr_b200.autograd (IPE and direction encodings are data; decoder on torch; compositing with interval edges through
    the host-built backward body) against autograd of the oracle, on the golden mip scene."""
    import helpers as H
    from nvsr_b200 import autograd as A
    g = H.golden("e2e_mip_det.npz")
    sid = str(g["scene_id"])
    mc, mf = H.load_mip_scene(str(g["scene_file"]))
    opt = scene.render_options(int(g["num_coarse"]), int(g["num_fine"]), perturb=True, white_background=True, noise_std=0.3, mip=True)
    scfg = H.scene_cfg_from(g)
    batch = torch.stack([H.T(g["ro"]).reshape(-1, 3), H.T(g["rd"]).reshape(-1, 3)], 0)
    n, Nc, Nf = batch.shape[1], int(g["num_coarse"]), int(g["num_fine"])
    gen = torch.Generator().manual_seed(6)
    rnd = {"t_rand": torch.rand(n, Nc + 1, generator=gen), "u": torch.rand(n, Nf + 1, generator=gen),
           "noise_c": torch.randn(n, Nc, generator=gen), "noise_f": torch.randn(n, Nc + Nf + 1, generator=gen)}
    target = torch.rand(n, 3, generator=gen)
    params = [p for m in (mc, mf) for p in m.parameters()]

    def step(fn):
        for p in params:
            p.grad = None
        out = fn()
        (((out[0] - target) ** 2).mean() + ((out[3] - target) ** 2).mean()).backward()
        return out, [None if p.grad is None else p.grad.clone() for p in params]

    enc_o = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
    encd_o = lambda x: O.positional_encoding(x, 4, True)
    Hh, Ww, f = int(g["H"]), int(g["W"]), float(g["focal"])
    out_o, want = step(lambda: O.run_one_iter_of_nerf(Hh, Ww, f, mc, mf, batch, opt, sid, "train", encode_position_fn=enc_o,
                                                      encode_direction_fn=encd_o, scene_config=scfg, randoms=rnd))
    out_p, got = step(lambda: A._run_one_iter(Hh, Ww, f, mc, mf, batch, opt, sid, "train", scfg, rnd,
                                              nvsr_b200.IntegratedPositionalEncoding(3, 7)))
    for j in (0, 2, 3, 5):
        assert torch.allclose(out_p[j], out_o[j], atol=2e-5), j
    n_checked = 0
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if b is not None and float(b.abs().max()) > 0:
            _close(a, b, rel=5e-4)
            n_checked += 1
    assert n_checked >= 16


def test_sr_scene_gradients_reach_the_sr_network(host_ops):
    """Super-resolution configuration (BASELINE config 3a; models.py:296-310): the fine model reads its position planes
    through an SR network.  In training the gather's plane gradient must flow on into that network and into the LR planes
    (torch autograd behind the TriPlaneGather Function) — checked against autograd of the oracle with a small
    differentiable SR module."""
    from nvsr_b200 import autograd as A

    class TinySR(torch.nn.Module):                       # differentiable stand-in of PlanesSR.forward(plane_name)
        def __init__(self, planes, channels, scale):
            super().__init__()
            self.planes, self.scale = planes, scale
            self.conv = torch.nn.Conv2d(channels, channels, 3, padding=1)

        def forward(self, name):
            up = torch.nn.functional.interpolate(self.planes[name], scale_factor=self.scale, mode="bilinear", align_corners=True)
            return up + 0.1 * self.conv(up)

    mc, mf, sid = scene.make_synthetic_scene(plane_res=6, view_res=5, channels=8, seed=2, sr_scale=2)
    torch.manual_seed(0)
    mf.assign_SR_model(TinySR(mf.planes_, 8, 2))
    g = torch.Generator().manual_seed(3)
    n, S = 17, 6
    ro = torch.randn(n, 3, generator=g) * 0.2
    rd = torch.randn(n, 3, generator=g)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(0.3 + 2.0 * torch.rand(n, S, generator=g), -1).values
    gout = torch.randn(n, S, 4, generator=g)
    params = list(mf.planes_.values()) + list(mf.SR_model.conv.parameters())

    def grads(rf):
        for p in params:
            p.grad = None
        (rf * gout).sum().backward()
        return [p.grad.clone() for p in params]

    got = grads(A.planes_model_forward(mf, sid, ro, rd, z, vd))
    mf.set_cur_scene_id(sid)
    pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
    x6 = torch.cat([pts, vd[:, None, :].expand(pts.shape)], -1).reshape(-1, 6)
    want = grads(O.planes_model_forward(mf, x6).reshape(n, S, 4))
    assert len(got) == 6 and all(float(w.abs().max()) > 0 for w in want)   # 3 LR planes + view plane + conv weight/bias
    for a, b in zip(got, want):
        _close(a, b, rel=2e-4)


def test_coarse_pass_honours_optional_no_grad(host_ops):
    """train_utils.py:88 / train_nerf.py:560: with `model_coarse.optional_no_grad = torch.no_grad` the coarse maps carry
    no graph (the coarse decoder receives no gradient), the fine pass still does."""
    from nvsr_b200 import autograd as A
    mc, mf, sid = scene.make_synthetic_scene(plane_res=8, view_res=4, channels=8, seed=5)
    g = torch.Generator().manual_seed(1)
    batch = torch.stack([torch.randn(9, 3, generator=g) * 0.2, torch.randn(9, 3, generator=g)], 0)
    opt, scfg = scene.render_options(6, 4), scene.scene_cfg(0.3, 2.3, True)
    mc.optional_no_grad = torch.no_grad
    out = A._run_one_iter(3, 3, 4.0, mc, mf, batch, opt, sid, "train", scfg, None)
    assert not out[0].requires_grad and out[3].requires_grad
    out[3].sum().backward()
    assert all(p.grad is None for p in mc.density_dec.parameters()) and any(p.grad is not None for p in mf.rgb_dec.parameters())
    mc.optional_no_grad = __import__("contextlib").nullcontext
    assert A._run_one_iter(3, 3, 4.0, mc, mf, batch, opt, sid, "train", scfg, None)[0].requires_grad


def test_fast_frozen_coarse_route_plumbing(host_ops, monkeypatch):
    """Control flow of `autograd.set_fast_frozen_coarse(True)` on the CPU: the forward-kernel chunk renderer is replaced by
    a stand-in that computes the same coarse pass with the oracle, so the fast route must reproduce the default route
    exactly (same draws handed over, merged depths used for the fine pass, coarse maps without a graph)."""
    from nvsr_b200 import autograd as A, render
    mc, mf, sid = scene.make_synthetic_scene(plane_res=8, view_res=4, channels=8, seed=6)
    mc.optional_no_grad = torch.no_grad
    g = torch.Generator().manual_seed(2)
    n, Nc, Nf = 11, 7, 5
    batch = torch.stack([torch.randn(n, 3, generator=g) * 0.2, torch.randn(n, 3, generator=g)], 0)
    opt, scfg = scene.render_options(Nc, Nf, perturb=True, noise_std=0.3, white_background=True), scene.scene_cfg(0.3, 2.3, True)
    rnd = {"t_rand": torch.rand(n, Nc, generator=g), "u": torch.rand(n, Nf, generator=g),
           "noise_c": torch.randn(n, Nc, generator=g), "noise_f": torch.randn(n, Nc + Nf, generator=g)}
    seen = {}

    def fake_chunk(pc, pf, ro, rd, vd, near, far, cfg, randoms, trace, coarse_only=False):
        assert pc == "pc" and pf is None and coarse_only and set(randoms) == {"t_rand", "u", "noise_c"}
        seen["called"] = True
        tr = {}
        rays = torch.cat([ro, rd, torch.full((n, 1), near), torch.full((n, 1), far), vd], -1)
        out = O.predict_and_render_radiance(rays, mc, mf, opt, sid, mode="train", randoms=dict(randoms, noise_f=rnd["noise_f"]), trace=tr)
        return {"rgb": out[0], "disp": out[1], "acc": out[2], "z_merged": tr["z_fine"]}, None

    monkeypatch.setattr(render, "_planes_pass", lambda *a, **k: "pc")
    monkeypatch.setattr(render, "_render_planes_chunk", fake_chunk)
    slow = A._run_one_iter(3, 3, 4.0, mc, mf, batch, opt, sid, "train", scfg, rnd)
    A.set_fast_frozen_coarse(True)
    try:
        fast = A._run_one_iter(3, 3, 4.0, mc, mf, batch, opt, sid, "train", scfg, rnd)
    finally:
        A.set_fast_frozen_coarse(False)
    assert seen.get("called") and not fast[0].requires_grad and fast[3].requires_grad
    for j in (0, 2, 3, 5):
        assert torch.allclose(fast[j], slow[j], atol=2e-6), j


@pytest.mark.parametrize("n,S,mip", [(1, 1, False), (3, 2, False), (2, 1, True), (0, 5, False)])
def test_composite_bwd_body_degenerate_sizes(hc, n, S, mip):
    """A single sample per ray (only the 1e10 tail interval / one mip interval), two samples, and an empty ray set."""
    g = torch.Generator().manual_seed(n * 10 + S)
    raw = torch.randn(n, S, 4, generator=g)
    raw[..., 3] = raw[..., 3].abs() * 0.3 + 0.05 if n else raw[..., 3]
    z = torch.sort(2.0 + torch.rand(n, S + int(mip), generator=g), -1).values
    rd = torch.randn(n, 3, generator=g)
    g_rgb = torch.randn(n, 3, generator=g)
    d_raw = torch.full((n, S, 4), 7.0)
    hc.hc_composite_bwd(_p(raw), _p(z), _p(rd), None, C.c_int64(n), S, 0, int(mip), _p(g_rgb), None, None, None, _p(d_raw))
    if n == 0:
        return
    raw_req = raw.clone().requires_grad_(True)
    rgb = O.volume_render_radiance_field(raw_req, z, rd, 0.0, False, mip_nerf=mip)[0]
    (rgb * g_rgb).sum().backward()
    _close(d_raw[..., :3], raw_req.grad[..., :3])
    assert torch.allclose(d_raw[..., 3], raw_req.grad[..., 3], rtol=1e-4, atol=1e-7 * float(raw_req.grad.abs().max() + 1))
