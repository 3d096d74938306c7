"""GPU parity of the decoder TRAINING path on the tensor cores (csrc/train_tc.cu and the TRAIN
variant of the forward chain): every kernel against a plain PyTorch restatement with the SAME 16-bit rounding points,
then the whole differentiable render step (`autograd.set_decoder('tc')`, the default) against the fp32 parity mode of the
same step (`'fp32'`: the model's own nn.Linear layers under torch autograd, which tests/test_gpu_next_rows.py pins to
the reference's golden gradients)."""
import pytest
import torch

import nvsr_b200
from nvsr_b200 import autograd as A, ops, scene
from nvsr_b200._lib import NVSR_F16

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def img_to_rows(img):
    """tile image [tiles, C/8, 128, 8] -> [tiles*128, C] fp32"""
    t, c8, r, e = img.shape
    return img.permute(0, 2, 1, 3).reshape(t * r, c8 * e).float()


def rows_to_img(x):
    """[tiles*128, C] -> fp16 tile image"""
    rows, c = x.shape
    return x.half().reshape(rows // 128, 128, c // 8, 8).permute(0, 2, 1, 3).contiguous()


def _chain(seed, k0, head_n, rb=None, n_rays=64, S=40):
    """random fp16-representable chain k0 -> 128 x4 -> head_n and a random feature image in the BLOCKED layout"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    rows = ops.rows_padded(n_rays, S, ops.ROWS_BLOCKED)
    W = [(torch.randn(128, k0 if l == 0 else 128, generator=g) * (1.5 / (k0 if l == 0 else 128) ** 0.5)).half().float().to(DEV)
         for l in range(4)]
    B = [(torch.randn(128, generator=g) * 0.3).to(DEV) for _ in range(4)]
    hw = (torch.randn(head_n, 128, generator=g) * 0.2).to(DEV)
    hb = torch.randn(head_n, generator=g).to(DEV)
    x0 = (torch.randn(rows, k0, generator=g) * 0.7).half().float().to(DEV)
    return W, B, hw, hb, x0, rows


def _layers(W, B, hw, hb, head_ch, row_bias=None):
    wimg = [ops.pack_weight16(w, dtype=NVSR_F16) for w in W]
    L = [ops.ChainLayer(wimg[l], None if (l == 0 and row_bias is not None) else B[l].contiguous(), W[l].shape[1], 128, True,
                        row_bias=row_bias if l == 0 else None, head_w=hw.contiguous() if l == 3 else None,
                        head_b=hb.contiguous() if l == 3 else None, head_ch=head_ch) for l in range(4)]
    return wimg, L


@pytest.mark.parametrize("k0,head_n,head_ch,per_ray,n", [(48, 1, 3, False, 64), (144, 3, 0, True, 64), (32, 1, 3, False, 64),
                                                          (144, 3, 0, True, 61)])
def test_train_forward_activations(k0, head_n, head_ch, per_ray, n):
    S = 40       # n = 61: padding rays in the last ray block — their bias row is zero, their activations finite
    W, B, hw, hb, x0, rows = _chain(1, k0, head_n, n_rays=n, S=S)
    rb = None
    if per_ray:
        rb = (torch.randn(n, 128, generator=torch.Generator().manual_seed(5)) * 0.5).to(DEV).contiguous()
    wimg, L = _layers(W, B, hw, hb, head_ch, rb)
    raw = ops.raw_buffer(n, S, ops.ROWS_BLOCKED, DEV).zero_()
    raw_inf = raw.clone()
    feat = rows_to_img(x0)
    acts = ops.mlp_chain_train(feat, L, rows, raw, S, n)
    ops.mlp_chain(feat, L, rows, raw_inf, NVSR_F16, S, n, ops.ROWS_BLOCKED)
    torch.cuda.synchronize()
    # the training forward IS the inference chain: identical heads, bit for bit
    assert torch.equal(raw[head_ch:head_ch + head_n], raw_inf[head_ch:head_ch + head_n])
    # torch restatement with the same rounding points (row -> ray through the BLOCKED order for the per-ray bias)
    ts = -(-S // 16)
    r = torch.arange(rows, device=DEV)
    ray = (r // 128 // ts) * 8 + (r % 8)
    x = x0
    for l in range(4):
        b = rb[ray.clamp(max=n - 1)] * (ray < n)[:, None] if (l == 0 and per_ray) else B[l]
        h = torch.relu(x @ W[l].t() + b)
        got = img_to_rows(acts[l])
        want = h.half().float()
        d = (got - want).abs()
        # a different summation order moves a value by fp32 noise, which may flip one fp16 rounding (1 ulp = 2^-10 rel)
        assert float(d.max()) <= 2.5e-3 * float(want.abs().max()), (l, float(d.max()))
        assert float((d > 1e-6).float().mean()) < 0.02, (l, float((d > 1e-6).float().mean()))
        x = got      # continue from the kernel's own operand, as the kernel does
    head = x @ hw.t() + hb        # (the kernel's heads read the unrounded activations: looser)
    got_h = raw[head_ch:head_ch + head_n, :rows].t()
    valid = ray < n
    assert float((got_h - head).abs()[valid].max()) <= 5e-3 * max(1.0, float(head.abs().max()))


@pytest.mark.parametrize("k0,head_n,head_ch", [(48, 1, 3), (144, 3, 0)])
def test_dgrad_chain_vs_torch(k0, head_n, head_ch):
    n, S = 61, 40          # ragged: padding rays and padding samples inside the last tiles
    W, B, hw, hb, x0, rows = _chain(2, k0, head_n, n_rays=n, S=S)
    # (the rgb chain's first-layer bias is per ray in the forward; the backward does not depend on it)
    rb = (torch.randn(n, 128, generator=torch.Generator().manual_seed(6)) * 0.5).to(DEV).contiguous() if head_n == 3 else None
    wimg, L = _layers(W, B, hw, hb, head_ch, rb)
    raw = ops.raw_buffer(n, S, ops.ROWS_BLOCKED, DEV).zero_()
    acts = ops.mlp_chain_train(rows_to_img(x0), L, rows, raw, S, n)
    g = torch.Generator().manual_seed(9)
    d_rf = torch.zeros(n, S, 4)
    d_rf[..., head_ch:head_ch + head_n] = torch.randn(n, S, head_n, generator=g) * 1e-4      # mse-sized gradients
    d_rf = d_rf.to(DEV)
    d_raw = ops.nsc_to_planar_blocked(d_rf, n, S)
    scale = 1024.0
    gi, dout, d_x0 = ops.mlp_dgrad(wimg, k0, hw, head_ch, d_raw, scale, acts, n, S)
    torch.cuda.synchronize()
    # torch restatement, same rounding points
    X = [img_to_rows(a) for a in acts]                       # x_1 .. x_4
    dsc = (d_raw[head_ch:head_ch + head_n, :rows].t() * scale)
    gl = ((dsc @ hw) * (X[3] > 0)).half().float()
    want = {3: gl}
    for l in (3, 2, 1):
        gl = ((gl @ W[l]) * (X[l - 1] > 0)).half().float()
        want[l - 1] = gl
    for l in range(4):
        got = img_to_rows(gi[l])
        d = (got - want[l]).abs()
        assert float(d.max()) <= 2.5e-3 * float(want[l].abs().max()) + 1e-7, (l, float(d.max()), float(want[l].abs().max()))
    # head gradient image: columns >= head_n zero
    do = img_to_rows(dout)
    assert torch.equal(do[:, head_n:], torch.zeros_like(do[:, head_n:]))
    assert float((do[:, :head_n] - dsc.half().float()).abs().max()) == 0.0
    # d_x0 from the kernel's own g_0, ray-major rows, unscaled
    ts = -(-S // 16)
    r = torch.arange(rows, device=DEV)
    ray, s = (r // 128 // ts) * 8 + (r % 8), (r // 128 % ts) * 16 + (r % 128) // 8
    valid = (ray < n) & (s < S)
    dx = (img_to_rows(gi[0]) @ W[0]) / scale
    ref = torch.zeros(n * S, k0, device=DEV)
    ref[(ray * S + s)[valid]] = dx[valid]
    assert float((d_x0 - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-12


@pytest.mark.parametrize("k0,head_n,head_ch,frac", [(48, 1, 3, 0.25), (144, 3, 0, 0.6), (32, 1, 3, 0.004)])
def test_dgrad_row_list_equals_dense_rows(k0, head_n, head_ch, frac):
    """nvsr_mlp_dgrad in row-list mode (row_ids): every listed row gets bit for bit the deltas / d_x0 of the dense pass,
    the LIST-ordered activation / feature images are the listed rows of the forward's images, the list tail is zero;
    and the weight gradients over the list equal the dense ones to the order of the fp32 sums."""
    n, S = 61, 40
    W, B, hw, hb, x0, rows = _chain(4, k0, head_n, n_rays=n, S=S)
    rb = (torch.randn(n, 128, generator=torch.Generator().manual_seed(6)) * 0.5).to(DEV).contiguous() if head_n == 3 else None
    wimg, L = _layers(W, B, hw, hb, head_ch, rb)
    raw = ops.raw_buffer(n, S, ops.ROWS_BLOCKED, DEV).zero_()
    x0_img = rows_to_img(x0)
    acts = ops.mlp_chain_train(x0_img, L, rows, raw, S, n)
    g = torch.Generator().manual_seed(10)
    d_rf = torch.zeros(n, S, 4)
    d_rf[..., head_ch:head_ch + head_n] = torch.randn(n, S, head_n, generator=g) * 1e-4 * (torch.rand(n, S, 1, generator=g) < frac)
    d_raw = ops.nsc_to_planar_blocked(d_rf.to(DEV), n, S)
    scale = 1024.0
    gd, doutd, dxd = ops.mlp_dgrad(wimg, k0, hw, head_ch, d_raw, scale, acts, n, S)
    ids, count = ops.nonzero_rows(d_raw)
    gl, doutl, dxl, acts_l, x0_l = ops.mlp_dgrad(wimg, k0, hw, head_ch, d_raw, scale, acts, n, S, row_count=count, row_ids=ids,
                                                 x0_img=x0_img)
    k = int(count)
    assert 0 < k < n * S
    kp = -(-k // 128) * 128
    sel = ids[:k].long()
    for l in range(4):
        assert torch.equal(img_to_rows(gl[l])[:k], img_to_rows(gd[l])[sel]), l
        assert not img_to_rows(gl[l])[k:kp].any()
        assert torch.equal(img_to_rows(acts_l[l])[:k], img_to_rows(acts[l])[sel]) and not img_to_rows(acts_l[l])[k:kp].any()
    assert torch.equal(img_to_rows(x0_l)[:k], x0[sel]) and not img_to_rows(x0_l)[k:kp].any()
    assert torch.equal(img_to_rows(doutl)[:k], img_to_rows(doutd)[sel])
    ts = -(-S // 16)
    ray, smp = (sel // 128 // ts) * 8 + (sel % 8), (sel // 128 % ts) * 16 + (sel % 128) // 8
    assert torch.equal(dxl[:k], dxd[ray * S + smp]) and not dxl[k:kp].any()
    # weight gradients: the list's products == the dense products
    def wgrads(gi, x0i, ai, do, rc):
        dws = [torch.zeros(128, k0 if l == 0 else 128, device=DEV) for l in range(4)]
        dbs = [torch.zeros(128, device=DEV) for _ in range(4)]
        dwh = torch.zeros(128, 16, device=DEV)
        ops.mlp_wgrad_chain(gi, x0i, k0, ai, do, 1.0 / scale, dws, dbs, dwh, row_count=rc)
        return dws + dbs + [dwh]
    for a, b in zip(wgrads(gl, x0_l, acts_l, doutl, count), wgrads(gd, x0_img, acts, doutd, None)):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12


@pytest.mark.parametrize("n_b", [16, 48, 128, 144])
def test_wgrad_and_ray_sum_vs_torch(n_b):
    g = torch.Generator().manual_seed(n_b)
    n, S = 300, 70
    rows = ops.rows_padded(n, S, ops.ROWS_BLOCKED)
    a = (torch.randn(rows, 128, generator=g) * 0.05).to(DEV)
    a[torch.rand(rows, 128, generator=g).to(DEV) < 0.5] = 0.0                 # masked deltas
    b = torch.relu(torch.randn(rows, n_b, generator=g)).to(DEV)
    ai, bi = rows_to_img(a), rows_to_img(b)
    dw = torch.zeros(128, n_b, device=DEV)
    db = torch.zeros(128, device=DEV)
    inv = 1.0 / 1024.0
    ops.mlp_wgrad(ai, bi, n_b, inv, dw, db)
    af, bf = img_to_rows(ai).double(), img_to_rows(bi).double()
    want = (af.t() @ bf) * inv
    assert float((dw.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    wantb = af.sum(0) * inv
    assert float((db.double() - wantb).abs().max()) <= 2e-5 * float(wantb.abs().max())
    # accumulation: a second call adds
    ops.mlp_wgrad(ai, bi, n_b, inv, dw, None)
    assert float((dw.double() - 2 * want).abs().max()) <= 4e-5 * float(want.abs().max())
    # per-ray sums of the 128-channel image
    rs = ops.ray_sum(ai, n, S, inv)
    ts = -(-S // 16)
    r = torch.arange(rows, device=DEV)
    ray, s = (r // 128 // ts) * 8 + (r % 8), (r // 128 % ts) * 16 + (r % 128) // 8
    valid = (ray < n) & (s < S)
    ref = torch.zeros(n, 128, device=DEV, dtype=torch.float64).index_add_(0, ray[valid], af[valid]) * inv
    assert float((rs.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def _step(mode, mc, mf, sid, batch, opt, scfg, rnd, target, H, W, focal):
    A.set_decoder(mode)
    for m in (mc, mf):
        for p in m.parameters():
            p.grad = None
    out = A.run_one_iter_of_nerf(H, W, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms=rnd)
    loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
    loss.backward()
    grads = {}
    for prefix, m in (("coarse.", mc), ("fine.", mf)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                grads[prefix + k] = p.grad.detach().clone()
    return float(loss.detach()), out, grads


def _q(t):
    """fp16 rounding with a straight-through gradient (what mixed-precision training differentiates)"""
    return t + (t.half().float() - t).detach()


def _emulated_planes_forward(model, scene_id, ro, rd, z, viewdirs, sigma_noise=None):
    """`autograd.planes_model_forward` in stock PyTorch ops with the tcgen05 path's rounding points: fp16 planes, fp16
    features, fp16 weights, fp16 hidden activations (straight-through), fp32 accumulation, fp32 biases, heads on the
    unrounded last activations, the rgb chain's view columns as an fp32 per-ray bias."""
    model.set_cur_scene_id(scene_id)
    geom = A.Geometry.of_model(model, scene_id)
    planes = [model.planes(d, super_resolve=False) for d in range(4)]
    n, S = z.shape
    feat_p, feat_m = A.TriPlaneGather.apply(_q(planes[0]), _q(planes[1]), _q(planes[2]), ro, rd, z, geom)
    vfeat = A.ViewdirGather.apply(planes[3], viewdirs, geom)
    feat_p, feat_m = _q(feat_p), _q(feat_m)
    h = feat_m
    for i, lin in enumerate(model.density_dec["0"]):
        h = torch.relu(h @ _q(lin.weight).t() + lin.bias)
        if i < 3:
            h = _q(h)
    alpha = model.fc_alpha["0"](h)
    L = list(model.rgb_dec["0"])
    C3 = feat_p.shape[1]
    rb = vfeat @ L[0].weight[:, C3:].t() + L[0].bias
    h = _q(torch.relu(feat_p @ _q(L[0].weight[:, :C3]).t() + rb[:, None, :].expand(n, S, 128).reshape(n * S, 128)))
    for i, lin in enumerate(L[1:]):
        h = torch.relu(h @ _q(lin.weight).t() + lin.bias)
        if i < 2:
            h = _q(h)
    rgb = model.fc_rgb["0"](h)
    return torch.cat([rgb, alpha], -1).reshape(n, S, 4)


def _compare(gtc, gref, what):
    worst = {}
    for k in gref:
        a, b = gtc[k].double().flatten(), gref[k].double().flatten()
        worst[k] = (float((a - b).norm() / (b.norm() + 1e-30)), float((a @ b) / (a.norm() * b.norm() + 1e-30)))
    print(f"tc vs {what} gradients (relative L2, cosine), worst 6:")
    for k, (rel, cos) in sorted(worst.items(), key=lambda kv: -kv[1][0])[:6]:
        print(f"  {k:44s} {rel:.3e} {cos:.6f}")
    return worst


@pytest.mark.parametrize("res,nc,nf", [(32, 64, 128), (19, 24, 40)])
def test_train_step_tc_vs_same_rounding_and_fp32_mode(monkeypatch, res, nc, nf):
    """One training step (1 024 rays, 64 + 128 samples, perturbation, density noise, white background) with the decoder
    forward + backward on tcgen05, fine depths teacher-forced, against
      (a) the SAME step differentiated by torch autograd through a stock-PyTorch restatement with the same 16-bit
          rounding points (`_emulated_planes_forward`): every gradient within 1 % of its norm, cosine >= 0.9999 — this
          pins the kernels' backward arithmetic end to end (the stage tests above pin each kernel);
      (b) the step in the fp32 parity mode: the mixed-precision contract.  The decoder weights (sums over every row)
          agree to a few percent.  The plane gradients carry the noise of the reference's own discontinuity: the density
          gradient of a sample is switched by relu(sigma + noise) (volume_rendering_utils.py:29-35), the 16-bit forward
          moves sigma by up to ~0.1, so ~0.5 % of the samples land on the other side and a texel sums only ~20 rows —
          measured 6-8 % relative L2 at cosine 0.997, independent of the loss scale (tests/diag_train_tc_scale.py);
          with radiance_field_noise_std > 0 the threshold is dithered by design anyway."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=DEV)
    for m in (mc, mf):
        m.train()
    Hh = Ww = res          # (19 x 19 = 361 rays with 24 + 40 samples: padding rays and padding samples in the tiles)
    pose, focal = scene.blender_camera(Hh)
    opt, scfg = scene.render_options(nc, nf, perturb=True, white_background=True, noise_std=0.2), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(Hh, Ww, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    n = batch.shape[1]
    g = torch.Generator().manual_seed(3)
    rnd = dict(t_rand=torch.rand(n, nc, generator=g), u=torch.rand(n, nf, generator=g),
               noise_c=torch.randn(n, nc, generator=g), noise_f=torch.randn(n, nc + nf, generator=g))
    target = torch.rand(n, 3, generator=g).to(DEV)
    try:
        tr = {}
        l32, o32, g32 = _step("fp32", mc, mf, sid, batch, opt, scfg, dict(rnd, trace=tr), target, Hh, Ww, focal)
        forced = dict(rnd, z_fine=tr["z_fine"])
        ltc, otc, gtc = _step("tc", mc, mf, sid, batch, opt, scfg, forced, target, Hh, Ww, focal)
        with monkeypatch.context() as mp:
            mp.setattr(A, "planes_model_forward", _emulated_planes_forward)
            lem, oem, gem = _step("fp32", mc, mf, sid, batch, opt, scfg, forced, target, Hh, Ww, focal)
    finally:
        A.set_decoder("tc")
    assert set(gtc) == set(g32) == set(gem)
    # (a) same rounding points
    assert abs(ltc - lem) <= 2e-5 * max(1.0, abs(lem)), (ltc, lem)
    assert float((otc[3] - oem[3]).abs().max()) <= 2e-3
    wa = _compare(gtc, gem, "same-rounding torch autograd")
    bad = {k: v for k, v in wa.items() if v[0] > 1e-2 or v[1] < 0.9999}
    assert not bad, bad
    # (b) fp32 parity mode
    assert abs(ltc - l32) <= 1e-3 * max(1.0, abs(l32)), (ltc, l32)
    wb = _compare(gtc, g32, "fp32-mode")
    bad = {k: v for k, v in wb.items() if (v[0] > (0.12 if "planes_" in k else 0.05)) or v[1] < (0.993 if "planes_" in k else 0.9985)}
    assert not bad, bad


def test_training_with_tc_decoder_tracks_fp32_mode():
    """Functional contract of mixed precision: fitting a student scene to a teacher's renderings with Adam (120 steps,
    1 024 rays, 64 + 64 samples, perturbation + density noise) through the tcgen05 training decoder ends at the same loss
    as the fp32 parity mode from the same start, batches and random draws (within 15 %; measured: 1.0003 after 200 steps,
    profiles/r2_train_demo.json), and both have dropped."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("train_demo", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                             "scripts", "train_demo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.main(["--steps", "120", "--rays", "1024"])
    print({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "curve"}) for k, v in r.items()})
    assert r["tc"]["last"] < 0.8 * r["tc"]["first"] and r["fp32"]["last"] < 0.8 * r["fp32"]["first"]
    assert 0.85 <= r["last_loss_ratio_tc_over_fp32"] <= 1.15, r["last_loss_ratio_tc_over_fp32"]


def test_frozen_decoder_step_gives_the_same_plane_gradients():
    """The phase in which only planes / the SR model train (decoder frozen, train_nerf.py:560): PlanesRadianceTC skips the
    weight-gradient kernels; the plane gradients are the ones of the unfrozen step, bit for bit."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=48, view_res=12, seed=5, device=DEV)
    for m in (mc, mf):
        m.train()
    Hh = Ww = 24
    pose, focal = scene.blender_camera(Hh)
    opt, scfg = scene.render_options(32, 32, perturb=True), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(Hh, Ww, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    n = batch.shape[1]
    g = torch.Generator().manual_seed(1)
    rnd = dict(t_rand=torch.rand(n, 32, generator=g), u=torch.rand(n, 32, generator=g))
    target = torch.rand(n, 3, generator=g).to(DEV)
    _, _, g_all = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, Hh, Ww, focal)
    for m in (mc, mf):
        for k, p in m.named_parameters():
            if "planes_" not in k:
                p.requires_grad_(False)
    ops.LAUNCHES.clear()
    _, _, g_frozen = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, Hh, Ww, focal)
    assert ops.LAUNCHES.get("nvsr_mlp_wgrad", 0) == 0 and ops.LAUNCHES.get("nvsr_mlp_dgrad", 0) == 4
    assert set(g_frozen) == {k for k in g_all if "planes_" in k}
    for k in g_frozen:
        assert torch.equal(g_frozen[k], g_all[k]) or float((g_frozen[k] - g_all[k]).abs().max()) <= 1e-6 * float(g_all[k].abs().max()), k


# ---- row-list ("sparse") backward: csrc/train_tc.cu nonzero_rows / compact_rows / *_rows ----------------------------
@pytest.mark.parametrize("n_rays,S,frac", [(64, 40, 0.2), (19, 24, 0.9), (40, 64, 0.0)])
def test_row_list_stages_vs_torch(n_rays, S, frac):
    """nonzero_rows lists exactly the rows with a non-zero raw gradient; compact_rows moves those rows of tile images and
    of d_raw to list order (tail of the last tile zero); ray_sum_rows adds the list rows to their rays."""
    g = torch.Generator().manual_seed(n_rays + S)
    rows = ops.rows_padded(n_rays, S, ops.ROWS_BLOCKED)
    d_rf = torch.randn(n_rays, S, 4, generator=g) * (torch.rand(n_rays, S, 1, generator=g) < frac)
    d_rf[0, 0, 2] = float("nan") if frac > 0 else 0.0          # NaN rows are listed
    d_raw = ops.nsc_to_planar_blocked(d_rf.to(DEV), n_rays, S)
    ids, count = ops.nonzero_rows(d_raw)
    k = int(count)
    want = torch.nonzero(((d_raw != 0) | d_raw.isnan()).any(0)).flatten()
    assert k == want.numel()
    got = ids[:k].long().sort().values
    assert torch.equal(got, want)
    imgs = [torch.randn(rows, c, generator=g).half().to(DEV) for c in (32, 128, 96)]
    timgs = [rows_to_img(x.float()) for x in imgs]
    cimgs, c_raw = ops.compact_rows(timgs, d_raw, ids, count)
    kp = -(-k // 128) * 128
    sel = ids[:k].long()
    for x, ci in zip(imgs, cimgs):
        r = img_to_rows(ci)
        assert torch.equal(r[:k], x[sel].float())
        assert not r[k:kp].any()
    assert torch.equal(c_raw[:, :k].nan_to_num(7.0), d_raw[:, sel].nan_to_num(7.0)) and not c_raw[:, k:kp].any()
    # ray sums of the 128-channel image in list order
    out = ops.ray_sum_rows(cimgs[1], ids, count, n_rays, S, 0.5)
    tile, r = sel // 128, sel % 128
    ray = (tile // (-(-S // ops.BLK_SAMPLES))) * ops.BLK_RAYS + (r % ops.BLK_RAYS)
    ref = torch.zeros(n_rays, 128, dtype=torch.float64, device=DEV).index_add_(0, ray, imgs[1][sel].double() * 0.5)
    assert float((out.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("res,nc,nf,white,noise", [(32, 64, 128, True, 0.2), (19, 24, 40, False, 0.0)])
def test_row_list_backward_equals_dense_backward(res, nc, nf, white, noise):
    """The default backward of the 'tc' decoder visits only the samples with a non-zero raw gradient.  Same step, same
    draws, with the row list on and off: every gradient agrees to the order of the fp32 sums (the per-row arithmetic
    is the same kernels on the same operands), and the list is a strict subset of the samples."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=DEV)
    for m in (mc, mf):
        m.train()
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True, white_background=white, noise_std=noise), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    n = batch.shape[1]
    g = torch.Generator().manual_seed(11)
    rnd = dict(t_rand=torch.rand(n, nc, generator=g), u=torch.rand(n, nf, generator=g))
    if noise:
        rnd.update(noise_c=torch.randn(n, nc, generator=g), noise_f=torch.randn(n, nc + nf, generator=g))
    target = torch.rand(n, 3, generator=g).to(DEV)
    try:
        A.set_sparse_forward(False)      # (the sparse forward brings its own list: tested below)
        A.set_sparse_backward(False)
        ops.LAUNCHES.clear()
        l_d, _, g_dense = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        assert ops.LAUNCHES.get("nvsr_nonzero_rows", 0) == 0
        A.set_sparse_backward(True)
        ops.LAUNCHES.clear()
        l_s, _, g_rows = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        assert ops.LAUNCHES.get("nvsr_nonzero_rows", 0) == 2
    finally:
        A.set_sparse_backward(True)
        A.set_sparse_forward(True)
    assert l_d == l_s and set(g_dense) == set(g_rows)
    for k in g_dense:
        a, b = g_rows[k].double(), g_dense[k].double()
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-12, (k, float((a - b).abs().max()), float(b.abs().max()))


def test_row_list_backward_with_no_gradient_rows():
    """an all-zero upstream gradient: the list is empty, no tile is visited, every gradient is exactly zero"""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=32, view_res=8, seed=2, device=DEV)
    pose, focal = scene.blender_camera(16)
    opt, scfg = scene.render_options(24, 24, perturb=False), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(16, 16, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    for m in (mc, mf):
        m.train()
        for p in m.parameters():
            p.grad = None
    out = A.run_one_iter_of_nerf(16, 16, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms={})
    ((out[0] * 0.0).sum() + (out[3] * 0.0).sum()).backward()
    gs = [p.grad for m in (mc, mf) for p in m.parameters() if p.grad is not None]
    assert gs and all(not bool(x.any()) for x in gs)


def test_graphed_train_step_matches_eager_step():
    """autograd.GraphedStep: the whole step (forward, sparse backward, SGD update) captured into a CUDA graph; replays
    with new ray batches give the same losses and the same parameters as the eager loop from the same start (plain SGD:
    an Adam update divides by |gradient| and turns the fp32 summation-order noise of near-zero entries into +-lr)."""
    import copy
    mc0, mf0, sid = scene.make_synthetic_scene(plane_res=48, view_res=12, seed=4, device=DEV)
    res, nc, nf, n = 32, 32, 32, 512
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    rays = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    g = torch.Generator().manual_seed(2)
    steps = 6
    picks = [torch.randperm(res * res, generator=g)[:n].to(DEV) for _ in range(steps)]
    targets = [torch.rand(n, 3, generator=g).to(DEV) for _ in range(steps)]
    draws = [dict(t_rand=torch.rand(n, nc, generator=g).to(DEV), u=torch.rand(n, nf, generator=g).to(DEV)) for _ in range(steps)]

    def run(graphed):
        mc, mf = copy.deepcopy(mc0), copy.deepcopy(mf0)
        for m in (mc, mf):
            m.train()
        params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
        optim = torch.optim.SGD(params, lr=0.2)
        batch, target = rays[:, picks[0]].clone(), targets[0].clone()
        rnd = {k: v.clone() for k, v in draws[0].items()}
        snapshot = [p.detach().clone() for p in params]

        def step():
            optim.zero_grad(set_to_none=True)
            out = A.run_one_iter_of_nerf(res, res, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms=rnd)
            loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
            loss.backward()
            optim.step()
            return loss
        fn = step
        if graphed:
            fn = A.GraphedStep(step, warmup=2)
            # the warm-up and the capture ran real updates: restart from the snapshot
            with torch.no_grad():
                for p, s0 in zip(params, snapshot):
                    p.copy_(s0)
        losses = []
        for i in range(steps):
            batch.copy_(rays[:, picks[i]]), target.copy_(targets[i])
            for k in rnd:
                rnd[k].copy_(draws[i][k])
            losses.append(float(fn().detach()))
        return losses, [p.detach().clone() for p in params], snapshot

    le, pe, p0 = run(False)
    lg, pg, _ = run(True)
    assert all(abs(a - b) <= 1e-4 * abs(a) for a, b in zip(le, lg)), (le, lg)
    moved = max(float((a - s0).abs().max()) for a, s0 in zip(pe, p0))
    assert moved > 1e-4          # the six updates did something
    for a, b, s0 in zip(pe, pg, p0):
        assert float((a - b).abs().max()) <= 1e-3 * float((a - s0).abs().max()) + 1e-7


def test_pack_weights16_one_launch_and_deferred_range_check():
    """nvsr_pack_weights16: every image equals nvsr_pack_weight16's, one launch for the lot, and the maximum |w| lands in
    the device scalar of the deferred fp16 range check — which raises one step late (or at flush) without any host read
    inside the step."""
    g = torch.Generator().manual_seed(0)
    big = torch.randn(128, 160, generator=g).to(DEV)
    ws = [torch.randn(128, 48, generator=g).to(DEV), big[:, :144], torch.randn(128, 128, generator=g).to(DEV) * 3.0,
          torch.randn(3, 128, generator=g).to(DEV)]
    rc = ops.DeferredRangeCheck()
    ops.LAUNCHES.clear()
    imgs = ops.pack_weights16(ws, NVSR_F16, range_check=rc)
    assert ops.LAUNCHES.get("nvsr_pack_weights16") == 1
    for w, im in zip(ws, imgs):
        assert torch.equal(im, ops.pack_weight16(w.contiguous(), dtype=NVSR_F16))
    want = max(float(w.abs().max()) for w in ws)
    assert float(rc.dev[torch.device(DEV)]) == want
    rc.commit()
    rc.flush()                                   # in range: silent
    assert float(rc.dev[torch.device(DEV)]) == 0.0
    ws[2][5, 7] = 1.0e5
    ops.pack_weights16(ws, NVSR_F16, range_check=rc)
    rc.commit()
    with pytest.raises(nvsr_b200.NvsrError, match="fp16 range"):
        rc.flush()
    ws[2][5, 7] = float("nan")
    ops.pack_weights16(ws, NVSR_F16, range_check=rc)
    rc.commit()
    with pytest.raises(nvsr_b200.NvsrError, match="fp16 range"):
        rc.flush()


@pytest.mark.parametrize("res,nc,nf,white,noise", [(32, 64, 128, True, 0.2), (19, 24, 40, False, 0.0)])
def test_sparse_training_forward_equals_dense_forward(res, nc, nf, white, noise):
    """`set_sparse_forward` (default on): the rgb chain of the training forward runs over the samples with
    relu(sigma + noise) > 0 only.  Same step with it off: bit-identical maps and loss (a dropped sample has weight exactly
    0), gradients equal to the order of the fp32 sums; the rgb chain sees a strict subset of the rows."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=DEV)
    for m in (mc, mf):
        m.train()
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True, white_background=white, noise_std=noise), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    n = batch.shape[1]
    g = torch.Generator().manual_seed(12)
    rnd = dict(t_rand=torch.rand(n, nc, generator=g), u=torch.rand(n, nf, generator=g))
    if noise:
        rnd.update(noise_c=torch.randn(n, nc, generator=g), noise_f=torch.randn(n, nc + nf, generator=g))
    target = torch.rand(n, 3, generator=g).to(DEV)
    try:
        A.set_sparse_forward(False)
        ops.LAUNCHES.clear()
        l_d, o_d, g_d = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        assert ops.LAUNCHES.get("nvsr_keep_rows", 0) == 0 and ops.LAUNCHES.get("nvsr_nonzero_rows", 0) == 2
        A.set_sparse_forward(True)
        ops.LAUNCHES.clear()
        l_s, o_s, g_s = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        assert ops.LAUNCHES.get("nvsr_keep_rows", 0) == 2 and ops.LAUNCHES.get("nvsr_nonzero_rows", 0) == 0
        assert ops.LAUNCHES.get("nvsr_sample_gather_rows", 0) == 2
    finally:
        A.set_sparse_forward(True)
    assert l_d == l_s
    for k in (0, 2, 3, 5):
        assert torch.equal(o_d[k], o_s[k]), k
    assert set(g_d) == set(g_s)
    for k in g_d:
        a, b = g_s[k].double(), g_d[k].double()
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-12, (k, float((a - b).abs().max()), float(b.abs().max()))


def test_device_rng_step_needs_no_randoms_and_is_capturable():
    """`set_device_rng(True)`: the random draws come from torch's CUDA generator, so a step without caller-provided
    `randoms` has no host upload, captures into a CUDA graph, and every replay draws new numbers."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=48, view_res=12, seed=6, device=DEV)
    for m in (mc, mf):
        m.train()
    res, nc, nf = 24, 32, 32
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True, noise_std=0.1), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    target = torch.rand(batch.shape[1], 3, generator=torch.Generator().manual_seed(0)).to(DEV)
    params = [p for m in (mc, mf) for p in m.parameters()]

    def step():
        for p in params:
            p.grad = None
        out = A.run_one_iter_of_nerf(res, res, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg)
        loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
        loss.backward()
        return loss
    A.set_device_rng(True)
    try:
        torch.manual_seed(1)
        a = float(step().detach())
        torch.manual_seed(1)
        b = float(step().detach())
        assert a == b                                    # torch's CUDA generator: seedable like the CPU one
        graphed = A.GraphedStep(step, warmup=2)
        losses = [float(graphed().detach()) for _ in range(4)]
        assert len(set(losses)) == 4, losses             # new draws in every replay
        assert all(abs(x - a) < 0.05 * abs(a) + 0.05 for x in losses)
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in params if p.requires_grad)
    finally:
        A.set_device_rng(False)


def test_render_after_graphed_training_sees_the_updated_parameters():
    """An optimizer step inside a replayed CUDA graph changes planes and weights without moving any `_version` counter —
    the signature the packed-plane / packed-weight caches key on.  GraphedStep bumps the caches' generation after every
    replay: a render after graphed training equals the render after `clear_caches()` bit for bit (and is not the
    pre-training image the caches held)."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=48, view_res=12, seed=8, device=DEV)
    res, nc, nf = 32, 32, 32
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True), scene.scene_cfg()
    vopt = scene.render_options(nc, nf)
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0).contiguous()
    n = batch.shape[1]
    g = torch.Generator().manual_seed(4)
    rnd = dict(t_rand=torch.rand(n, nc, generator=g).to(DEV), u=torch.rand(n, nf, generator=g).to(DEV))
    target = torch.rand(n, 3, generator=g).to(DEV)

    def render():
        with torch.no_grad():
            out = nvsr_b200.run_one_iter_of_nerf(res, res, focal, mc, mf, batch, vopt, sid, "validation", scene_config=scfg)
        return out[3].clone()
    before = render()
    for m in (mc, mf):
        m.train()
    params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
    optim = torch.optim.SGD(params, lr=0.5)

    def step():
        optim.zero_grad(set_to_none=True)
        out = A.run_one_iter_of_nerf(res, res, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms=rnd)
        loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
        loss.backward()
        optim.step()
        return loss
    graphed = A.GraphedStep(step, warmup=1)
    for _ in range(2):
        graphed()
    mid = render()                                     # caches now hold the scene as of replay 2 (versions as captured)
    for _ in range(4):
        graphed()                                      # ... and these move the parameters with no version tick
    after = render()
    scene.clear_caches()
    fresh = render()
    assert torch.equal(after, fresh)
    assert float((after - mid).abs().max()) > 1e-4 and float((mid - before).abs().max()) > 1e-4


def test_precise_density_forward_closes_most_of_the_gap_to_the_fp32_mode():
    """`set_precise_density(True)`: sigma of the training forward from the split-operand chain (fp32 grade).  Free-running
    step with density noise against the fp32 parity mode: the rendered maps come within 1e-3 (the fp16 forward: several
    1e-3) and the worst plane-gradient deviation drops well below the plain 'tc' step's, whose excess is the
    relu(sigma + noise) flips of the fp16 sigma."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=DEV)
    for m in (mc, mf):
        m.train()
    res, nc, nf = 32, 64, 128
    pose, focal = scene.blender_camera(res)
    opt, scfg = scene.render_options(nc, nf, perturb=True, white_background=True, noise_std=0.2), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(DEV))
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    n = batch.shape[1]
    g = torch.Generator().manual_seed(3)
    rnd = dict(t_rand=torch.rand(n, nc, generator=g), u=torch.rand(n, nf, generator=g),
               noise_c=torch.randn(n, nc, generator=g), noise_f=torch.randn(n, nc + nf, generator=g))
    target = torch.rand(n, 3, generator=g).to(DEV)
    try:
        _, o32, g32 = _step("fp32", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        _, otc, gtc = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        A.set_precise_density(True)
        ops.LAUNCHES.clear()
        _, opr, gpr = _step("tc", mc, mf, sid, batch, opt, scfg, rnd, target, res, res, focal)
        assert ops.LAUNCHES.get("nvsr_mlp_chain_split", 0) == 2
    finally:
        A.set_precise_density(False)
        A.set_decoder("tc")

    def worst(ga):
        return max(float((ga[k].double() - g32[k].double()).norm() / (g32[k].double().norm() + 1e-30)) for k in g32 if "planes_" in k)
    e_tc, e_pr = float((otc[0] - o32[0]).abs().max()), float((opr[0] - o32[0]).abs().max())
    w_tc, w_pr = worst(gtc), worst(gpr)
    print("coarse map error tc %.2e precise %.2e | worst plane-gradient rel L2 tc %.3f precise %.3f" % (e_tc, e_pr, w_tc, w_pr))
    assert e_pr <= 1e-3 and e_pr < 0.5 * e_tc
    assert w_pr < 0.7 * w_tc
