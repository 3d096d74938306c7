"""GPU parity, stage by stage: the CUDA path (through the C-ABI) against golden vectors produced by
the reference and against the CPU oracle on seeded inputs.

Tolerances (stated, per BASELINE.json north_star):
  fp32 mode  : rgb/depth/acc/weights within 1e-3 abs of the reference (observed ~1e-6);
  bf16 mode  : decoder outputs carry bf16 rounding of features/activations: stated in BF16_*.
  integer/index work (searchsorted indices, ray order): bit-exact.
"""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import T, golden
from oracle import nvsr_oracle as O

import nvsr_b200
from nvsr_b200 import NVSR_BF16, NVSR_F16, NVSR_F32, ops, scene
from nvsr_b200._lib import FEAT_ROWMAJOR_F32, FEAT_TILE_BF16, FEAT_TILE_F16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_get_ray_bundle_golden(tag):
    g = golden(f"stage_raybundle_{tag}.npz")
    ro, rd = nvsr_b200.get_ray_bundle(int(g["H"]), int(g["W"]), g["focal"].tolist(), T(g["pose"], DEV), int(g["pad"]),
                                      float(g["offset"]))
    H.assert_close(ro, g["ro"], 0, what="ro")
    H.assert_close(rd, g["rd"], 1e-6, what="rd")


def test_get_ray_bundle_800_bands_and_order():
    pose, focal = scene.blender_camera(800)
    ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(DEV))
    ro_o, rd_o = O.get_ray_bundle(800, 800, focal, pose)
    H.assert_close(rd, rd_o, 1e-6, what="rd 800")
    frac_exact = float((rd.cpu() == rd_o).float().mean())
    assert frac_exact > 0.99, frac_exact
    bands = [nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(DEV), row_range=(r, r + 100))[1] for r in range(0, 800, 100)]
    assert torch.equal(torch.cat(bands, 0), rd)  # row-band sharding preserves ray order bit-exactly
    assert ro.shape == (800, 800, 3) and torch.equal(ro[5, 7].cpu(), pose[:3, 3])


def test_prepare_rays_ndc_golden():
    g = golden("stage_ndc.npz")
    ro, rd, vd = ops.prepare_rays(T(g["ro"], DEV), T(g["rd"], DEV), True, int(g["H"]), int(g["W"]), float(g["focal"]), 1.0)
    H.assert_close(ro, g["ro_ndc"], 2e-6, 2e-6, what="ro_ndc")
    H.assert_close(rd, g["rd_ndc"], 2e-6, 2e-6, what="rd_ndc")
    ref = T(g["rd"]) / T(g["rd"]).norm(p=2, dim=-1).unsqueeze(-1)
    H.assert_close(vd, ref, 2e-7, what="viewdirs")


@pytest.mark.parametrize("dtype,tdt", [(NVSR_F16, torch.float16), (NVSR_BF16, torch.bfloat16), (NVSR_F32, torch.float32)])
@pytest.mark.parametrize("shape", [(48, 200, 200), (16, 5, 9), (8, 3, 1)])
def test_pack_plane_layouts(dtype, tdt, shape):
    """nvsr_pack_plane: fp32 -> channels-last [Rh,Rw,C]; 16-bit -> x-pair records [Rh,C/8,Rw,2,8] (texel chunk +
    its right neighbour's, the last column paired with itself), values rounded to nearest and saturated."""
    torch.manual_seed(11)
    c, rh, rw = shape
    plane = torch.randn(1, c, rh, rw) * 3
    plane[0, 0, 0, 0] = 1e6     # beyond fp16
    plane[0, 1, 0, 0] = -1e6
    if dtype == NVSR_F16:
        # the host wrapper refuses a plane that would saturate (silently wrong values otherwise) ...
        with pytest.raises(nvsr_b200.NvsrError, match="fp16 range"):
            ops.pack_plane(plane.to(DEV), dtype)
        # ... and the kernel itself saturates to the largest finite value instead of producing inf
        import ctypes as C
        from nvsr_b200 import _lib
        p_dev = plane.to(DEV).contiguous()
        raw_out = torch.empty((rh, c // 8, rw, 2, 8), dtype=tdt, device=DEV)
        st = _lib.load().nvsr_pack_plane(C.c_void_p(p_dev.data_ptr()), c, rh, rw, C.c_void_p(raw_out.data_ptr()), dtype,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert st == 0
        out = raw_out.cpu()
    else:
        out = ops.pack_plane(plane.to(DEV), dtype).cpu()
    if dtype == NVSR_F32:
        assert out.shape == (rh, rw, c)
        assert torch.equal(out, plane[0].permute(1, 2, 0).contiguous())
        return
    assert out.shape == (rh, c // 8, rw, 2, 8) and out.dtype == tdt
    fmax = torch.finfo(tdt).max
    ref = plane[0].clamp(-fmax, fmax).to(tdt)                      # [C,Rh,Rw], round to nearest even
    left = ref.reshape(c // 8, 8, rh, rw).permute(2, 0, 3, 1)      # [Rh,C/8,Rw,8]
    right = left[:, :, torch.clamp(torch.arange(rw) + 1, max=rw - 1)]
    assert torch.equal(out[..., 0, :], left)
    assert torch.equal(out[..., 1, :], right)
    assert torch.isfinite(out.float()).all()


def _forward_points(model, sid, x6, precision):
    """TwoDimPlanesModel.forward(x[n,6]) through the kernels: points as 1-sample rays (rd = 0)."""
    model.set_cur_scene_id(sid)
    packed = scene.pack_scene_planes(model, sid, precision)
    dec = scene.pack_planes_decoder(model, precision)
    n = x6.shape[0]
    ro = x6[:, :3].contiguous()
    rd = torch.zeros_like(ro)
    layout = ops.FEAT_LAYOUT[precision]
    fp, fm, _ = ops.sample_gather(ro, rd, 0.0, 1.0, packed, layout, z_in=torch.zeros(n, 1, device=DEV))
    vfeat = ops.viewdir_gather(x6[:, 3:].contiguous(), packed)
    rb = ops.row_bias(vfeat, dec.view_w, dec.view_b)
    order = ops.LAYOUT_ROWS[layout]
    rows = ops.rows_padded(n, 1, order)
    raw = ops.raw_buffer(n, 1, order, DEV).zero_()
    ops.mlp_chain(fm, dec.density, rows, raw, precision, 1, n, order)
    ops.mlp_chain(fp, dec.rgb_chain(rb), rows, raw, precision, 1, n, order)
    return fp, fm, vfeat, ops.raw_to_nsc(raw, n, 1, order)[:, 0, :]


def _untile(img, n_rays, n_samples=1):
    """BLOCKED [tiles, K/8, 128, 8] 16-bit tile image -> [n_rays*n_samples, K] fp32 in ray-major order"""
    t, kc, r, e = img.shape
    flat = img.permute(0, 2, 1, 3).reshape(t * r, kc * e).float()          # [padded rows, K], BLOCKED order
    nb, ts = -(-n_rays // 8), -(-n_samples // 16)
    x = flat.reshape(nb, ts, 16, 8, kc * e).permute(0, 3, 1, 2, 4).reshape(nb * 8, ts * 16, kc * e)
    return x[:n_rays, :n_samples].reshape(n_rays * n_samples, kc * e)


def test_planes_forward_fp32_golden():
    g = golden("stage_planes_forward.npz")
    sid = str(g["scene_id"])
    mc, _ = H.load_planes_scene("scene_planes_small.npz", sid, DEV)
    fp, fm, vfeat, out = _forward_points(mc, sid, T(g["x6"], DEV), NVSR_F32)
    for d in range(3):
        H.assert_close(fp[:, d * 48:(d + 1) * 48], g[f"pos{d}"], 2e-6, what=f"pos{d}")
    mean = (T(g["pos0"]) + T(g["pos1"]) + T(g["pos2"])) / 3
    H.assert_close(fm, mean, 2e-6, what="mean")
    H.assert_close(vfeat, g["view"], 5e-6, what="view")
    H.assert_close(out, g["out"], 1e-3, 1e-5, what="forward out")


# raw decoder outputs (pre-sigmoid; |sigma| up to ~50 on these scenes), 16-bit operand modes
RAW_TOL = {NVSR_BF16: (0.40, 0.02), NVSR_F16: (0.05, 0.003)}
FEAT_TOL = {NVSR_BF16: (2e-2, 1e-2), NVSR_F16: (2e-3, 1.5e-3)}


@pytest.mark.parametrize("precision", [NVSR_BF16, NVSR_F16], ids=["bf16", "fp16"])
def test_planes_forward_16bit_golden(precision):
    g = golden("stage_planes_forward.npz")
    sid = str(g["scene_id"])
    mc, _ = H.load_planes_scene("scene_planes_small.npz", sid, DEV)
    n = g["x6"].shape[0]
    fp, fm, vfeat, out = _forward_points(mc, sid, T(g["x6"], DEV), precision)
    feats = _untile(fp, n)
    for d in range(3):
        H.assert_close(feats[:, d * 48:(d + 1) * 48], g[f"pos{d}"], *FEAT_TOL[precision], what=f"16-bit pos{d}")
    err = (out.cpu() - T(g["out"])).abs()
    print("16-bit forward (%s): max abs err rgb %.4f sigma %.4f (|sigma| max %.1f)" % (
        "bf16" if precision == NVSR_BF16 else "fp16", float(err[:, :3].max()), float(err[:, 3].max()),
        float(np.abs(g["out"][:, 3]).max())))
    H.assert_close(out, g["out"], *RAW_TOL[precision], what="16-bit forward out")


@pytest.mark.parametrize("precision", [NVSR_F32, NVSR_BF16, NVSR_F16], ids=["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("rows", [1, 127, 128, 1000, 40000])
def test_mlp_chain_vs_torch(precision, rows):
    """decoder chain kernels vs a plain fp32 torch evaluation of the same layers (ragged tile counts)"""
    torch.manual_seed(rows)
    k0, S = 144, 8
    n_rays = (rows + S - 1) // S
    lins = [torch.nn.Linear(k0, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 128)]
    head = torch.nn.Linear(128, 3)
    x = torch.randn(rows, k0)
    rbias = torch.randn(n_rays, 128)
    with torch.no_grad():
        if precision != NVSR_F32:
            q = (lambda t: t.bfloat16().float()) if precision == NVSR_BF16 else (lambda t: t.half().float())
            h = q(x)
            for i, l in enumerate(lins):
                w = q(l.weight)
                b = rbias[torch.arange(rows) // S] if i == 0 else l.bias
                h = torch.relu(h @ w.t() + b)
                if i < 3:
                    h = q(h)
            ref = h @ head.weight.t() + head.bias
        else:
            h = x
            for i, l in enumerate(lins):
                b = rbias[torch.arange(rows) // S] if i == 0 else l.bias
                h = torch.relu(h @ l.weight.t() + b)
            ref = h @ head.weight.t() + head.bias
    layers = []
    for i, l in enumerate(lins):
        w = l.weight.detach().to(DEV)
        wp = ops.pack_weight16(w, dtype=precision) if precision != NVSR_F32 else w.contiguous()
        layers.append(ops.ChainLayer(wp, None if i == 0 else l.bias.detach().to(DEV), l.in_features, 128, True,
                                     row_bias=rbias.to(DEV) if i == 0 else None,
                                     head_w=head.weight.detach().to(DEV) if i == 3 else None,
                                     head_b=head.bias.detach().to(DEV) if i == 3 else None, head_ch=0))
    if precision != NVSR_F32:
        tiles = (rows + 127) // 128
        xin = torch.zeros(tiles * 128, k0)
        xin[:rows] = x
        inp = xin.reshape(tiles, 128, k0 // 8, 8).permute(0, 2, 1, 3).contiguous().to(ops.TORCH_DTYPE[precision]).to(DEV)
    else:
        inp = x.to(DEV)
    raw = torch.full((4, (rows + 127) // 128 * 128), float("nan"), device=DEV)
    ops.mlp_chain(inp, layers, rows, raw, precision, S, n_rays)
    torch.cuda.synchronize()
    out = raw[:3, :rows].t()
    # the torch evaluation rounds operands exactly like the kernel: only accumulation order (and rare
    # one-ulp re-rounding of a hidden activation) differs
    tol = {NVSR_F32: 2e-4, NVSR_BF16: 4e-3, NVSR_F16: 5e-4}[precision]
    H.assert_close(out, ref, tol, tol, what=f"mlp rows={rows}")
    assert torch.isnan(raw[3]).all()  # untouched channel stays untouched


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("white", dict(white_background=True)), ("mip", dict(mip_nerf=True))])
def test_volume_render_golden(tag, kw):
    g = golden(f"stage_composite_{tag}.npz")
    o = nvsr_b200.volume_render_radiance_field(T(g["raw"], DEV), T(g["z"], DEV), T(g["rd"], DEV), **kw)
    for name, v in zip(("rgb", "disp", "acc", "weights", "depth"), o):
        H.assert_close(v, g[name], 2e-6, 1e-5, what=name)   # NaN pattern of disp is checked inside


@pytest.mark.parametrize("tag", ["rand", "dyadic"])
def test_sample_pdf_golden(tag):
    g = golden(f"stage_samplepdf_{tag}.npz")
    bins, w = T(g["bins"], DEV), T(g["weights"], DEV)
    # (i) stage test: same cdf, same u  =>  indices bit-exact, always
    smp, inds, _ = nvsr_b200.sample_pdf(bins, None, 48, det=True, cdf=T(g["cdf"], DEV), return_all=True)
    assert torch.equal(inds.cpu(), T(g["inds"]))
    H.assert_close(smp, g["samples"], 1e-6, what="samples from golden cdf")
    # (ii)/(iii) full path: `total` is summed in ATen's own CPU order, the cumsum in fp64 like ATen's,
    # so for IDENTICAL weights the cdf and the indices are bit-exact for dyadic AND random weights
    smp, inds, cdf = nvsr_b200.sample_pdf(bins, w, 48, det=True, return_all=True)
    H.assert_close(cdf, g["cdf"], 1.2e-7, what="cdf")
    exact = float((cdf.cpu() == T(g["cdf"])).float().mean())
    print(f"sample_pdf[{tag}]: cdf bit-exact fraction {exact:.4f}, index mismatches {int((inds.cpu() != T(g['inds'])).sum())}")
    mism = H.check_resampling(inds, smp, T(g["inds"]), T(g["samples"]), T(g["cdf"]), T(g["bins"]), T(g["u"]), tag,
                              max_flip_frac=0.0 if tag == "dyadic" else 0.005)
    if tag == "dyadic":
        assert not mism.any()


def test_sample_pdf_random_u_golden():
    g = golden("stage_samplepdf_random_u.npz")
    bins, w, u = T(g["bins"]), T(g["weights"]), T(g["u"])
    smp_ref, inds_ref, cdf_ref = O.sample_pdf(bins, w, 31, det=False, u=u, return_all=True)
    H.assert_close(smp_ref, g["samples"], 2e-6, what="oracle vs golden")
    smp, inds, _ = nvsr_b200.sample_pdf(bins.to(DEV), w.to(DEV), 31, det=False, u=u.to(DEV), return_all=True)
    H.check_resampling(inds, smp, inds_ref, smp_ref, cdf_ref, bins, u, "random u")


def test_sample_pdf_large_bins_cascade():
    """n = 1022 weights exercises the 4-level cascade of the emulated ATen summation order"""
    torch.manual_seed(9)
    n, B = 64, 1023
    bins = torch.sort(torch.rand(n, B) * 4 + 2, -1)[0]
    w = torch.rand(n, B - 1)
    smp_ref, inds_ref, cdf_ref = O.sample_pdf(bins, w, 200, det=True, return_all=True)
    smp, inds, cdf = nvsr_b200.sample_pdf(bins.to(DEV), w.to(DEV), 200, det=True, return_all=True)
    print("cascade: cdf bit-exact fraction", float((cdf.cpu() == cdf_ref).float().mean()))
    H.assert_close(cdf, cdf_ref, 1.2e-7, what="cdf 1023")
    H.check_resampling(inds, smp, inds_ref, smp_ref, cdf_ref, bins, torch.linspace(0, 1, 200), "cascade")


def test_composite_resample_merge_properties():
    """sort(cat(z_vals, z_samples)) by rank-merge: output sorted and a permutation of the inputs,
    for deterministic u (fast path) and random u (all-pairs path)."""
    torch.manual_seed(3)
    n, S, nf = 513, 64, 128
    z = torch.sort(torch.rand(n, S) * 4 + 2, -1)[0].to(DEV)
    raw = (torch.randn(4, n * S) * 3).to(DEV)
    rd = torch.randn(n, 3).to(DEV)
    for u in (torch.linspace(0, 1, nf).to(DEV), torch.rand(n, nf).to(DEV)):
        o = ops.composite(raw, z, rd, S, n_fine=nf, u=u, want_samples=True, want_inds=True, want_weights=True)
        zm = o["z_merged"]
        assert bool((zm[:, 1:] >= zm[:, :-1]).all())
        ref = torch.sort(torch.cat([z, o["z_samples"]], -1), -1)[0]
        assert torch.equal(zm, ref)
        # oracle on the same weights
        w = o["weights"].cpu()
        mid = 0.5 * (z[:, 1:] + z[:, :-1]).cpu()
        smp, inds, cdf = O.sample_pdf(mid, w[:, 1:-1], nf, det=True, u=u.cpu(), return_all=True)
        H.check_resampling(o["inds"], o["z_samples"], inds, smp, cdf, mid, u.cpu(), "composite resample")


@pytest.mark.parametrize("n,S,nf", [(513, 64, 128), (37, 193, 0), (8, 16, 5), (1, 70, 0)])
def test_composite_blocked_equals_ray_major(n, S, nf):
    """The BLOCKED raw order (8 rays x 16 samples per tile, what the tcgen05 decoder writes) must give
    bit-identical maps, indices and merged depths to the ray-major order, ragged sizes included."""
    torch.manual_seed(5)
    z = torch.sort(torch.rand(n, S) * 4 + 2, -1)[0].to(DEV)
    rf = (torch.randn(n, S, 4) * 3).to(DEV)
    rd = torch.randn(n, 3).to(DEV)
    u = torch.linspace(0, 1, nf).to(DEV) if nf else None
    raw_rm = ops.raw_to_planar(rf)
    nb, ts = -(-n // 8), -(-S // 16)
    pad = torch.full((nb * 8, ts * 16, 4), float("nan"), device=DEV)       # padding rows must never be read
    pad[:n, :S] = rf
    raw_bl = pad.reshape(nb, 8, ts, 16, 4).permute(4, 0, 2, 3, 1).reshape(4, -1).contiguous()
    assert torch.equal(ops.raw_to_nsc(raw_bl, n, S, ops.ROWS_BLOCKED), rf)
    kw = dict(n_fine=nf, u=u, want_samples=nf > 0, want_inds=nf > 0, want_weights=True)
    a = ops.composite(raw_rm, z, rd, S, row_order=ops.ROWS_RAY_MAJOR, **kw)
    b = ops.composite(raw_bl, z, rd, S, row_order=ops.ROWS_BLOCKED, **kw)
    for k in a:
        assert torch.equal(a[k], b[k]) or (k == "disp" and torch.equal(torch.nan_to_num(a[k], 7.0), torch.nan_to_num(b[k], 7.0))), k


def test_composite_full_size_properties():
    """BASELINE config-2 chunk size (327 680 rays, 64 coarse -> 128 fine) through size-independent properties:
    merged depths sorted and a permutation of cat(z, z_samples); inds within range and consistent with the
    samples' bins; acc = sum(weights); the white-background identity; blocked == ray-major at full size."""
    torch.manual_seed(21)
    n, S, nf = 327680, 64, 128
    t = torch.linspace(0, 1, S, device=DEV)
    z = (2.0 * (1 - t) + 6.0 * t).expand(n, S).contiguous()
    rf = torch.randn(n, S, 4, device=DEV) * 2
    rf[..., 3] = rf[..., 3] * 3 - 4                       # sparse density: weights concentrate on a few bins
    rd = torch.randn(n, 3, device=DEV)
    u = torch.linspace(0, 1, nf, device=DEV)
    raw = ops.raw_to_planar(rf)
    o = ops.composite(raw, z, rd, S, n_fine=nf, u=u, want_samples=True, want_inds=True, want_weights=True)
    zm, zs, inds, w = o["z_merged"], o["z_samples"], o["inds"], o["weights"]
    assert bool((zm[:, 1:] >= zm[:, :-1]).all())
    assert torch.equal(zm, torch.sort(torch.cat([z, zs], -1), -1)[0])
    assert int(inds.min()) >= 1 and int(inds.max()) <= S - 1       # u in [0,1], cdf[0] = 0
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=S - 2)
    lo, hi = torch.gather(mid, 1, below), torch.gather(mid, 1, above)
    assert bool(((zs >= lo - 1e-6) & (zs <= hi + 1e-6)).all())     # every sample lies in the bin its index names
    H.assert_close(o["acc"], w.sum(-1), 2e-5, what="acc = sum(weights)")
    assert bool((w >= 0).all()) and float(o["acc"].max()) <= 1.0 + 1e-5
    ow = ops.composite(raw, z, rd, S, white_background=True)
    H.assert_close(ow["rgb"], o["rgb"] + (1.0 - o["acc"])[:, None], 1e-6, what="white background identity")
    # the BLOCKED order (what the decoder writes) gives bit-identical results at this size too
    pad = rf.reshape(n // 8, 8, S // 16, 16, 4).permute(4, 0, 2, 3, 1).reshape(4, -1).contiguous()
    ob = ops.composite(pad, z, rd, S, n_fine=nf, u=u, want_samples=True, want_inds=True, want_weights=True,
                       row_order=ops.ROWS_BLOCKED)
    for k in ("rgb", "acc", "depth", "weights", "inds", "z_samples", "z_merged"):
        assert torch.equal(o[k], ob[k]), k


@pytest.mark.parametrize("precision", [NVSR_F32, NVSR_F16])
def test_gather_constant_planes_and_linearity(precision):
    """Bilinear gather properties that hold at any size (1.3 M points here): constant planes interpolate to the
    constant (weights sum to 1), and fp32 gathering is linear in the planes."""
    torch.manual_seed(22)
    mc, _, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=1, device=DEV)
    n, S = 20480, 64
    pose, focal = scene.blender_camera(800)
    ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(DEV))
    ro, rd = ro.reshape(-1, 3)[250 * 800:250 * 800 + n].contiguous(), rd.reshape(-1, 3)[250 * 800:250 * 800 + n].contiguous()
    t_vals = torch.linspace(0, 1, S, device=DEV)
    layout = ops.FEAT_LAYOUT[precision]
    names = [scene.get_plane_name(sid, d) for d in range(3)]
    keep = {k: mc.planes_[k].detach().clone() for k in names}
    const = torch.linspace(-2.0, 2.0, 48, device=DEV).reshape(1, 48, 1, 1)

    def gather():
        scene.clear_caches()
        packed = scene.pack_scene_planes(mc, sid, precision)
        fp, fm, _ = ops.sample_gather(ro, rd, 2.0, 6.0, packed, layout, t_vals=t_vals)
        if precision == NVSR_F32:
            return fp, fm                                   # row-major fp32 [rows, 3C] / [rows, C]
        return _untile(fp, n, S).float(), _untile(fm, n, S).float()

    with torch.no_grad():
        for k in names:
            mc.planes_[k].copy_(const.expand_as(mc.planes_[k]))
        fp, fm = gather()
        ref = const.reshape(48).to(torch.float16 if precision == NVSR_F16 else torch.float32).float()
        tol = 2e-3 if precision == NVSR_F16 else 1e-6
        H.assert_close(fp, ref.repeat(3).expand_as(fp), tol, what="constant planes -> constant features")
        H.assert_close(fm, ref.expand_as(fm), tol, what="constant planes -> constant mean")
        if precision == NVSR_F32:
            a = {k: torch.randn_like(keep[k]) for k in names}
            b = {k: torch.randn_like(keep[k]) for k in names}
            outs = []
            for planes in (a, b, {k: a[k] + b[k] for k in names}):
                for k in names:
                    mc.planes_[k].copy_(planes[k])
                outs.append(gather()[0])
            H.assert_close(outs[2], outs[0] + outs[1], 2e-6, what="gather(A + B) = gather(A) + gather(B)")
        for k in names:
            mc.planes_[k].copy_(keep[k])
    scene.clear_caches()


def test_ipe_golden():
    g = golden("stage_ipe.npz")
    enc = ops.ipe(T(g["z"], DEV), T(g["ro"], DEV), T(g["rd"], DEV), float(g["radius"]), 6)
    H.assert_close(enc, g["enc"].reshape(-1, 36), 5e-6, what="ipe")
    for layout, tol in ((FEAT_TILE_BF16, 4e-3), (FEAT_TILE_F16, 5e-4)):
        tile = ops.ipe(T(g["z"], DEV), T(g["ro"], DEV), T(g["rd"], DEV), float(g["radius"]), 6, layout, 48)
        t_, kc_, r_, e_ = tile.shape   # IPE tile images are ray-major
        un = tile.permute(0, 2, 1, 3).reshape(t_ * r_, kc_ * e_)[:180].float()
        H.assert_close(un[:, :36], g["enc"].reshape(-1, 36), tol, what="ipe 16-bit tile")
        assert float(un[:, 36:].abs().max()) == 0.0
    H.assert_close(ops.dir_encoding(T(g["viewdirs"], DEV), 4, True), g["dir_enc"], 2e-6, what="dir enc")


def test_invalid_arguments_fail_loudly():
    with pytest.raises(nvsr_b200.NvsrError):
        ops.composite(torch.zeros(4, 64, device=DEV), torch.zeros(1, 2000, device=DEV), torch.zeros(1, 3, device=DEV), 2000)
    with pytest.raises(nvsr_b200.NvsrError):
        ops.pack_weight16(torch.zeros(128, 48, device=DEV), k_pad=40)


# ---- the same-signature stage drop-ins SURVEY.md 8(b) lists -----------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
def test_planes_model_forward_dropin(prec):
    """`nvsr_b200.planes_model_forward(model, x[n,6]) -> [n,4]` == TwoDimPlanesModel.forward (models.py:381-421) on the
    reference-generated stage golden: fp32 within the golden's own tolerance, 16-bit modes within the per-mode raw
    bounds of the parity gate."""
    import nvsr_b200
    import parity_attribution as PA
    g = golden("stage_planes_forward.npz")
    sid = str(g["scene_id"])
    mc, _ = H.load_planes_scene("scene_planes_small.npz", sid, DEV)
    mc.set_cur_scene_id(sid)
    nvsr_b200.set_precision(prec)
    try:
        with torch.no_grad():
            out = nvsr_b200.planes_model_forward(mc, T(g["x6"], DEV))
    finally:
        nvsr_b200.set_precision("fp16")
    want = T(g["out"])
    assert out.shape == want.shape
    bd = PA.BOUNDS[prec]
    d = (out.cpu() - want).abs()
    assert float(d[:, 3].max()) <= bd["sigma"] and float(d[:, :3].max()) <= bd["logit"], (float(d[:, 3].max()), float(d[:, :3].max()))


def test_proj_combination_sum():
    """proj_combination='sum' (models.py:355-357): density features = sum of the three projections, in every mode."""
    import copy
    import nvsr_b200
    import parity_attribution as PA
    mc, mf, sid = scene.make_synthetic_scene(plane_res=40, view_res=12, seed=5, device=DEV)
    mc.proj_combination = "sum"
    mc.set_cur_scene_id(sid)
    g = torch.Generator().manual_seed(2)
    x6 = torch.cat([torch.rand(777, 3, generator=g) * 2.4 - 1.2, torch.nn.functional.normalize(torch.randn(777, 3, generator=g), dim=-1)], -1)
    mo = copy.deepcopy(mc).cpu()
    with torch.no_grad():
        want = O.planes_model_forward(mo, x6)
        for prec in ("fp32", "fp16", "bf16"):
            nvsr_b200.set_precision(prec)
            out = nvsr_b200.planes_model_forward(mc, x6.to(DEV)).cpu()
            bd = PA.BOUNDS[prec]
            # sigma of a 'sum' model is ~3x the 'avg' one's scale: bound relative to that
            d = (out - want).abs()
            assert float(d[:, 3].max()) <= 3 * bd["sigma"] and float(d[:, :3].max()) <= bd["logit"], (prec, float(d[:, 3].max()))
    nvsr_b200.set_precision("fp16")
    mc.proj_combination = "avg"
    with torch.no_grad():
        nvsr_b200.set_precision("fp32")
        avg = nvsr_b200.planes_model_forward(mc, x6.to(DEV)).cpu()
        nvsr_b200.set_precision("fp16")
    assert float((avg[:, 3] - want[:, 3]).abs().max()) > 1e-2      # the switch really changes the density input


def test_cast_rays_and_ipe_dropins():
    """mip.cast_rays (mip.py:9-18) and IntegratedPositionalEncoding.forward((means, covs)) (mip.py:164-191) with the
    reference's own tensors at the boundary, against the reference-generated stage golden."""
    import nvsr_b200
    g = golden("stage_ipe.npz")
    z, ro, rd = T(g["z"], DEV), T(g["ro"], DEV), T(g["rd"], DEV)
    n = z.shape[0]
    radius = float(g["radius"])
    for radii in (radius, torch.full((n, 1), radius, device=DEV)):       # scalar, or the reference's constant column
        means, covs = nvsr_b200.cast_rays(z, ro, rd, radii, None)
        H.assert_close(means, g["means"], 2e-6, what="means")
        H.assert_close(covs, g["covs"], 1e-7, 1e-5, what="covs")
    enc = nvsr_b200.IntegratedPositionalEncoding(3, 7)
    assert enc.max_freq == 6 and enc.out_dims == 36
    out = enc((T(g["means"], DEV), T(g["covs"], DEV)))
    H.assert_close(out, g["enc"], 5e-6, what="ipe((means, covs))")
    assert out.shape == tuple(g["enc"].shape)


@pytest.mark.gpu
@pytest.mark.parametrize("sa,sb,n", [(64, 128, 1000), (24, 40, 37), (1, 0, 5), (128, 256, 129), (65, 130, 64), (200, 312, 9)])
def test_sort_cat_equals_torch_sort(sa, sb, n):
    """nvsr_sort_cat == torch.sort(torch.cat((a, b), -1), -1).values bit for bit (train_utils.py:144-156), including
    duplicates, negative values, infinities and a NaN (sorted last)."""
    g = torch.Generator().manual_seed(sa * 1000 + sb)
    a = torch.sort(2.0 + 4.0 * torch.rand(n, sa, generator=g), -1).values
    b = 2.0 + 4.0 * torch.rand(n, sb, generator=g)
    if sb:
        b[:, : sb // 3] = a[:, :1]                     # duplicates of a coarse depth
        b[0, -1] = float("inf")
        if n > 2:
            b[1, 0], b[2, 0] = -3.5, float("nan")
    a, b = a.to(DEV), b.to(DEV)
    got = ops.sort_cat(a, b)
    want = torch.sort(torch.cat((a, b), -1), -1).values
    assert got.shape == want.shape
    assert torch.equal(got.nan_to_num(nan=123.0), want.nan_to_num(nan=123.0))
    assert torch.equal(got.isnan(), want.isnan())


@pytest.mark.gpu
@pytest.mark.parametrize("channels,n,S", [(48, 61, 40), (32, 64, 64)])
def test_sample_gather_hilo_vs_fp32_gather(channels, n, S):
    """nvsr_sample_gather_hilo (the fp16-split mode's gather): featP is the 16-bit gather's featP bit for bit; the fp32
    tile image of the combined features, interpolated from the planes' hi + lo fp16 halves, equals the fp32 gather's
    combined features to a few 1e-7 of the feature scale (hi + lo carries ~22 bits of each texel); padding rows are 0."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=40, view_res=8, channels=channels, seed=9, device=DEV)
    g = torch.Generator().manual_seed(5)
    ro = (torch.randn(n, 3, generator=g) * 0.3).to(DEV)
    rd = torch.randn(n, 3, generator=g).to(DEV)
    z = torch.sort(0.2 + 2.5 * torch.rand(n, S, generator=g), -1).values.to(DEV)
    p32 = scene.pack_scene_planes(mf, sid, NVSR_F32)
    p16, lo = scene.pack_scene_planes_hilo(mf, sid)
    fp, fm32, z2 = ops.sample_gather_hilo(ro, rd, 0.0, 1.0, p16, lo, z_in=z)
    fp_ref, _, _ = ops.sample_gather(ro, rd, 0.0, 1.0, p16, FEAT_TILE_F16, z_in=z)
    assert torch.equal(fp, fp_ref) and z2 is z or torch.equal(z2, z)
    _, fm_ref, _ = ops.sample_gather(ro, rd, 0.0, 1.0, p32, FEAT_ROWMAJOR_F32, z_in=z, density_only=True)   # [n*S, C] ray-major
    t, c4, r, e = fm32.shape
    rows = fm32.permute(0, 2, 1, 3).reshape(t * r, c4 * e)                     # BLOCKED rows
    ts = -(-S // 16)
    i = torch.arange(t * r, device=DEV)
    ray, smp = (i // 128 // ts) * 8 + (i % 8), (i // 128 % ts) * 16 + (i % 128) // 8
    valid = (ray < n) & (smp < S)
    got = rows[valid]
    want = fm_ref[(ray * S + smp)[valid]]
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-6 * scale, (float((got - want).abs().max()), scale)
    assert not rows[~valid].any()
    # density-only form
    fp0, fm0, _ = ops.sample_gather_hilo(ro, rd, 0.0, 1.0, p16, lo, z_in=z, density_only=True)
    assert fp0 is None and torch.equal(fm0, fm32)
