"""GPU parity, end to end through the reference-shaped call surface (run_one_iter_of_nerf /
eval_nerf), against reference-generated goldens (small scenes) and the CPU oracle (full-size scene,
ray subset), plus size-independent properties at BASELINE config-2 size (800x800, 64+128)."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import T, golden
from oracle import nvsr_oracle as O
from test_oracle_golden import E2E, NAMES, run_oracle_e2e

import nvsr_b200
from nvsr_b200 import ops, scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# stated tolerances (north_star): fp32-accumulate mode 1e-3 abs on rgb/acc, depth 1e-3*(far-near);
# bf16 mode: features, weights and hidden activations are rounded to bf16 (fp32 accumulate)
FP32_TOL = 1e-3
BF16_TOL = 3e-2


def _check(name, g, out, tol):
    worst = {}
    for k, v in zip(NAMES, out[:6]):
        if k not in g:
            assert v is None
            continue
        if "disp" in k:
            # disp = 1/max(1e-10, depth/acc): unbounded; compare where the reference value is moderate
            ref = T(g[k])
            assert torch.equal(torch.isnan(v.cpu()), torch.isnan(ref)), f"{name}:{k} NaN pattern"
            m = ~torch.isnan(ref) & (ref.abs() < 10)
            d = ((v.cpu() - ref).abs() / (1 + ref.abs()))[m]
            worst[k] = float(d.max()) if d.numel() else 0.0
            assert worst[k] <= 10 * tol, (name, k, worst[k])
        else:
            d = (v.cpu() - T(g[k])).abs()
            worst[k] = float(d.max())
            assert worst[k] <= tol, (name, k, worst[k])
    print(name, {k: "%.2e" % v for k, v in worst.items()})


@pytest.mark.parametrize("name", E2E)
def test_e2e_fp32_vs_reference_golden(name):
    nvsr_b200.set_precision("fp32")
    g, out = run_oracle_e2e(name, DEV, runner=nvsr_b200.run_one_iter_of_nerf)
    _check(name, g, out, FP32_TOL)
    assert out[6] is None and out[7] is None and out[8] is None


@pytest.mark.parametrize("name", [n for n in E2E if "mip" not in n])
def test_e2e_bf16_vs_reference_golden(name):
    nvsr_b200.set_precision("bf16")
    g, out = run_oracle_e2e(name, DEV, runner=nvsr_b200.run_one_iter_of_nerf)
    _check(name, g, out, BF16_TOL)


def test_trace_indices_bit_exact_given_same_weights():
    """Bin indices along the real pipeline: feed the GPU's own coarse weights and z to the oracle's
    sample_pdf; indices must agree except where u is within 2 ulp of a cdf edge."""
    nvsr_b200.set_precision("fp32")
    tr = {}
    g, out = run_oracle_e2e("e2e_planes_det.npz", DEV, runner=nvsr_b200.run_one_iter_of_nerf, trace=tr)
    z, w = tr["z_coarse"].cpu(), tr["weights_coarse"].cpu()
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    smp, inds, cdf = O.sample_pdf(mid, w[:, 1:-1], int(g["num_fine"]), det=True, return_all=True)
    mism = tr["inds"].cpu() != inds
    u = torch.linspace(0, 1, int(g["num_fine"]))[None].expand_as(mism)
    for r, j in zip(*torch.nonzero(mism, as_tuple=True)):
        assert float((cdf[r] - u[r, j]).abs().min()) <= 2.4e-7
    H.assert_close(tr["z_samples"], smp, 2e-6, what="z_samples")
    zf = torch.sort(torch.cat([z, tr["z_samples"].cpu()], -1), -1)[0]
    assert torch.equal(tr["z_fine"].cpu(), zf)


@pytest.fixture(scope="module")
def big_scene():
    mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=DEV)
    pose, focal = scene.blender_camera(800)
    return mc, mf, sid, pose.to(DEV), focal


def test_full_size_subset_vs_oracle(big_scene):
    """config-2 scene (R=200, 64+128): 1024 rays spread over the 800x800 frame vs the CPU oracle."""
    mc, mf, sid, pose, focal = big_scene
    opt, scfg = scene.render_options(64, 128), scene.scene_cfg()
    ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose)
    idx = torch.randperm(640000, generator=torch.Generator().manual_seed(0))[:1024].to(DEV)
    batch = torch.stack([ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx]], 0)
    import copy
    mc_c, mf_c = copy.deepcopy(mc).cpu(), copy.deepcopy(mf).cpu()
    with torch.no_grad():
        ref = O.run_one_iter_of_nerf(800, 800, focal, mc_c, mf_c, batch.cpu(), opt, sid, "validation", scene_config=scfg)
        stats = {}
        for prec, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
            nvsr_b200.set_precision(prec)
            out = nvsr_b200.run_one_iter_of_nerf(800, 800, focal, mc, mf, batch, opt, sid, "validation", scene_config=scfg)
            for k, a, b in zip(NAMES, out[:6], ref[:6]):
                if "disp" in k:
                    continue
                d = float((a.cpu() - b).abs().max())
                stats[(prec, k)] = d
                assert d <= tol, (prec, k, d)
    print({f"{p}:{k}": "%.2e" % v for (p, k), v in stats.items()})
    acc = ref[5]
    assert 0.02 < float((acc > 0.5).float().mean()) < 0.98   # the synthetic scene is not degenerate


def test_full_frame_properties(big_scene):
    """800x800, 64+128 (BASELINE config 2): chunk-size invariance, row-band sharding == full frame,
    eval_nerf shape contract."""
    mc, mf, sid, pose, focal = big_scene
    opt, scfg = scene.render_options(64, 128), scene.scene_cfg()
    nvsr_b200.set_precision("bf16")
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose)
        nvsr_b200.set_ray_chunk(32768)
        full = nvsr_b200.eval_nerf(800, 800, focal, mc, mf, ro, rd, opt, sid, scene_config=scfg)
        assert full[0].shape == (800, 800, 3) and full[3].shape == (800, 800, 3)
        assert all(full[i] is None for i in (1, 2, 4, 5, 6, 7, 8))
        assert bool(torch.isfinite(full[3]).all())
        nvsr_b200.set_ray_chunk(20000)   # ragged chunks
        again = nvsr_b200.eval_nerf(800, 800, focal, mc, mf, ro, rd, opt, sid, scene_config=scfg)
        assert torch.equal(full[0], again[0]) and torch.equal(full[3], again[3])
        nvsr_b200.set_ray_chunk(32768)
        band = nvsr_b200.render_frame(800, 800, focal, pose, mc, mf, opt, sid, scfg, row_range=(300, 400))
        assert torch.equal(band[3].reshape(100, 800, 3), full[3][300:400])
