"""GPU parity of the plane super-resolution step: `nvsr_b200.sr.PlaneSuperResolver` — EDSR conv
chain (cuDNN, channels-last) + the hand-written `nvsr_sr_finalize` kernel that writes the gather's plane image directly —
against planes super-resolved by the reference's own PlanesSR + EDSR (tests/golden/sr_*.npz, make_golden_sr.py), and end
to end through the render path against the oracle (which restates PlanesSR.forward; pinned to the same goldens on the CPU)."""
import pytest
import torch

import helpers as H
import nvsr_b200
import parity_attribution as PA
from nvsr_b200 import NVSR_BF16, NVSR_F16, NVSR_F32, scene, sr as SR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["sr_small.npz", "sr_x2_norm.npz", "sr_ragged.npz"])
def test_super_resolve_matches_reference(name):
    sr, g, want = H.load_sr_model(name, DEV)
    res = SR.resolver_of(sr)
    for pname, ref in want.items():
        packed, nchw = res.super_resolve(pname, NVSR_F32, want_nchw=True)
        H.assert_close(nchw, ref, 2e-5, 1e-4, what=f"{pname} fp32 NCHW")           # cuDNN fp32 vs ATen CPU summation order
        assert torch.equal(packed, nchw[0].permute(1, 2, 0).contiguous())          # channels-last image == the same values
        c, rh, rw = ref.shape[1:]
        for dtype, tdt, tol in ((NVSR_F16, torch.float16, 2e-2), (NVSR_BF16, torch.bfloat16, 1.2e-1)):
            img, _ = res.super_resolve(pname, dtype)
            assert img.shape == (rh, c // 8, rw, 2, 8) and img.dtype == tdt
            left = img[..., 0, :].float().permute(1, 3, 0, 2).reshape(c, rh, rw).cpu()     # [C/8,8,Rh,Rw] -> [C,Rh,Rw]
            d = (left - ref[0]).abs() / (1 + ref[0].abs())
            assert float(d.max()) <= tol, (pname, dtype, float(d.max()))
            # x-pair records: the right half is the same chunk of the texel to the right (last column paired with itself)
            right = img[..., 1, :]
            shifted = img[..., 0, :][:, :, torch.clamp(torch.arange(rw, device=DEV) + 1, max=rw - 1)]
            assert torch.equal(right, shifted)


def test_sr_plane_is_cached_per_plane_version():
    sr, g, want = H.load_sr_model("sr_small.npz", DEV)
    res = SR.resolver_of(sr)
    pname = next(iter(want))
    a, _ = res.super_resolve(pname, NVSR_F16)
    b, _ = res.super_resolve(pname, NVSR_F16)
    assert a is b                                   # device-resident: computed once, never re-uploaded
    sr.LR_planes[pname].mul_(1.5)                   # an optimizer step bumps the version
    c, _ = res.super_resolve(pname, NVSR_F16)
    assert c is not a and not torch.equal(c, a)


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16", "fp16-split"])
def test_render_with_native_sr_vs_oracle(prec):
    """BASELINE config 3a's structure at test size: the fine model reads planes super-resolved by an EDSR-shaped SR
    model (x2, hidden 32, 2 blocks) on the device; the coarse model reads the LR planes.  Whole chain through the
    explained-outlier gate against the oracle, which super-resolves on the CPU with its restatement of PlanesSR."""
    mc, mf, sid = scene.make_synthetic_scene(plane_res=32, view_res=8, seed=21, device=DEV, sr_scale=2, sr_hidden=32,
                                             sr_blocks=2)
    pose, focal = scene.blender_camera(30)
    opt, scfg = scene.render_options(32, 48), scene.scene_cfg()
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(30, 30, focal, pose.to(DEV), 0, 0.25)
    batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    c = dict(H=30, W=30, focal=focal, mc=mc, mf=mf, sid=sid, opt=opt, scfg=scfg, batch=batch, enc=None, encd=None, kind="planes")
    nvsr_b200.set_precision(prec)
    nvsr_b200.set_sparse_rgb(False)
    try:
        tr = {}
        with torch.no_grad():
            out = nvsr_b200.run_one_iter_of_nerf(30, 30, focal, mc, mf, batch, opt, sid, "validation", scene_config=scfg, trace=tr)
        rep = PA.check_chain(c, prec, out, tr)
        print(prec, {k: v for k, v in rep.items() if "unexplained" in k})
        # the fine pass really read other planes than the coarse pass
        assert not torch.equal(out[0], out[3])
    finally:
        nvsr_b200.set_sparse_rgb(True)
        nvsr_b200.set_precision("fp16")
