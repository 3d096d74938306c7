"""Shared test utilities: golden-fixture loading into the stand-in models of nvsr_b200.scene."""
import os

import numpy as np
import torch

import nvsr_b200
from nvsr_b200 import scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def T(a, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class StoredSR(torch.nn.Module):
    """An 'SR model' whose outputs were computed by the reference's PlanesSR+EDSR and stored."""

    def __init__(self, sr_planes):
        super().__init__()
        self.SR_planes = dict(sr_planes)

    def forward(self, name):
        return self.SR_planes[name]


def _load_state(model, g, prefix):
    sd = {}
    for k, v in g.items():
        if k.startswith(prefix):
            sd[k[len(prefix):].replace("__", ".")] = T(v)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "planes_" not in m and "rot_mats" not in m and "SR_model" not in m]
    assert not missing and not unexpected, (missing, unexpected)


def load_planes_scene(scene_file, scene_id, device="cpu", lr_scene_id=None):
    g = golden(scene_file)
    pairs = {scene_id: lr_scene_id} if lr_scene_id else None
    coarse = scene.TriPlaneModel(scene_coupler=scene.SingleSceneCoupler(pairs))
    fine = scene.TriPlaneModel(scene_coupler=scene.SingleSceneCoupler(pairs))
    fine.coord_projector = coarse.coord_projector
    _load_state(coarse, g, "coarse__")
    _load_state(fine, g, "fine__")
    planes = torch.nn.ParameterDict()
    for k, v in g.items():
        if k.startswith("plane__"):
            planes[k[len("plane__"):]] = torch.nn.Parameter(T(v))
    box = torch.tensor(scene.DEFAULT_BOX, dtype=torch.float64)
    for m in (coarse, fine):
        m.planes_ = planes
        m.box_coords = {scene_id: box}
        if lr_scene_id:
            m.box_coords[lr_scene_id] = box
        m.eval()
    coarse.to(device)
    fine.to(device)
    sr = {k[len("sr_plane__"):]: T(v, device) for k, v in g.items() if k.startswith("sr_plane__")}
    if sr:
        fine.assign_SR_model(StoredSR(sr))
    return coarse, fine


def load_mip_scene(scene_file, device="cpu"):
    g = golden(scene_file)
    coarse, fine = scene.MipMLP(), scene.MipMLP()
    _load_state(coarse, g, "coarse__")
    _load_state(fine, g, "fine__")
    return coarse.eval().to(device), fine.eval().to(device)


def options_from(g, mip=False):
    return scene.render_options(num_coarse=int(g["num_coarse"]), num_fine=int(g["num_fine"]),
                                perturb=bool(g["perturb"]), lindisp=bool(g["lindisp"]),
                                white_background=bool(g["white_background"]), noise_std=float(g["noise_std"]), mip=mip)


def scene_cfg_from(g):
    return scene.scene_cfg(float(g["near"]), float(g["far"]), bool(g["no_ndc"]))


def randoms_from(g, device="cpu"):
    return {k: T(g[k], device) for k in ("t_rand", "u", "noise_c", "noise_f") if k in g}


def assert_close(a, b, atol, rtol=0.0, what=""):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu() if torch.is_tensor(b) else torch.from_numpy(np.asarray(b)).float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    assert torch.equal(nan_a, nan_b), f"{what}: NaN pattern differs"
    d = (a - b).abs()[~nan_a]
    tol = atol + rtol * b.abs()[~nan_a]
    if d.numel() and not bool((d <= tol).all()):
        i = int(torch.argmax(d - tol))
        raise AssertionError(f"{what}: max |diff| {float(d.max()):.3e} (tol {atol:g}+{rtol:g}*|ref|), "
                             f"worst at flat index {i}: {float(a[~nan_a][i])} vs {float(b[~nan_a][i])}")


def pdf_sample_tolerance(cdf, bins, inds, eps=2.4e-7):
    """Per-sample bound on |delta z_sample| caused by an `eps` (2 ulp at 1.0) perturbation of the cdf:
    t = (u - cdf_below)/denom, so delta z <= width * 3*eps/denom.  Near-empty bins (denom ~ 1e-5) are
    ill-conditioned in the reference itself; across its `denom < 1e-5 -> 1` switch the sample may move
    by a whole bin."""
    B = cdf.shape[-1]
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=B - 1)
    denom = cdf.gather(1, above) - cdf.gather(1, below)
    width = (bins.gather(1, above) - bins.gather(1, below)).abs()
    switch = (denom - 1e-5).abs() < 2e-6
    d = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    tol = 2e-6 + width * 3 * eps / d
    return torch.where(switch, width + 2e-6, tol)


def check_resampling(inds_gpu, zs_gpu, inds_ref, zs_ref, cdf_ref, bins_ref, u, what="", max_flip_frac=0.02):
    """Bin indices must be bit-exact except where u sits within 2 ulp of a cdf edge ("legit flips");
    samples must agree within the conditioning bound of the reference's own formula."""
    inds_gpu, zs_gpu = inds_gpu.cpu(), zs_gpu.cpu()
    u = u.expand_as(inds_ref) if u.dim() == 2 else u[None].expand_as(inds_ref)
    mism = inds_gpu != inds_ref
    for r, j in zip(*torch.nonzero(mism, as_tuple=True)):
        edge = float((cdf_ref[r] - u[r, j]).abs().min())
        assert edge <= 3.6e-7, f"{what}: index mismatch at ray {int(r)} sample {int(j)} but u is {edge:.2e} from the nearest cdf edge"
    assert float(mism.float().mean()) <= max_flip_frac, (what, float(mism.float().mean()))
    tol = pdf_sample_tolerance(cdf_ref, bins_ref, inds_ref)
    width = (bins_ref[:, 1:] - bins_ref[:, :-1]).abs().max(-1, keepdim=True)[0]
    tol = torch.where(mism, 2 * width.expand_as(tol), tol)
    d = (zs_gpu - zs_ref).abs()
    bad = d > tol
    assert not bool(bad.any()), f"{what}: z_samples off by {float((d - tol).max()):.3e} beyond the conditioning bound"
    return mism


def load_sr_model(golden_name, device="cpu"):
    """The stand-in PlanesSRModel loaded with the EDSR weights of an SR golden (tests/golden/make_golden_sr.py) and its LR
    planes registered -> (sr model, golden dict, {plane name: SR plane of the reference})."""
    g = golden(golden_name)
    sr = scene.PlanesSRModel(int(g["scale"]), int(g["channels"]), int(g["channels"]), int(g["hidden"]), int(g["n_blocks"]))
    sd = {k[len("w__"):].replace("__", "."): T(v) for k, v in g.items() if k.startswith("w__")}
    for k in [k for k in sd if k.startswith("planes_mean_NON_LEARNED") or k.startswith("planes_std_NON_LEARNED")]:
        setattr(sr, k, torch.nn.Parameter(sd.pop(k), requires_grad=False))
    missing, unexpected = sr.load_state_dict(sd, strict=False)
    assert not [m for m in missing if "NON_LEARNED" not in m] and not unexpected, (missing, unexpected)
    sr = sr.to(device).eval()
    want = {}
    for k, v in g.items():
        if k.startswith("lr__"):
            sr.set_LR_plane(T(v, device), k[len("lr__"):])
            want[k[len("lr__"):]] = T(g["sr__" + k[len("lr__"):]])
    return sr, g, want
