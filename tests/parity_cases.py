"""Seeded full-size parity cases (BASELINE configs 1, 2, 3b, 4 at their own sampling, on a ray subset the CPU oracle
finishes in seconds).  Models are built on the CPU from fixed seeds and moved to `device`, rays come from the oracle's
get_ray_bundle — so `build_case(name, "cpu")` reproduces bit-identical inputs wherever it runs (GPU box, build
container).  Test infrastructure: imports the oracle."""
import torch

from nvsr_b200 import scene
from oracle import nvsr_oracle as O

CASES = {
    # BASELINE config 2: 800x800 camera, planes 200^2, 64 + 128 samples; 1024 rays spread over the frame
    "big": dict(kind="planes", res=(800, 800), nc=64, nf=128, rays=1024, precisions=("fp32", "fp16", "bf16")),
    # BASELINE config 1: the whole 100x100 frame, 64 coarse samples, no fine pass
    "cfg1": dict(kind="planes", res=(100, 100), nc=64, nf=0, rays=None, precisions=("fp32", "fp16", "bf16")),
    # BASELINE config 4's shape: forward-facing NDC rays 1008x756, near 0 / far 1, 128 + 256 samples
    "cfg4": dict(kind="planes", res=(756, 1008), nc=128, nf=256, rays=256, ndc=True, precisions=("fp32", "fp16", "bf16")),
    # BASELINE config 3b: mip-NeRF IPE + FlexibleNeRFModel, 65 + 129 interval edges
    "mipbig": dict(kind="mip", res=(800, 800), nc=64, nf=128, rays=512, precisions=("fp32", "fp16", "bf16")),
}


_DEFAULT = object()


def build_case(name, device="cpu", rays=_DEFAULT):
    """`rays`: how many rays of the frame to take (None = the whole frame); default = the case's own figure"""
    c = dict(CASES[name])
    if rays is not _DEFAULT:
        c["rays"] = rays
    H, W = c["res"]
    enc = encd = None
    if c["kind"] == "planes":
        mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device="cpu")
        opt = scene.render_options(c["nc"], c["nf"])
    else:
        mc, mf = scene.make_mip_models(seed=0, device="cpu")
        sid = "synth_DS2"
        opt = scene.render_options(c["nc"], c["nf"], mip=True)
    if c.get("ndc"):
        pose, focal = torch.eye(4), 0.8 * W
        scfg = scene.scene_cfg(near=0.0, far=1.0, no_ndc=False)
    else:
        pose, focal = scene.blender_camera(W)
        scfg = scene.scene_cfg()
    ro, rd = O.get_ray_bundle(H, W, focal, pose)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    if c["rays"] is not None:
        idx = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:c["rays"]]
        ro, rd = ro[idx], rd[idx]
    batch = torch.stack([ro, rd], 0).contiguous()
    dev = torch.device(device)
    if dev.type != "cpu":
        mc, mf, batch = mc.to(dev), mf.to(dev), batch.to(dev)
    if c["kind"] == "mip":
        if dev.type == "cpu":
            enc = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
            encd = lambda x: O.positional_encoding(x, 4, True)
        else:
            import nvsr_b200
            enc, encd = nvsr_b200.IntegratedPositionalEncoding(3, 7), object()
    return dict(H=H, W=W, focal=focal, mc=mc, mf=mf, sid=sid, opt=opt, scfg=scfg, batch=batch, enc=enc, encd=encd,
                nc=c["nc"], nf=c["nf"], kind=c["kind"])


def golden_case(name, device="cpu"):
    """One of the reference-generated end-to-end goldens (tests/golden/e2e_*.npz) as a case dict (+ 'randoms', 'golden')."""
    import helpers as H
    g = H.golden(name)
    sid = str(g["scene_id"])
    mip = name.startswith("e2e_mip")
    if mip:
        mc, mf = H.load_mip_scene(str(g["scene_file"]), device)
    else:
        lr = str(g["lr_scene_id"]) if "lr_scene_id" in g else None
        mc, mf = H.load_planes_scene(str(g["scene_file"]), sid, device, lr_scene_id=lr)
    opt, scfg = H.options_from(g, mip), H.scene_cfg_from(g)
    batch = torch.stack([H.T(g["ro"], device).reshape(-1, 3), H.T(g["rd"], device).reshape(-1, 3)], 0)
    enc = encd = None
    if mip:
        if torch.device(device).type == "cpu":
            enc = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
            encd = lambda x: O.positional_encoding(x, 4, True)
        else:
            import nvsr_b200
            enc, encd = nvsr_b200.IntegratedPositionalEncoding(3, 7), object()
    return dict(H=int(g["H"]), W=int(g["W"]), focal=float(g["focal"]), mc=mc, mf=mf, sid=sid, opt=opt, scfg=scfg, batch=batch,
                enc=enc, encd=encd, nc=int(g["num_coarse"]), nf=int(g["num_fine"]), kind="mip" if mip else "planes",
                randoms=H.randoms_from(g, device), golden=g)
