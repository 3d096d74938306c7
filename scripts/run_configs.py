"""BASELINE.json configs 3a / 3b / 4 at FULL size on one B200 (config 2 is bench.py's workload): frame time, rays/s,
decoder evaluations/s and sanity of the maps.  These are the other rows of SURVEY.md §8a exercised end to end:
a5 with 4x super-resolved 800^2 planes on the fine pass, a9 + a6' (IPE + FlexibleNeRFModel), a3 (NDC rays).

    python scripts/run_configs.py [--frames 3] > profiles/<round>_configs.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nvsr_b200  # noqa: E402
from nvsr_b200 import scene  # noqa: E402


def timed_frames(fn, frames):
    fn()                       # warm-up: packs planes / weights, SR inference, allocator
    torch.cuda.synchronize()
    ts = []
    for _ in range(frames):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out


def report(name, ms, out, n_rays, nc, nf, extra):
    rgb, acc = out[3], out[5]
    ok = bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(acc).all())
    r = {"config": name, "ms_per_frame": ms, "rays_per_s": n_rays / ms * 1e3,
         "decoder_evals_per_s": n_rays * (nc + nc + nf) / ms * 1e3, "rays": n_rays, "samples": "%d+%d" % (nc, nf),
         "rgb_fine_mean": float(rgb.mean()), "acc_fine_mean": float(acc.mean()),
         "acc_fine_frac_above_half": float((acc > 0.5).float().mean()), "finite": ok}
    r.update(extra)
    print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--precision", default="fp16")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    nvsr_b200.set_precision(args.precision)
    with torch.no_grad():
        # ---- config 3a: fine model reads 4x super-resolved planes (200^2 -> 800^2), coarse reads the LR planes
        mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev, sr_scale=4)
        pose, focal = scene.blender_camera(800)
        pose = pose.to(dev)
        opt, scfg = scene.render_options(64, 128), scene.scene_cfg()
        off = (2 - 1) / (2 * 2)   # downsampling_offset of a DS2 scene (train_nerf.py:610)
        ms, out = timed_frames(lambda: nvsr_b200.render_frame(800, 800, focal, pose, mc, mf, opt, sid, scfg,
                                                              downsampling_offset=off), args.frames)
        report("cfg3a_800x800_64+128_SRplanes_200->800", ms, out, 640000, 64, 128,
               {"planes": "coarse 3x48x200^2, fine 3x48x800^2 (stock-PyTorch SR, once per scene)"})
        del mc, mf
        scene.clear_caches() if hasattr(scene, "clear_caches") else None
        torch.cuda.empty_cache()

        # ---- config 3b: mip-NeRF IPE + FlexibleNeRFModel
        mc, mf = scene.make_mip_models(seed=0, device=dev)
        opt = scene.render_options(64, 128, mip=True)
        enc = nvsr_b200.IntegratedPositionalEncoding(3, 7)
        ms, out = timed_frames(lambda: nvsr_b200.render_frame(800, 800, focal, pose, mc, mf, opt, "synth_DS2", scfg,
                                                              encode_position_fn=enc, encode_direction_fn=object()),
                               args.frames)
        report("cfg3b_800x800_65+129_mip_IPE", ms, out, 640000, 64, 128, {"decoder": "FlexibleNeRFModel 36->128x4 (+27 dir)"})
        del mc, mf
        torch.cuda.empty_cache()

        # ---- config 4: LLFF-shaped forward-facing scene, NDC rays, 1008x756, 128 + 256 samples
        mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev)
        H, W = 756, 1008
        pose4 = torch.eye(4, device=dev)
        opt = scene.render_options(128, 256)
        scfg4 = scene.scene_cfg(near=0.0, far=1.0, no_ndc=False)
        ms, out = timed_frames(lambda: nvsr_b200.render_frame(H, W, 0.8 * W, pose4, mc, mf, opt, sid, scfg4), args.frames)
        report("cfg4_1008x756_128+256_ndc", ms, out, H * W, 128, 256, {"planes": "3x48x200^2"})


if __name__ == "__main__":
    main()
