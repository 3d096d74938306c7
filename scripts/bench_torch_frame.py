#!/usr/bin/env python
"""Informative second baseline ("optionally also time the reference CUDA path on 1 B200"): the 800x800
frame of bench.py rendered with STOCK PyTorch ops on the GPU — F.grid_sample, nn.Linear, cumprod, searchsorted, sort —
i.e. what the reference's own code executes on a CUDA device, in ray chunks like its `chunksize` (train_utils.py:228-234).
Not the bench.py metric and not a parity oracle (tests/ hold those): a round-2 tool that answers "what do the kernels
displace on this very GPU".

    python scripts/bench_torch_frame.py [--res 800] [--chunk 16384] [--frames 3]     # prints one JSON line
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvsr_b200  # noqa: E402
from nvsr_b200 import scene  # noqa: E402
from bench_train_step import torch_step  # noqa: E402  (the same step, written with stock torch ops)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--chunk", type=int, default=16384, help="rays per chunk (the reference chunks at 131 072 POINTS)")
    ap.add_argument("--frames", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev)
    for m in (mc, mf):
        m.box_coords = {k: v.to(dev) for k, v in m.box_coords.items()}
    pose, focal = scene.blender_camera(args.res)
    Nc, Nf = 64, 128
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(args.res, args.res, focal, pose.to(dev))
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        t = torch.linspace(0.0, 1.0, Nc).to(dev)
        z_row = 2.0 * (1.0 - t) + 6.0 * t
        u_row = torch.linspace(0.0, 1.0, Nf).to(dev)

        def frame():
            out = []
            for i in range(0, ro.shape[0], args.chunk):
                n = min(args.chunk, ro.shape[0] - i)
                rgb_c, rgb_f = torch_step(mc, mf, sid, ro[i:i + n], rd[i:i + n], vd[i:i + n], z_row.expand(n, Nc).contiguous(),
                                          u_row.expand(n, Nf).contiguous(), False)
                out.append(rgb_f)
            return torch.cat(out, 0)

        frame()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.frames):
            img = frame()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.frames
        ours = nvsr_b200.render_frame(args.res, args.res, focal, pose.to(dev), mc, mf, scene.render_options(Nc, Nf), sid, scene.scene_cfg())
        diff = (ours[3] - img).abs()
    print(json.dumps({"baseline": "stock PyTorch ops on the GPU (reference algorithm)", "res": args.res, "chunk_rays": args.chunk,
                      "ms_per_frame": ms, "rays_per_s": args.res * args.res / ms * 1e3,
                      "rgb_fine_vs_nvsr_mean_abs": float(diff.mean()), "rgb_fine_vs_nvsr_p99_abs": float(diff.flatten().quantile(0.99))}))


if __name__ == "__main__":
    main()
