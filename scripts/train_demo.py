#!/usr/bin/env python
"""Functional check of the training path: a student tri-plane scene is fitted to renderings of a
teacher scene with Adam through `nvsr_b200.autograd.run_one_iter_of_nerf` — once with the decoder forward / backward on
tcgen05 (fp16 operands, loss-scaled deltas: the default) and once in the fp32 parity mode (model's nn.Linear under torch
autograd), from the same initial weights, on the same ray batches and the same random draws.  Prints one JSON object with
both loss curves; mixed precision is fit for purpose if the curves track each other.

    python scripts/train_demo.py [--steps 200] [--rays 1024]
"""
import argparse
import copy
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvsr_b200  # noqa: E402
from nvsr_b200 import autograd as A, scene  # noqa: E402


def run(mode, student, sid, batches, targets, randoms, opt_cfg, scfg, res, focal, lr):
    A.set_decoder(mode)
    mc, mf = student
    params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
    optim = torch.optim.Adam(params, lr=lr)
    losses = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i, (b, t, r) in enumerate(zip(batches, targets, randoms)):
        if i == 5:
            e0.record()
        optim.zero_grad(set_to_none=True)
        out = A.run_one_iter_of_nerf(res, res, focal, mc, mf, b, opt_cfg, sid, "train", scene_config=scfg, randoms=r)
        loss = F.mse_loss(out[0], t) + F.mse_loss(out[3], t)
        loss.backward()
        optim.step()
        losses.append(float(loss.detach()))
    e1.record()
    torch.cuda.synchronize()
    return losses, e0.elapsed_time(e1) / max(1, len(batches) - 5)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--plane-res", type=int, default=64)
    ap.add_argument("--lr", type=float, default=2e-3)
    args = ap.parse_args(argv)
    dev = torch.device("cuda", 0)
    res, Nc, Nf = 64, 64, 64
    teacher_c, teacher_f, sid = scene.make_synthetic_scene(plane_res=args.plane_res, view_res=16, seed=1, device=dev)
    student = scene.make_synthetic_scene(plane_res=args.plane_res, view_res=16, seed=2, device=dev)[:2]
    for m in student:
        m.train()
    opt_cfg = scene.render_options(Nc, Nf, perturb=True, noise_std=0.1)
    scfg = scene.scene_cfg()
    g = torch.Generator().manual_seed(0)
    batches, targets, randoms = [], [], []
    nvsr_b200.set_precision("fp32")      # teacher images: the 1e-3 parity mode of the forward path
    with torch.no_grad():
        for i in range(args.steps):
            pose, focal = scene.blender_camera(res, theta=float(torch.rand(1, generator=g)) * 360.0 - 180.0)
            ro, rd = nvsr_b200.get_ray_bundle(res, res, focal, pose.to(dev))
            pick = torch.randperm(res * res, generator=g)[:args.rays].to(dev)
            b = torch.stack([ro.reshape(-1, 3)[pick], rd.reshape(-1, 3)[pick]], 0).contiguous()
            out = nvsr_b200.run_one_iter_of_nerf(res, res, focal, teacher_c, teacher_f, b, scene.render_options(Nc, Nf), sid,
                                                 "validation", scene_config=scfg)
            batches.append(b)
            targets.append(out[3].clone())
            randoms.append(dict(t_rand=torch.rand(args.rays, Nc, generator=g), u=torch.rand(args.rays, Nf, generator=g),
                                noise_c=torch.randn(args.rays, Nc, generator=g), noise_f=torch.randn(args.rays, Nc + Nf, generator=g)))
    nvsr_b200.set_precision("fp16")
    res_out = {"steps": args.steps, "rays": args.rays, "samples": [Nc, Nf], "plane_res": args.plane_res, "lr": args.lr}
    try:
        for mode in ("tc", "fp32"):
            st = tuple(copy.deepcopy(m) for m in student)
            for m_new, m_old in zip(st, student):
                m_new.box_coords = {k: v.clone() for k, v in m_old.box_coords.items()}
            losses, ms = run(mode, st, sid, batches, targets, randoms, opt_cfg, scfg, res, focal, args.lr)
            k = max(1, args.steps // 10)
            res_out[mode] = {"first": sum(losses[:k]) / k, "last": sum(losses[-k:]) / k, "ms_per_step": ms,
                             "curve": [round(sum(losses[i:i + k]) / len(losses[i:i + k]), 6) for i in range(0, args.steps, k)]}
    finally:
        A.set_decoder("tc")
    res_out["last_loss_ratio_tc_over_fp32"] = res_out["tc"]["last"] / res_out["fp32"]["last"]
    print(json.dumps(res_out))
    return res_out


if __name__ == "__main__":
    main()
