#!/usr/bin/env python
"""Training-step timing for the §8f-rank-1 backward path (round-2 tool; not the bench.py metric).

One step = the reference's train iteration on the render path (train_nerf.py:860-905): 4 096 rays
(config/TrainModels.yml:8), 64 coarse + 128 fine samples with perturbation, mse on rgb_coarse + rgb_fine, backward to
the tri-planes and both decoders.  Two arms, same scene, same rays, same random draws:
  nvsr   nvsr_b200.autograd.run_one_iter_of_nerf: gather / compositing forward+backward kernels and the decoder's forward,
         data gradient and weight gradients on tcgen05 ('tc', default); also timed in its fp32 parity mode (decoder on torch)
  torch  the same step written with stock PyTorch ops on the GPU (F.grid_sample, nn.Linear, cumprod, searchsorted) —
         what the reference's own code executes on a CUDA device
Prints one JSON object with ms per step (CUDA events, after warm-up), the gradient agreement between the arms and the
per-kernel times of ours.

    python scripts/bench_train_step.py [--rays 4096] [--steps 20] [--plane-res 200]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvsr_b200  # noqa: E402
from nvsr_b200 import autograd as A, ops, scene  # noqa: E402


def torch_planes_forward(model, sid, ro, rd, z, vd):
    """TwoDimPlanesModel.forward with stock torch ops (models.py:381-421)."""
    n, S = z.shape
    model.set_cur_scene_id(sid)
    box = model.box_coords[sid].to(ro)
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
    el = torch.atan2(vd[:, 2], torch.sqrt((vd[:, :2] ** 2).sum(-1)))
    az = torch.atan2(vd[:, 1], vd[:, 0])
    c = torch.cat([pts, torch.stack([az, el], -1)[:, None, :].expand(n, S, 2).reshape(-1, 2)], -1)
    cn = 2 * (c - box[:1]) / (box[1:] - box[:1]) - 1
    rots = model.coord_projector.rot_mats_NON_LEARNED
    pos = []
    for d in range(3):
        grid = (cn[:, :3] @ rots[d][:, 1:].to(cn)).reshape(1, -1, 1, 2)
        pos.append(F.grid_sample(model.planes(d, False), grid, mode="bilinear", align_corners=True, padding_mode="border")[0, :, :, 0].t())
    view = F.grid_sample(model.planes(3, False), cn[:, 3:].reshape(1, -1, 1, 2), mode="bilinear", align_corners=True,
                         padding_mode="border")[0, :, :, 0].t()
    h = torch.stack(pos, 0).mean(0)
    for lin in model.density_dec["0"]:
        h = torch.relu(lin(h))
    alpha = model.fc_alpha["0"](h)
    h = torch.cat(pos + [view], 1)
    for lin in model.rgb_dec["0"]:
        h = torch.relu(lin(h))
    return torch.cat([model.fc_rgb["0"](h), alpha], -1).reshape(n, S, 4)


def torch_render(rf, z, rd, white):
    """volume_render_radiance_field with stock torch ops (volume_rendering_utils.py:15-51)."""
    dists = torch.cat((z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)), -1) * rd.norm(dim=-1, keepdim=True)
    alpha = 1.0 - torch.exp(-torch.relu(rf[..., 3]) * dists)
    T = torch.roll(torch.cumprod(1.0 - alpha + 1e-10, -1), 1, -1)
    T = torch.cat((torch.ones_like(T[..., :1]), T[..., 1:]), -1)
    w = alpha * T
    rgb = (w[..., None] * torch.sigmoid(rf[..., :3])).sum(-2)
    if white:
        rgb = rgb + (1.0 - w.sum(-1, keepdim=True))
    return rgb, w


def torch_step(mc, mf, sid, ro, rd, vd, z, u, white):
    rf = torch_planes_forward(mc, sid, ro, rd, z, vd)
    rgb_c, w = torch_render(rf, z, rd, white)
    with torch.no_grad():   # sample_pdf_2 (nerf_helpers.py:668-702) + sort-merge (train_utils.py:144-156)
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        wt = w[..., 1:-1] + 1e-5
        cdf = torch.cumsum(wt / wt.sum(-1, keepdim=True), -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1).contiguous()
        inds = torch.searchsorted(cdf, u.contiguous(), side="right")
        below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
        cb, ca, bb, ba = cdf.gather(1, below), cdf.gather(1, above), mid.gather(1, below), mid.gather(1, above)
        den = ca - cb
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        zf = torch.sort(torch.cat((z, bb + (u - cb) / den * (ba - bb)), -1), -1).values
    rgb_f, _ = torch_render(torch_planes_forward(mf, sid, ro, rd, zf, vd), zf, rd, white)
    return rgb_c, rgb_f


def timed_replay(graphed, args):
    for _ in range(args.warmup):
        graphed()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graphed()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--plane-res", type=int, default=200)
    ap.add_argument("--no-torch", action="store_true", help="only the eager 'tc' arm (for a profiler's launch list)")
    ap.add_argument("--quick", action="store_true",
                    help="only the eager / graphed 'tc' arms and the stock-PyTorch arm (bench.py's `train_step` block)")
    args = ap.parse_args(argv)
    res = measure(args)
    print(json.dumps(res))
    return res


def measure(args):
    dev = torch.device("cuda", torch.cuda.current_device())
    mc, mf, sid = scene.make_synthetic_scene(plane_res=args.plane_res, view_res=32, seed=0, device=dev)
    for m in (mc, mf):
        m.box_coords = {k: v.to(dev) for k, v in m.box_coords.items()}
    pose, focal = scene.blender_camera(800)
    Nc, Nf = 64, 128
    opt, scfg = scene.render_options(Nc, Nf, perturb=True), scene.scene_cfg(2.0, 6.0, True)
    with torch.no_grad():
        ro_all, rd_all = nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(dev))
    g = torch.Generator().manual_seed(0)
    pick = torch.randperm(800 * 800, generator=g)[:args.rays].to(dev)
    ro, rd = ro_all.reshape(-1, 3)[pick].contiguous(), rd_all.reshape(-1, 3)[pick].contiguous()
    batch = torch.stack([ro, rd], 0)
    target = torch.rand(args.rays, 3, generator=g).to(dev)
    rnd = {"t_rand": torch.rand(args.rays, Nc, generator=g).to(dev), "u": torch.rand(args.rays, Nf, generator=g).to(dev)}
    params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())

    def zero():
        for p in params:
            p.grad = None

    def nvsr_arm():
        out = A.run_one_iter_of_nerf(800, 800, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms=rnd)
        (F.mse_loss(out[0], target) + F.mse_loss(out[3], target)).backward()

    vd = rd / rd.norm(dim=-1, keepdim=True)
    t = torch.linspace(0.0, 1.0, Nc).to(dev)
    zc = (2.0 * (1.0 - t) + 6.0 * t).expand(args.rays, Nc)
    mids = 0.5 * (zc[..., 1:] + zc[..., :-1])
    upper, lower = torch.cat((mids, zc[..., -1:]), -1), torch.cat((zc[..., :1], mids), -1)
    zc = (lower + (upper - lower) * rnd["t_rand"]).contiguous()

    def torch_arm():
        rgb_c, rgb_f = torch_step(mc, mf, sid, ro, rd, vd, zc, rnd["u"], False)
        (F.mse_loss(rgb_c, target) + F.mse_loss(rgb_f, target)).backward()

    def timed(fn):
        for _ in range(args.warmup):
            zero(), fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            zero(), fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    res = {"rays": args.rays, "samples": [Nc, Nf], "plane_res": args.plane_res}
    A.set_decoder("tc")          # decoder forward + backward on tcgen05 (the default of the differentiable path)
    res["nvsr_ms"] = timed(nvsr_arm)
    if args.no_torch:
        return res
    g_n = [None if p.grad is None else p.grad.clone() for p in params]
    # the same step captured once into a CUDA graph and replayed (autograd.GraphedStep): no host enqueue cost
    graphed = A.GraphedStep(lambda: (zero(), nvsr_arm()))   # .grad set to None inside: the capture allocates the gradients
    res["nvsr_graph_ms"] = timed_replay(graphed, args)
    g_g = [None if p.grad is None else p.grad.clone() for p in params]
    if getattr(args, "quick", False):
        res["torch_ms"] = timed(torch_arm)
        g_t = [None if p.grad is None else p.grad.clone() for p in params]
        res["max_rel_l2_grad_diff_tc_vs_torch"] = max(float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
                                                      for a, b in zip(g_n, g_t) if a is not None and b is not None)
        res["what"] = ("one training iteration of the render path (train_nerf.py:860-905): 4 096 rays, 64 + 128 samples with "
                       "perturbation, mse on both passes, backward to planes and decoders; nvsr = nvsr_b200.autograd (tcgen05 "
                       "decoder forward + backward), graph = the same step replayed from a CUDA graph, torch = stock PyTorch "
                       "ops on the same GPU")
        return res
    A.set_sparse_backward(False)
    res["nvsr_dense_backward_ms"] = timed(nvsr_arm)
    A.set_sparse_backward(True)
    # the drop-in flow (no `randoms` from the caller): draws on the CPU like the reference + upload, or on the device
    def dropin_arm():
        out = A.run_one_iter_of_nerf(800, 800, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg)
        (F.mse_loss(out[0], target) + F.mse_loss(out[3], target)).backward()
    res["nvsr_cpu_rng_ms"] = timed(dropin_arm)
    A.set_device_rng(True)
    res["nvsr_device_rng_ms"] = timed(dropin_arm)
    A.set_device_rng(False)
    A.set_decoder("fp32")        # fp32 parity mode: gather / compositing kernels + the model's nn.Linear under torch autograd
    res["nvsr_fp32_mode_ms"] = timed(nvsr_arm)
    g_f = [None if p.grad is None else p.grad.clone() for p in params]
    A.set_decoder("tc")
    res["torch_ms"] = timed(torch_arm)
    g_t = [None if p.grad is None else p.grad.clone() for p in params]

    def rel_l2(x, y):
        return max(float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)) for a, b in zip(x, y) if a is not None and b is not None)
    # 'tc' with the fp32-grade density forward (set_precise_density): gradients against the same stock-PyTorch step
    A.set_precise_density(True)
    res["nvsr_precise_density_ms"] = timed(nvsr_arm)
    g_p = [None if p.grad is None else p.grad.clone() for p in params]
    graphed_p = A.GraphedStep(lambda: (zero(), nvsr_arm()))
    res["nvsr_precise_density_graph_ms"] = timed_replay(graphed_p, args)
    A.set_precise_density(False)
    res["max_rel_l2_grad_diff_tc_precise_vs_torch"] = rel_l2(g_p, g_t)
    res["max_rel_l2_grad_diff_graph_vs_eager"] = rel_l2(g_g, g_n)
    res["max_rel_l2_grad_diff_tc_vs_torch"] = rel_l2(g_n, g_t)
    res["max_rel_l2_grad_diff_fp32_mode_vs_torch"] = rel_l2(g_f, g_t)
    ops.PROFILE = []
    zero(), nvsr_arm()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for name, a, b, meta in prof:
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
    res["kernels_ms"] = {k: {"launches": v[0], "total_ms": v[1]} for k, v in agg.items()}
    return res


if __name__ == "__main__":
    main()
