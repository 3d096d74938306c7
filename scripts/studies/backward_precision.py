"""Design study for the tensor-core decoder backward (NOTES.md backlog #1) — CPU only, no kernels involved.

Question: in which 16-bit format can the back-propagated deltas (and the stored activations) live so that the parameter
gradients of a training step stay close to the fp32 reference?  The step is the reference's (4 096-ray batches in the
real run, config/TrainModels.yml:8; here a lattice of the synthetic bench scene), loss = mse(rgb_coarse) + mse(rgb_fine).
The decoder's linear layers are wrapped so that, in backward, grad_output and the saved input are rounded to the format
under test (optionally after multiplying the loss by a power-of-two scale, undone on the weight gradients), with fp32
accumulation — the arithmetic a tcgen05 kind::f16 backward would perform.  Everything else stays fp32.

    python scripts/studies/backward_precision.py        # prints one line per format: worst relative gradient error
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import nvsr_b200  # noqa: E402
from nvsr_b200 import scene  # noqa: E402

FORMATS = {"fp32": None, "fp16": torch.float16, "bf16": torch.bfloat16}


class QLinear(torch.autograd.Function):
    """y = x W^T + b with the backward's operands rounded to `dtype` (saturating like cvt.satfinite)."""

    @staticmethod
    def forward(ctx, x, w, b, dtype):
        ctx.save_for_backward(x, w)
        ctx.dtype = dtype
        return torch.nn.functional.linear(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dt = ctx.dtype
        if dt is not None:
            lim = torch.finfo(dt).max
            q = lambda t: t.clamp(-lim, lim).to(dt).float()
            g, x, w = q(g), q(x), q(w)
        return g @ w, g.t() @ x, g.sum(0), None


def decoder(model, feat_m, feat_p, vfeat_rows, dtype):
    h = feat_m
    for lin in model.density_dec["0"]:
        h = torch.relu(QLinear.apply(h, lin.weight, lin.bias, dtype))
    alpha = QLinear.apply(h, model.fc_alpha["0"].weight, model.fc_alpha["0"].bias, dtype)
    h = torch.cat([feat_p, vfeat_rows], 1)
    for lin in model.rgb_dec["0"]:
        h = torch.relu(QLinear.apply(h, lin.weight, lin.bias, dtype))
    return torch.cat([QLinear.apply(h, model.fc_rgb["0"].weight, model.fc_rgb["0"].bias, dtype), alpha], -1)


def render(rf, z, rd):
    dists = torch.cat((z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)), -1) * rd.norm(dim=-1, keepdim=True)
    alpha = 1.0 - torch.exp(-torch.relu(rf[..., 3]) * dists)
    T = torch.cumprod(torch.cat((torch.ones_like(alpha[..., :1]), 1.0 - alpha[..., :-1] + 1e-10), -1), -1)
    w = alpha * T
    return (w[..., None] * torch.sigmoid(rf[..., :3])).sum(-2), w


def features(model, sid, ro, rd, z, vd):
    import torch.nn.functional as F
    n, S = z.shape
    model.set_cur_scene_id(sid)
    box = model.box_coords[sid].float()
    pts = (ro[:, None, :] + rd[:, None, :] * z[..., None]).reshape(-1, 3)
    el = torch.atan2(vd[:, 2], torch.sqrt((vd[:, :2] ** 2).sum(-1)))
    az = torch.atan2(vd[:, 1], vd[:, 0])
    cn = 2 * (pts - box[0, :3]) / (box[1, :3] - box[0, :3]) - 1
    rots = model.coord_projector.rot_mats_NON_LEARNED
    pos = [F.grid_sample(model.planes(d, False), (cn @ rots[d][:, 1:].float()).reshape(1, -1, 1, 2), mode="bilinear",
                         align_corners=True, padding_mode="border")[0, :, :, 0].t() for d in range(3)]
    va = 2 * (torch.stack([az, el], -1) - box[0, 3:]) / (box[1, 3:] - box[0, 3:]) - 1
    view = F.grid_sample(model.planes(3, False), va.reshape(1, -1, 1, 2), mode="bilinear", align_corners=True,
                         padding_mode="border")[0, :, :, 0].t()
    return torch.stack(pos, 0).mean(0), torch.cat(pos, 1), view[:, None, :].expand(n, S, view.shape[-1]).reshape(n * S, -1)


def step(mc, mf, sid, ro, rd, vd, zc, u, target, dtype, loss_scale):
    params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
    for p in params:
        p.grad = None
    rf = decoder(mc, *features(mc, sid, ro, rd, zc, vd), dtype).reshape(zc.shape[0], zc.shape[1], 4)
    rgb_c, w = render(rf, zc, rd)
    with torch.no_grad():
        mid = 0.5 * (zc[..., 1:] + zc[..., :-1])
        wt = w[..., 1:-1] + 1e-5
        cdf = torch.cumsum(wt / wt.sum(-1, keepdim=True), -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1).contiguous()
        inds = torch.searchsorted(cdf, u.contiguous(), side="right")
        below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
        cb, ca, bb, ba = cdf.gather(1, below), cdf.gather(1, above), mid.gather(1, below), mid.gather(1, above)
        den = torch.where(ca - cb < 1e-5, torch.ones_like(ca), ca - cb)
        zf = torch.sort(torch.cat((zc, bb + (u - cb) / den * (ba - bb)), -1), -1).values
    rf_f = decoder(mf, *features(mf, sid, ro, rd, zf, vd), dtype).reshape(zf.shape[0], zf.shape[1], 4)
    rgb_f, _ = render(rf_f, zf, rd)
    loss = torch.nn.functional.mse_loss(rgb_c, target) + torch.nn.functional.mse_loss(rgb_f, target)
    (loss * loss_scale).backward()
    return {id(p): p.grad / loss_scale for p in params if p.grad is not None}, params


def main():
    torch.manual_seed(0)
    mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0)
    pose, focal = scene.blender_camera(64)
    g = torch.Generator().manual_seed(1)
    n, Nc, Nf = 1024, 64, 128
    # rays of a 32x32 lattice of the frame, computed on the host (fp32)
    ii, jj = torch.meshgrid(torch.linspace(0, 63, 32), torch.linspace(0, 63, 32), indexing="xy")
    dirs = torch.stack([(ii - 32) / focal, -(jj - 32) / focal, -torch.ones_like(ii)], -1)
    rd = (dirs[..., None, :] * pose[:3, :3]).sum(-1).reshape(-1, 3)
    ro = pose[:3, 3].expand(rd.shape).contiguous()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    t = torch.linspace(0.0, 1.0, Nc)
    zc = (2.0 * (1.0 - t) + 6.0 * t).expand(n, Nc)
    mids = 0.5 * (zc[..., 1:] + zc[..., :-1])
    upper, lower = torch.cat((mids, zc[..., -1:]), -1), torch.cat((zc[..., :1], mids), -1)
    zc = (lower + (upper - lower) * torch.rand(n, Nc, generator=g)).contiguous()
    u = torch.rand(n, Nf, generator=g)
    target = torch.rand(n, 3, generator=g)
    ref, params = step(mc, mf, sid, ro, rd, vd, zc, u, target, None, 1.0)
    names = {id(p): k for m, tag in ((mc, "coarse."), (mf, "fine.")) for k, p in ((tag + k, p) for k, p in m.named_parameters())}
    print("format      loss_scale   worst rel err (decoder weights)   worst rel err (planes)   worst parameter")
    for fmt in ("fp16", "bf16"):
        for scale in (1.0, 2.0 ** 10, 2.0 ** 16):
            got, _ = step(mc, mf, sid, ro, rd, vd, zc, u, target, FORMATS[fmt], scale)
            worst_w = worst_p = 0.0
            worst_name = ""
            for p in params:
                if id(p) not in ref or float(ref[id(p)].abs().max()) == 0:
                    continue
                e = float((got[id(p)] - ref[id(p)]).abs().max() / ref[id(p)].abs().max())
                if "planes_" in names[id(p)]:
                    worst_p = max(worst_p, e)
                elif e > worst_w:
                    worst_w, worst_name = e, names[id(p)]
            print(f"{fmt:10s}  {scale:10.0f}   {worst_w:30.3e}   {worst_p:22.3e}   {worst_name}")


if __name__ == "__main__":
    main()
