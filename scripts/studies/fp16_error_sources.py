"""CPU study: which fp16 roundings of the decoder path make up the map error of the 16-bit mode (coarse pass of the
smoke scene), and what a split (hi + lo) operand scheme would leave.  q(x) = x.half().float()."""
import os, sys, itertools
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import nvsr_b200
from nvsr_b200 import scene
from oracle import nvsr_oracle as O

torch.manual_seed(0)
mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device="cpu")
pose, focal = scene.blender_camera(48)
ro, rd = O.get_ray_bundle(48, 48, focal, pose)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
n, S = ro.shape[0], 64
z = torch.linspace(2.0, 6.0, S).expand(n, S).contiguous()
vd = rd / rd.norm(dim=-1, keepdim=True)
pts = ro[:, None] + rd[:, None] * z[..., None]
x6 = torch.cat([pts, vd[:, None].expand(pts.shape)], -1).reshape(-1, 6)
q = lambda t: t.half().float()

def run(qp, qf, qw, qa, split=False):
    """qp: round planes, qf: features, qw: weights, qa: hidden activations; split: operands are hi+lo pairs (error ~2^-22)"""
    m = mc
    with torch.no_grad():
        if qp:
            import copy
            m = copy.deepcopy(mc)
            for k, p in m.planes_.items():
                p.data = q(p.data)
        m.set_cur_scene_id(sid)
        pos, view = O.planes_gather(m, x6)
        fP, fM = torch.cat(pos, 1), torch.stack(pos, 0).mean(0)
        if qf:
            fP, fM = q(fP), q(fM)
        W = (lambda w: q(w)) if qw else (lambda w: w)
        A = (lambda a: q(a)) if qa else (lambda a: a)
        h = fM
        L = list(m.density_dec["0"])
        for i, lin in enumerate(L):
            h = torch.relu(h @ W(lin.weight).t() + lin.bias)
            if i < len(L) - 1:
                h = A(h)
        sigma = m.fc_alpha["0"](h)
        L = list(m.rgb_dec["0"])
        w0 = L[0].weight
        h = torch.relu(fP @ W(w0[:, :fP.shape[1]]).t() + (view @ w0[:, fP.shape[1]:].t() + L[0].bias))
        h = A(h)
        for i, lin in enumerate(L[1:]):
            h = torch.relu(h @ W(lin.weight).t() + lin.bias)
            if i < len(L) - 2:
                h = A(h)
        rgb = m.fc_rgb["0"](h)
        raw = torch.cat([rgb, sigma], -1).reshape(n, S, 4)
        out = O.volume_render_radiance_field(raw, z, rd, 0.0, False)
    return raw, out

raw0, out0 = run(False, False, False, False)
print("lit fraction", float((raw0[..., 3] > 0).float().mean()))
for name, cfg in [("planes", (1, 0, 0, 0)), ("features", (0, 1, 0, 0)), ("weights", (0, 0, 1, 0)), ("activations", (0, 0, 0, 1)),
                  ("planes+features", (1, 1, 0, 0)), ("weights+activations", (0, 0, 1, 1)), ("all (= fp16 mode)", (1, 1, 1, 1)),
                  ("all but weights", (1, 1, 0, 1)), ("all but activations", (1, 1, 1, 0)), ("all but planes+features", (0, 0, 1, 1))]:
    raw, out = run(*cfg)
    ds = float((raw[..., 3] - raw0[..., 3]).abs().max())
    dl = float((raw[..., :3] - raw0[..., :3]).abs().max())
    # exclude last-sample steps (sign change of sigma_last)
    step = (raw[:, -1, 3] > 0) != (raw0[:, -1, 3] > 0)
    e = torch.maximum((out[0] - out0[0]).abs().max(-1)[0], (out[2] - out0[2]).abs())
    print(f"{name:28s} sigma maxdiff {ds:.2e}  logit maxdiff {dl:.2e}  map max err (no-step rays) {float(e[~step].max()):.2e}  mean {float(e[~step].mean()):.2e}")
