"""Diagnostic (not collected): order dependence of test_composite_bwd_matches_oracle_autograd[False-0.0-False]."""
import os, sys
import pytest, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
here = os.path.dirname(os.path.abspath(__file__))
rc = pytest.main([os.path.join(here, "test_gpu_train_tc.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"])
print("train_tc rc", rc)
import nvsr_b200
from nvsr_b200 import ops, autograd as A
sys.path.insert(0, os.path.dirname(here))
from oracle import nvsr_oracle as O
DEV = "cuda:0"
g = torch.Generator().manual_seed(21)
n, S = 1000, 96
raw = torch.randn(n, S, 4, generator=g) * 1.5
raw[..., 3] = raw[..., 3] * 4.0 - 1.0
raw[7, 10:20, 3] = 60.0
z = torch.sort(2.0 + 4.0 * torch.rand(n, S, generator=g), -1).values
rd = torch.randn(n, 3, generator=g)
g_rgb, g_acc, g_depth, g_w = (torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g))
outs = []
for rep in range(3):
    raw_o = raw.clone().requires_grad_(True)
    rgb, _, acc, w, depth = O.volume_render_radiance_field(raw_o, z, rd, 0.0, False, mip_nerf=False, noise=None)
    ((rgb * g_rgb).sum() + (acc * g_acc).sum() + (depth * g_depth).sum() + (w * g_w).sum()).backward()
    d_raw = ops.composite_bwd(raw.to(DEV), z.to(DEV), rd.to(DEV), g_rgb.to(DEV), g_acc.to(DEV), g_depth.to(DEV), g_w.to(DEV),
                              noise=None, white_background=False, mip=False).cpu()
    raw_g = raw.to(DEV).requires_grad_(True)
    out = A.volume_render_radiance_field(raw_g, z.to(DEV), rd.to(DEV), 0.0, False, mip_nerf=False, noise=None)
    ((out[0] * g_rgb.to(DEV)).sum() + (out[2] * g_acc.to(DEV)).sum() + (out[4] * g_depth.to(DEV)).sum() + (out[3] * g_w.to(DEV)).sum()).backward()
    ga = raw_g.grad.cpu()
    ref = raw_o.grad
    for name, t in (("stage", d_raw), ("autograd", ga)):
        e = (t[..., :3] - ref[..., :3]).abs()
        i = int(e.argmax())
        r_, s_, c_ = i // (S * 3), (i // 3) % S, i % 3
        print(rep, name, "max err rgb", float(e.max()), "at", (r_, s_, c_), "got", float(t[r_, s_, c_]), "ref", float(ref[r_, s_, c_]),
              "scale", float(ref[..., :3].abs().max()), "threads", torch.get_num_threads())
    outs.append((d_raw, ga, ref.clone()))
print("stage deterministic", all(torch.equal(outs[0][0], o[0]) for o in outs), "autograd deterministic", all(torch.equal(outs[0][1], o[1]) for o in outs),
      "oracle deterministic", all(torch.equal(outs[0][2], o[2]) for o in outs))
print("stage == autograd", torch.equal(outs[0][0], outs[0][1]), float((outs[0][0] - outs[0][1]).abs().max()))
