"""Diagnostic (not a test): tc-vs-fp32 gradient agreement of the train step as a function of the loss scale."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nvsr_b200
from nvsr_b200 import autograd as A, scene
import test_gpu_train_tc as T

DEV = "cuda:0"
mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=DEV)
for m in (mc, mf):
    m.train()
Hh = Ww = 32
pose, focal = scene.blender_camera(Hh)
opt, scfg = scene.render_options(64, 128, perturb=True, white_background=True, noise_std=0.2), scene.scene_cfg()
with torch.no_grad():
    ro, rd = nvsr_b200.get_ray_bundle(Hh, Ww, focal, pose.to(DEV))
batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
n = batch.shape[1]
g = torch.Generator().manual_seed(3)
rnd = dict(t_rand=torch.rand(n, 64, generator=g), u=torch.rand(n, 128, generator=g),
           noise_c=torch.randn(n, 64, generator=g), noise_f=torch.randn(n, 192, generator=g))
target = torch.rand(n, 3, generator=g).to(DEV)
tr = {}
l32, o32, g32 = T._step("fp32", mc, mf, sid, batch, opt, scfg, dict(rnd, trace=tr), target, Hh, Ww, focal)
for sc in (2.0 ** 6, 2.0 ** 10, 2.0 ** 14, 2.0 ** 18, 2.0 ** 22):
    A.set_loss_scale(sc)
    ltc, otc, gtc = T._step("tc", mc, mf, sid, batch, opt, scfg, dict(rnd, z_fine=tr["z_fine"]), target, Hh, Ww, focal)
    rows = []
    for k in g32:
        a, b = gtc[k].double().flatten(), g32[k].double().flatten()
        rows.append((float((a - b).norm() / (b.norm() + 1e-30)), k))
    rows.sort(reverse=True)
    print(f"scale 2^{int(torch.log2(torch.tensor(sc)))}: loss {ltc:.6f} vs {l32:.6f};", " ".join(f"{k.split('.')[-3] if 'planes' not in k else 'plane'}.{k.split('.')[-1][-9:]}={v:.3e}" for v, k in rows[:6]),
          "| finite:", all(bool(torch.isfinite(v).all()) for v in gtc.values()))
# forward-only effect: how far are the fp16-path maps from the fp32 path's
print("rgb_c maxdiff", float((otc[0] - o32[0]).abs().max()), "rgb_f maxdiff", float((otc[3] - o32[3]).abs().max()))
