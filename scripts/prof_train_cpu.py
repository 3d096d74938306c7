"""cProfile of the training step's HOST side (which Python calls the step spends its CPU time in)."""
import cProfile, pstats, io, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nvsr_b200
from nvsr_b200 import autograd as A, scene
dev = torch.device("cuda", 0)
mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev)
for m in (mc, mf):
    m.train()
pose, focal = scene.blender_camera(800)
opt, scfg = scene.render_options(64, 128, perturb=True), scene.scene_cfg()
with torch.no_grad():
    ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(dev))
g = torch.Generator().manual_seed(0)
pick = torch.randperm(640000, generator=g)[:4096].to(dev)
batch = torch.stack([ro.reshape(-1, 3)[pick], rd.reshape(-1, 3)[pick]], 0).contiguous()
target = torch.rand(4096, 3, generator=g).to(dev)
rnd = {"t_rand": torch.rand(4096, 64, generator=g).to(dev), "u": torch.rand(4096, 128, generator=g).to(dev)}
params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
def step():
    for p in params:
        p.grad = None
    out = A.run_one_iter_of_nerf(800, 800, focal, mc, mf, batch, opt, sid, "train", scene_config=scfg, randoms=rnd)
    (F.mse_loss(out[0], target) + F.mse_loss(out[3], target)).backward()
for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 20:.2f} ms per step; with final sync {1e3 * (t2 - t0) / 20:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
