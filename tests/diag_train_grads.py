"""Diagnostic (not a test): per-parameter error table of the train-step gradient test on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import helpers as H
from nvsr_b200 import autograd as A
from oracle import nvsr_oracle as O
import test_gpu_next_rows as T

DEV = "cuda:0"
A.set_decoder("fp32")
g, sid, opt, scfg, batch = T._train_case()
target = H.T(g["target"], DEV)
rnd = H.randoms_from(g, DEV)
mc_o, mf_o = H.load_planes_scene(str(g["scene_file"]), sid, "cpu")
tc = {}
with torch.no_grad():
    O.run_one_iter_of_nerf(int(g["H"]), int(g["W"]), float(g["focal"]), mc_o, mf_o, batch, opt, sid, "train",
                           scene_config=scfg, randoms=H.randoms_from(g), trace=tc)
for forced in (True, False):
    mc, mf = H.load_planes_scene(str(g["scene_file"]), sid, DEV)
    named = T._named_params(mc, mf)
    r = dict(rnd, z_fine=tc["z_fine"].to(DEV)) if forced else dict(rnd)
    out = A.run_one_iter_of_nerf(int(g["H"]), int(g["W"]), float(g["focal"]), mc, mf, batch.to(DEV), opt, sid, "train",
                                 scene_config=scfg, randoms=r)
    loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
    loss.backward()
    print("forced" if forced else "free", "loss", float(loss), "golden", float(g["loss"]),
          "rgb_c err", float((out[0].cpu() - torch.from_numpy(g["rgb_coarse"])).abs().max()),
          "rgb_f err", float((out[3].cpu() - torch.from_numpy(g["rgb_fine"])).abs().max()))
    for k in [k[len("grad__"):] for k in g if k.startswith("grad__")]:
        want = torch.from_numpy(g["grad__" + k]); got = named[k].grad.cpu()
        d = (got - want).abs(); i = int(d.argmax())
        print(f"  {k:45s} scale {float(want.abs().max()):.3e} err {float(d.max()):.3e} rel {float(d.max())/float(want.abs().max()):.2e} "
              f"at {i}: got {float(got.flatten()[i]):.4e} want {float(want.flatten()[i]):.4e}  n(err>1e-3 scale) {int((d > 1e-3*want.abs().max()).sum())}")
