"""nvsr_b200.autograd against the LIVE reference's own model classes and training step (only where the reference checkout
exists; run as a subprocess by tests/test_oracle_live_reference.py because importing the reference needs the shims of
SURVEY.md §8c).

The reference's TwoDimPlanesModel / FlexibleNeRFModel objects are rendered in train mode (a) by the reference's
train_utils.run_one_iter_of_nerf + loss.backward() (train_nerf.py:860-905) and (b) by nvsr_b200.autograd's composition
with the C-ABI calls replaced by the host stand-ins of tests/host_ops.py (the kernels' own bodies, built for the host).
Same random draws (recorded from the reference's CPU generator).  Prints {check: max relative gradient difference}.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402  (applies the shims and imports the reference modules)
import host_ops as HO  # noqa: E402
import nvsr_b200  # noqa: E402
from nvsr_b200 import autograd as A, ops  # noqa: E402

for name, fn in HO.standins(HO.build_hostcheck()).items():
    setattr(ops, name, fn)
A.set_decoder("fp32")     # the fp32 parity mode of the differentiable path (the tcgen05 decoder has no host stand-in)
out = {}


def grads_of(params, fn, target):
    for p in params:
        p.grad = None
    o = fn()
    (torch.nn.functional.mse_loss(o[0], target) + torch.nn.functional.mse_loss(o[3], target)).backward()
    return o, [None if p.grad is None else p.grad.clone() for p in params]


def compare(tag, got, want, o_p, o_r):
    worst = 0.0
    n = 0
    for a, b in zip(got, want):
        assert (a is None) == (b is None), tag
        if b is not None and float(b.abs().max()) > 0:
            worst = max(worst, float((a - b).abs().max() / b.abs().max()))
            n += 1
    assert n >= 16, (tag, n)
    out[tag + "_grad_rel"] = worst
    out[tag + "_rgb_abs"] = max(float((o_p[j].detach() - o_r[j].detach()).abs().max()) for j in (0, 3))


# ---- tri-plane model: the reference's TwoDimPlanesModel pair, perturbation + noise + white background
sid = "live_DS2_PlRes12_6"
mc, mf, _ = G.build_planes_models(sid, 12, 6, seed=3)
opt = G.options(12, 10, perturb=True, white=True, noise=0.4)
scfg = G.CfgNode(dict(near=2.0, far=6.0, no_ndc=True))
pose = torch.from_numpy(G.pose_spherical(55.0, -30.0, 4.0)).float()
ro, rd = G.nerf_helpers.get_ray_bundle(9, 9, 11.0, pose)
batch = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
target = torch.rand(81, 3, generator=torch.Generator().manual_seed(1))
params = list({id(p): p for m in (mc, mf) for p in m.parameters() if p.requires_grad}.values())
torch.manual_seed(11)
with G.Recorder() as rec:
    o_r, want = grads_of(params, lambda: G.train_utils.run_one_iter_of_nerf(9, 9, 11.0, mc, mf, batch, opt, sid, mode="train",
                                                                          scene_config=scfg), target)
rnd = {"t_rand": rec.rand[0], "u": rec.rand[1], "noise_c": rec.randn[0], "noise_f": rec.randn[1]}
o_p, got = grads_of(params, lambda: A._run_one_iter(9, 9, 11.0, mc, mf, batch, opt, sid, "train", scfg, rnd), target)
compare("planes", got, want, o_p, o_r)

# ---- mip / IPE family: the reference's FlexibleNeRFModel pair with its IntegratedPositionalEncoding
torch.manual_seed(4)
kw = dict(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True, use_viewdirs=True)
fc, ff = G.models.FlexibleNeRFModel(**kw), G.models.FlexibleNeRFModel(**kw)
fc.optional_no_grad = G.nerf_helpers.null_with     # train_nerf.py:349
enc = G.mip.IntegratedPositionalEncoding(3, 7)
encd = lambda x: G.nerf_helpers.positional_encoding(x, 4, True)
opt_m = G.options(10, 8, perturb=True, white=False, noise=0.2, mip_enc=True)
params_m = [p for m in (fc, ff) for p in m.parameters()]
torch.manual_seed(12)
with G.Recorder() as rec:
    o_r, want = grads_of(params_m, lambda: G.train_utils.run_one_iter_of_nerf(9, 9, 11.0, fc, ff, batch, opt_m, "lego_DS2", mode="train",
                                                                            encode_position_fn=enc, encode_direction_fn=encd,
                                                                            scene_config=scfg), target)
rnd = {"t_rand": rec.rand[0], "u": rec.rand[1], "noise_c": rec.randn[0], "noise_f": rec.randn[1]}
o_p, got = grads_of(params_m, lambda: A._run_one_iter(9, 9, 11.0, fc, ff, batch, opt_m, "lego_DS2", "train", scfg, rnd,
                                                      nvsr_b200.IntegratedPositionalEncoding(3, 7)), target)
compare("mip", got, want, o_p, o_r)

print("LIVE_BACKWARD_JSON " + json.dumps(out))
