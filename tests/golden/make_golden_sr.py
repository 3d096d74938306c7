"""Golden vectors of the plane super-resolution step made by RUNNING THE REFERENCE's own
`PlanesSR` + `EDSR` (models.py:773-926) on seeded LR planes — CPU, this container only (same shims as make_golden.py).

    python tests/golden/make_golden_sr.py        # writes tests/golden/sr_*.npz

Each fixture holds: the EDSR state dict, the constructor arguments, the LR planes, and the SR planes
`PlanesSR.forward(plane_name)` returned (the tensor it caches in `SR_planes`, models.py:925)."""
import os
import sys
import types

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

REF = os.environ.get("NVSR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
scipy.signal.gaussian = scipy.signal.windows.gaussian
sys.modules.setdefault("imageio", types.ModuleType("imageio"))
if not torch.cuda.is_available():
    torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)
import models  # noqa: E402
from cfgnode import CfgNode  # noqa: E402


def case(name, seed, channels, res, hidden, n_blocks, scale, n_planes=3, weight_gain=10.0, normalize=False):
    torch.manual_seed(seed)
    sr = models.PlanesSR(model_arch=models.EDSR, scale_factor=scale, in_channels=channels, out_channels=channels,
                         sr_config=CfgNode({"model": {"hidden_size": hidden, "n_blocks": n_blocks},
                                            "input_normalization": normalize}), plane_interp="bilinear").eval()
    sr.align_corners = True                      # TwoDimPlanesModel.assign_SR_model sets it (models.py:253)
    with torch.no_grad():
        # the reference initialises every conv with std sqrt(2/n)/10 (models.py:843-846): the SR residual of a fresh model is
        # ~1e-6 of the plane's scale and would hide any error of the conv chain — scale the weights up so that the
        # residual is a visible fraction of the output
        for m in sr.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.mul_(weight_gain)
        if normalize:
            sr.normalization_params({"mean": torch.randn(channels) * 0.2, "std": torch.rand(channels) * 0.5 + 0.75})
    out = {"scale": scale, "hidden": hidden, "n_blocks": n_blocks, "channels": channels, "res": res,
           "required_padding": sr.inner_model.required_padding, "hr_overpadding": sr.HR_overpadding,
           "normalize": int(normalize)}
    for k, v in sr.state_dict().items():
        out["w__" + k.replace(".", "__")] = v.detach().numpy()
    with torch.no_grad():
        for d in range(n_planes):
            lr = torch.randn(1, channels, res, res + (d if name.endswith("ragged") else 0)) * 0.5
            pname = "sc%s_D%d" % (name, d)
            sr.set_LR_plane(lr, pname, save_interpolated=False)
            hr = sr(pname)
            assert pname in sr.SR_planes and not torch.isnan(hr).any()
            out["lr__" + pname] = lr.numpy()
            out["sr__" + pname] = hr.numpy()
            res_only = hr - sr.interpolate_LR(pname)
            print(name, pname, tuple(lr.shape), "->", tuple(hr.shape), "| residual rms %.3e of output rms %.3e"
                  % (float(res_only.pow(2).mean().sqrt()), float(hr.pow(2).mean().sqrt())))
    np.savez_compressed(os.path.join(OUT, "sr_%s.npz" % name), **out)


if __name__ == "__main__":
    case("small", 3, channels=48, res=8, hidden=16, n_blocks=2, scale=4)
    case("x2_norm", 4, channels=16, res=12, hidden=32, n_blocks=3, scale=2, n_planes=2, normalize=True)
    case("ragged", 5, channels=8, res=9, hidden=24, n_blocks=1, scale=4, n_planes=2)
