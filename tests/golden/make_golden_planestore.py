"""A `.par` scene file written by the REFERENCE's own code: models.create_plane +
models.get_plane_name + nerf_helpers.safe_saving with the content dict of PlanesOptimizer.save_params (models.py:667-668).

    python tests/golden/make_golden_planestore.py            # writes tests/golden/coarse_tiny_DS2_PlRes6_4.par (+ .npz twin)
    python tests/golden/make_golden_planestore.py --load F   # nerf_helpers.safe_loading(F,'par') -> prints a JSON summary
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (shims + reference modules)

SCENE = "tiny_DS2_PlRes6_4"


def summary(content):
    return {"keys": sorted(content.keys()),
            "planes": {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().max())] for k, v in content["params"].items()},
            "box": torch.as_tensor(content["coords_normalization"]).double().tolist(),
            "n_opt_states": len(content["opt_states"])}


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--load":
        _orig = torch.load
        torch.load = lambda *a, **k: _orig(*a, **{**k, "weights_only": False})   # torch 1.12 semantics of the reference
        best = len(sys.argv) > 3 and sys.argv[3] == "best"
        print("PLANESTORE_JSON " + json.dumps(summary(G.nerf_helpers.safe_loading(sys.argv[2], "par", best=best))))
        return
    torch.manual_seed(9)
    params = torch.nn.ParameterDict([(G.models.get_plane_name(SCENE, d), G.models.create_plane(6 if d < 3 else 4, 8, 0.5))
                                     for d in range(4)])
    box = torch.tensor(G.BOX, dtype=torch.float64)
    f = os.path.join(HERE, "coarse_%s.par" % SCENE)
    G.nerf_helpers.safe_saving(f, content={"params": params, "opt_states": [None] * 4, "coords_normalization": box}, suffix="par")
    G.npz("coarse_%s_par_twin.npz" % SCENE, box=box, **{k: v for k, v in params.items()})
    print("wrote", f, os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
