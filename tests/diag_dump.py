"""Diagnostics: dump the CUDA path's per-stage traces (every precision mode) for a set of seeded cases into
gpurun_out/dump_<case>_<prec>.npz, so that the attribution logic of the parity tests (tests/parity_attribution.py) can be
developed against real device outputs on a box without a GPU.  The scenes are seeded and built on the CPU first, so
the same models can be re-created offline for the oracle.  Lives under tests/ (diagnostics of the parity tests).

    python tests/diag_dump.py [case ...]        # cases: big cfg1 cfg4 mipbig goldens
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch

import nvsr_b200
from nvsr_b200 import scene
from parity_cases import CASES, build_case   # shared with the offline analysis and the GPU tests

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
NAMES = ["rgb_coarse", "disp_coarse", "acc_coarse", "rgb_fine", "disp_fine", "acc_fine"]


def dump(tag, prec, out, tr):
    d = {}
    for k, v in zip(NAMES, out[:6]):
        if v is not None:
            d[k] = v.detach().float().cpu().numpy()
    for k, v in tr.items():
        v = v.detach().cpu()[:1536]        # traces of the first 1536 rays (gpurun_out/ comes back only below 64 MiB)
        d["tr_" + k] = v.numpy().astype(np.int16) if v.dtype == torch.int64 else v.float().numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, f"dump_{tag}_{prec}.npz"), **d)
    print("dumped", tag, prec, {k: v.shape for k, v in d.items()}, flush=True)


def main():
    cases = sys.argv[1:] or list(CASES) + ["goldens"]
    dev = "cuda:0"
    for name in cases:
        if name == "goldens":
            from test_oracle_golden import E2E, run_oracle_e2e
            for gname in E2E:
                for prec in ("fp32", "fp16", "bf16"):
                    nvsr_b200.set_precision(prec)
                    tr = {}
                    _, out = run_oracle_e2e(gname, dev, runner=nvsr_b200.run_one_iter_of_nerf, trace=tr)
                    dump("golden_" + gname.replace(".npz", ""), prec, out, tr)
            continue
        c = build_case(name, dev)
        for prec in CASES[name]["precisions"]:
            nvsr_b200.set_precision(prec)
            nvsr_b200.set_sparse_rgb(False)
            tr = {}
            with torch.no_grad():
                out = nvsr_b200.run_one_iter_of_nerf(c["H"], c["W"], c["focal"], c["mc"], c["mf"], c["batch"], c["opt"], c["sid"],
                                                     "validation", encode_position_fn=c["enc"], encode_direction_fn=c["encd"],
                                                     scene_config=c["scfg"], trace=tr)
            torch.cuda.synchronize()
            dump(name, prec, out, tr)
        del c
        scene.clear_caches()
        torch.cuda.empty_cache()
    nvsr_b200.set_sparse_rgb(True)
    nvsr_b200.set_precision("fp16")


if __name__ == "__main__":
    main()
