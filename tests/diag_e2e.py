"""Stage-by-stage diff of the CUDA path vs the CPU oracle on a golden case (diagnostics; lives under tests/ because it
runs the oracle — only tests/, smoke() and bench.py's CPU baseline may)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import nvsr_b200
from test_oracle_golden import run_oracle_e2e, NAMES

names = sys.argv[2:] or ["e2e_planes_det.npz", "e2e_planes_sr.npz", "e2e_planes_perturb.npz"]
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
nvsr_b200.set_precision(prec)
for name in names:
    tg, tc = {}, {}
    g, out_g = run_oracle_e2e(name, "cuda:0", runner=nvsr_b200.run_one_iter_of_nerf, trace=tg)
    _, out_c = run_oracle_e2e(name, "cpu", trace=tc)
    print("=====", name, prec)
    for k in tc:
        if k not in tg:
            continue
        a, b = tg[k].cpu(), tc[k]
        if a.dtype == torch.int64:
            print(f"  {k:16s} mismatches {int((a != b).sum())}/{a.numel()}")
            continue
        d = (a - b).abs()
        d = d[~torch.isnan(d)]
        print(f"  {k:16s} max {float(d.max()):.3e}  mean {float(d.mean()):.3e}  #>1e-4 {int((d > 1e-4).sum())}/{d.numel()}  ref absmax {float(b[~torch.isnan(b)].abs().max()):.3g}")
    for k, a, b in zip(NAMES, out_g[:6], out_c[:6]):
        if a is None:
            continue
        d = (a.cpu() - b).abs()
        d = d[~torch.isnan(d)]
        gd = (b - torch.from_numpy(g[k])).abs()
        gd = gd[~torch.isnan(gd)]
        print(f"  OUT {k:12s} gpu-vs-oracle max {float(d.max()):.3e} mean {float(d.mean()):.3e} | oracle-vs-golden max {float(gd.max()):.3e}")
    if "raw_coarse" in tg:
        # where is the worst coarse ray?
        d = (out_g[0].cpu() - out_c[0]).abs().max(-1)[0]
        r = int(d.argmax())
        print("  worst coarse ray", r, float(d[r]), "acc", float(out_g[2][r]), float(out_c[2][r]))
        print("   raw sigma gpu ", [round(float(x), 4) for x in tg["raw_coarse"][r, :, 3].cpu()])
        print("   raw sigma cpu ", [round(float(x), 4) for x in tc["raw_coarse"][r, :, 3]])
