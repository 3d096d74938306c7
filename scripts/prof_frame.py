"""One (or a few) frames of the bench workload with nothing else around it — the target of the
`ncu --set full` captures (a number printed under ncu is never a bench value).

    python scripts/prof_frame.py [--frames 1] [--rows 800] [--precision fp16] [--ray-chunk N]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import nvsr_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--rows", type=int, default=bench.RES, help="render only the first ROWS image rows")
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--ray-chunk", type=int, default=0)
    ap.add_argument("--sparse", action="store_true", help="profile the sparse colour path (default: every sample dense)")
    ap.add_argument("--config", default="cfg2")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    nvsr_b200.set_precision(args.precision)
    if args.ray_chunk:
        nvsr_b200.set_ray_chunk(args.ray_chunk)
    nvsr_b200.set_sparse_rgb(args.sparse)
    w = bench.build_workload(args.config, dev)
    with torch.no_grad():
        for _ in range(args.frames):
            out = w.render(nvsr_b200, 0, min(args.rows, w.H))
    torch.cuda.synchronize()
    print("rgb_fine mean", float(out[3].mean()), "acc_fine mean", float(out[5].mean()))


if __name__ == "__main__":
    main()
