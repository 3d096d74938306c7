"""BASELINE.json config 5: several scenes' tri-planes sharing ONE decoder pair, an orbit video per scene.

Single process: the scenes are rendered round-robin, one frame each, so every frame switches scene (packed planes
are cached per scene: after the first frame of a scene a switch costs nothing).  Under torchrun (one process per
GPU): scene-parallel, rank r renders the scenes r, r + world, ... — no communication per frame, one barrier at the
end; the aggregate frame rate is printed by rank 0.

    python scripts/run_multiscene.py [--scenes 8] [--frames 8] [--res 800]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_multiscene.py
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nvsr_b200  # noqa: E402
from nvsr_b200 import scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--frames", type=int, default=8, help="orbit poses per scene (the full config has 200)")
    ap.add_argument("--res", type=int, default=800)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    mc, mf, sid0 = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev, scene_id="s0_DS2_PlRes200_32")
    sids = [sid0] + [scene.add_synthetic_scene(mc, mf, "s%d_DS2_PlRes200_32" % i, seed=100 + i) for i in range(1, args.scenes)]
    mine = sids[rank::world]
    opt, scfg = scene.render_options(64, 128), scene.scene_cfg()
    thetas = torch.linspace(-180.0, 180.0, args.frames + 1)[:-1].tolist()
    poses = [scene.blender_camera(args.res, theta=t)[0].to(dev) for t in thetas]
    focal = scene.blender_camera(args.res)[1]
    acc = {}
    with torch.no_grad():
        for sid in mine:                       # first frame of every scene: packs its planes (untimed warm-up)
            nvsr_b200.render_frame(args.res, args.res, focal, poses[0], mc, mf, opt, sid, scfg)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in range(args.frames):
            for sid in mine:                   # round-robin: a scene switch before every frame
                out = nvsr_b200.render_frame(args.res, args.res, focal, poses[f], mc, mf, opt, sid, scfg)
                acc[sid] = out[5].mean()
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    n_frames = args.frames * args.scenes
    if rank == 0:
        print(json.dumps({"config": "cfg5_%d_scenes_one_decoder_%dx%d_64+128" % (args.scenes, args.res, args.res),
                          "n_gpus": world, "frames": n_frames, "ms_total": float(ms), "ms_per_frame_per_gpu": float(ms) / (args.frames * len(mine)),
                          "frames_per_s": n_frames / float(ms) * 1e3, "rays_per_s": n_frames * args.res * args.res / float(ms) * 1e3,
                          "acc_fine_mean_of_rank0_scenes": {k: round(float(v), 4) for k, v in acc.items()}}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
