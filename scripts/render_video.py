"""The evaluation/video path end to end through the §8f rows either side of the render (round-2 tool, GPU needed):
scenes stored as the reference's `.par` files -> PlaneStore (prefetch of the NEXT scene while the current one renders) ->
orbit poses (load_blender.py:308-311) -> render_frame -> FrameSink (uint8 on the device, side-stream pinned copies,
PNG files named like write_image's, train_nerf.py:268).

    python scripts/render_video.py --out /tmp/video [--scenes 2] [--frames 8] [--res 200]

Writes <out>/planes/coarse_<scene>.par (synthetic scenes, written through the store itself), then
<out>/<scene>/<i>.png, and prints one JSON line: frames, ms per frame with the sink on the critical path vs rendering alone.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nvsr_b200  # noqa: E402
from nvsr_b200 import frames, plane_store, scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--scenes", type=int, default=2)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--res", type=int, default=200)
    ap.add_argument("--plane-res", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    planes_dir = os.path.join(args.out, "planes")
    os.makedirs(planes_dir, exist_ok=True)
    # synthetic scenes, stored in the reference's on-disk format by the store itself
    mc, mf, sid0 = scene.make_synthetic_scene(plane_res=args.plane_res, view_res=32, seed=0, device=dev,
                                              scene_id="s0_DS2_PlRes%d_32" % args.plane_res)
    sids = [sid0] + [scene.add_synthetic_scene(mc, mf, "s%d_DS2_PlRes%d_32" % (i, args.plane_res), plane_res=args.plane_res, seed=100 + i)
                     for i in range(1, args.scenes)]
    store = plane_store.PlaneStore(planes_dir, device=dev, prepack=nvsr_b200.get_precision())
    for sid in sids:
        store.write(sid, {k: mc.planes_[k] for k in plane_store.plane_names(sid)}, mc.box_coords[sid])
    opt, scfg = scene.render_options(64, 128), scene.scene_cfg()
    focal = scene.blender_camera(args.res)[1]
    poses = torch.from_numpy(frames.orbit_poses(args.frames)).to(dev)

    def video(with_sink):
        store.evict(sids[0])
        store.prefetch(sids[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():
            for i, sid in enumerate(sids):
                store.attach([mc, mf], sid)                       # render stream waits on the copy event only
                if i + 1 < len(sids):
                    store.evict(sids[i + 1]), store.prefetch(sids[i + 1])   # next scene loads while this one renders
                sink = frames.FrameSink(frames.png_writer(os.path.join(args.out, sid))) if with_sink else None
                for f in range(args.frames):
                    out = nvsr_b200.render_frame(args.res, args.res, focal, poses[f], mc, mf, opt, sid, scfg)
                    if sink is not None:
                        sink.submit(out[3].reshape(args.res, args.res, 3))
                if sink is not None:
                    sink.flush()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (len(sids) * args.frames)

    video(False)                                                  # warm-up: packs planes, builds caches
    ms_render = video(False)
    ms_sink = video(True)
    print(json.dumps({"scenes": len(sids), "frames_per_scene": args.frames, "res": args.res,
                      "ms_per_frame_render_only": ms_render, "ms_per_frame_with_store_and_png_sink": ms_sink,
                      "png_files": sum(len(os.listdir(os.path.join(args.out, s))) for s in sids)}))


if __name__ == "__main__":
    main()
