"""N > 1 host logic on CPU: world_size-2 `gloo` process group, each rank renders its row band (with the
oracle standing in for the CUDA render, which needs a GPU), one all_gather per frame, and the assembled
frame must equal the single-process full-frame render bit for bit (ray order preserved)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nvsr_b200
from nvsr_b200 import scene, sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(res):
    mc, mf, sid = scene.make_synthetic_scene(plane_res=16, view_res=8, seed=3)
    pose, focal = scene.blender_camera(res)
    return mc, mf, sid, pose, focal, scene.render_options(8, 8), scene.scene_cfg()


def _band_renderer(res):
    from oracle import nvsr_oracle as O
    mc, mf, sid, pose, focal, opt, scfg = _scene(res)
    ro, rd = O.get_ray_bundle(res, res, focal, pose)

    def render_band(r0, r1):
        batch = torch.stack([ro[r0:r1].reshape(-1, 3), rd[r0:r1].reshape(-1, 3)], 0)
        with torch.no_grad():
            return O.run_one_iter_of_nerf(res, res, focal, mc, mf, batch, opt, sid, "validation", scene_config=scfg)
    return render_band


def _worker(rank, world, port, res, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = sharding.FrameSharder(res, res, rank, world, "cpu")
        buf = sh.render(_band_renderer(res))
        dist.barrier()
        if rank == 0:
            torch.save({k: v.clone() for k, v in sh.frame(buf).items()}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,res", [(2, 12), (2, 11)])   # even and ragged band split
def test_two_rank_row_bands_equal_full_frame(tmp_path, world, res):
    out_path = str(tmp_path / "frame.pt")
    mp.spawn(_worker, args=(world, _free_port(), res, out_path), nprocs=world, join=True)
    got = torch.load(out_path)
    full = _band_renderer(res)(0, res)
    names = ("rgb_coarse", "disp_coarse", "acc_coarse", "rgb_fine", "disp_fine", "acc_fine")
    for k, ref in zip(names, full[:6]):
        ref = ref.reshape(res, res, -1).squeeze(-1) if ref.dim() == 1 else ref.reshape(res, res, -1)
        assert torch.equal(torch.nan_to_num(got[k], 7.0), torch.nan_to_num(ref, 7.0)), k


def test_row_band_partition_properties():
    for h in (1, 7, 100, 756, 800):
        for w in (1, 2, 3, 4, 8, 16):
            bands = [sharding.row_band(h, r, w) for r in range(w)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(b[1] == c[0] for b, c in zip(bands, bands[1:]))           # contiguous, ordered
            assert max(b[1] - b[0] for b in bands) == sharding.rows_per_rank(h, w)
