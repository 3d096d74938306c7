"""Multi-GPU frame rendering: rays are independent, so a frame is cut into contiguous row bands, one per
rank (one process per GPU), and the only exchange is ONE all_gather of the finished image tiles per frame
(NCCL over NVLink on the B200 box; gloo in the CPU tests).  SURVEY.md §8(e).

The reference has no multi-device path (single process, single GPU — train_nerf.py:284-287); ray order
(`get_ray_bundle`'s row-major [H,W], nerf_helpers.py:507-549) is preserved by construction: band r holds
image rows [r*rows_per, (r+1)*rows_per) and bands are concatenated in rank order.
"""
import torch

N_MAPS = 10  # rgb_coarse(3) disp_coarse acc_coarse rgb_fine(3) disp_fine acc_fine


def rows_per_rank(height, world):
    return (height + world - 1) // world


def row_band(height, rank, world):
    """image rows [r0, r1) rendered by `rank` (possibly empty for trailing ranks when world > height)"""
    per = rows_per_rank(height, world)
    return min(height, rank * per), min(height, (rank + 1) * per)


def pack_tile(out, n_local, tile):
    """9-tuple of run_one_iter_of_nerf -> rows [0, n_local) of the [rows_per*W, 10] send tile (padding rows
    keep their previous contents; they are dropped by `unpack_frame`)."""
    tile[:n_local, 0:3], tile[:n_local, 3], tile[:n_local, 4] = out[0], out[1], out[2]
    if out[3] is not None:
        tile[:n_local, 5:8], tile[:n_local, 8], tile[:n_local, 9] = out[3], out[4], out[5]
    return tile


def gather_tiles(tile, gathered, group=None):
    """the one collective of a frame: equal-sized padded tiles, concatenated in rank order"""
    import torch.distributed as dist
    dist.all_gather_into_tensor(gathered, tile, group=group)
    return gathered


def unpack_frame(gathered, height, width, world):
    """[world*rows_per*W, 10] -> dict of [H,W,...] maps in the reference's ray order"""
    per = rows_per_rank(height, world)
    g = gathered.reshape(world * per, width, N_MAPS)[:height]
    return {"rgb_coarse": g[..., 0:3], "disp_coarse": g[..., 3], "acc_coarse": g[..., 4],
            "rgb_fine": g[..., 5:8], "disp_fine": g[..., 8], "acc_fine": g[..., 9]}


class FrameSharder:
    """Per-rank state of a sharded render: band, send tile, gather buffer."""

    def __init__(self, height, width, rank, world, device, group=None):
        self.height, self.width, self.rank, self.world, self.group = height, width, rank, world, group
        self.r0, self.r1 = row_band(height, rank, world)
        self.per = rows_per_rank(height, world)
        self.n_local = (self.r1 - self.r0) * width
        self.tile = torch.zeros((self.per * width, N_MAPS), dtype=torch.float32, device=device)
        self.gathered = (torch.zeros((world * self.per * width, N_MAPS), dtype=torch.float32, device=device)
                         if world > 1 else None)

    def render(self, render_band):
        """render_band(r0, r1) -> 9-tuple for image rows [r0, r1); returns the gathered [world*per*W, 10]
        buffer (every rank holds the whole frame) — or this rank's tile when world == 1."""
        if self.n_local > 0:
            pack_tile(render_band(self.r0, self.r1), self.n_local, self.tile)
        if self.world > 1:
            return gather_tiles(self.tile, self.gathered, self.group)
        return self.tile

    def frame(self, buf):
        return unpack_frame(buf, self.height, self.width, self.world)
