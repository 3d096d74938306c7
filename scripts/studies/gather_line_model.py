"""CPU study: how many distinct 128-byte lines / 32-byte sectors the gather's texel loads touch per warp-level load
instruction on the bench scene (coarse and fine pass), for the current plane layout + lane mapping and for candidates.
Calibrated against the ncu wavefront counts of profiles/r2_step1_ncu_full.txt (fine / coarse = 19.1 / 11.6 per LDG.256).

    python scripts/studies/gather_line_model.py
"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import nvsr_oracle as O

torch.set_num_threads(os.cpu_count() or 8)
w = bench.build_workload("cfg2", "cpu")
H = W = 800
ro, rd = O.get_ray_bundle(H, W, w.focal, w.pose)
# a patch of 32 image rows x 64 columns around the object's silhouette: rays adjacent along x (8-ray blocks)
y0, x0, ph, pw = 380, 330, 16, 64
ro_p, rd_p = ro[y0:y0 + ph, x0:x0 + pw].reshape(-1, 3), rd[y0:y0 + ph, x0:x0 + pw].reshape(-1, 3)
batch = torch.stack([ro_p, rd_p], 0)
tr = {}
with torch.no_grad():
    O.run_one_iter_of_nerf(H, W, w.focal, w.mc, w.mf, batch, w.opt, w.scene_id(False), "validation", scene_config=w.scfg, trace=tr)
zc, zf = tr["z_coarse"].numpy(), tr["z_fine"].numpy()
n = ro_p.shape[0]
model = w.mf
model.set_cur_scene_id(w.scene_id(False))
box = model.box_coords[model.cur_id].double().numpy()
rots = [r.detach().numpy() for r in model.coord_projector.rot_mats_NON_LEARNED]
R = 200


def cells(z):
    """[n, S, 3 planes, 2] integer (x0, y0) of the bilinear footprint"""
    pts = ro_p.numpy()[:, None, :] + rd_p.numpy()[:, None, :] * z[..., None]
    cn = 2 * (pts - box[0, :3]) / (box[1, :3] - box[0, :3]) - 1
    out = np.zeros(z.shape + (3, 2), dtype=np.int64)
    for d in range(3):
        g = cn @ rots[d][:, 1:]
        ix = np.clip((g[..., 0] + 1) / 2 * (R - 1), 0, R - 1)
        iy = np.clip((g[..., 1] + 1) / 2 * (R - 1), 0, R - 1)
        out[..., d, 0] = np.minimum(np.floor(ix), R - 1)
        out[..., d, 1] = np.minimum(np.floor(iy), R - 1)
    return out


def addr_chunk_major(y, c, x):       # current: [Rh][C/8][Rw] records of 32 B
    return ((y * 6 + c) * R + x) * 32


def addr_texel_major(y, c, x):       # candidate: [Rh][Rw][C/8] records of 32 B
    return ((y * R + x) * 6 + c) * 32


def stats(z, mapping, layout, name):
    S = z.shape[1]
    c = cells(z)
    ts = -(-S // 16)
    lines_w, lines_q, sect_w = [], [], []
    for blk in range(n // 8):
        for t in range(ts):
            # rows of the tile: r -> (sample, ray)
            r = np.arange(128)
            if mapping == "rays8":            # current BLOCKED order: row = (s % 16) * 8 + ray % 8
                s_of, ray_of = t * 16 + (r >> 3), blk * 8 + (r & 7)
            elif mapping == "r2s4":           # quarter = 2 rays x 4 consecutive samples
                s_of, ray_of = t * 16 + ((r >> 5) << 2) + (r & 3), blk * 8 + ((r >> 2) & 7)
            elif mapping == "r4s2":
                s_of, ray_of = t * 16 + ((r >> 4) << 1) + (r & 1), blk * 8 + ((r >> 1) & 7)
            else:                             # "s8": quarter = 8 consecutive samples of one ray
                s_of, ray_of = t * 16 + ((r >> 6) << 3) + (r & 7), blk * 8 + ((r >> 3) & 7)
            ok = s_of < S
            s_c = np.minimum(s_of, S - 1)
            cc = c[ray_of, s_c]              # [128, 3, 2]
            for d in range(3):
                for dy in (0, 1):
                    y = np.minimum(cc[:, d, 1] + dy, R - 1)
                    x = cc[:, d, 0]
                    if layout == "coop":      # texel-major, lanes = (row, chunk): a warp = 5.33 rows x 6 chunks
                        a_all = np.concatenate([addr_texel_major(y, ch, x)[:, None] for ch in range(6)], 1).reshape(-1)
                        for w0 in range(0, 768, 32):
                            a = a_all[w0:w0 + 32]
                            lines_w.append(len(np.unique(a // 128)))
                            sect_w.append(len(np.unique(a // 32)))
                            lines_q.append(sum(len(np.unique(a[q:q + 8] // 128)) for q in range(0, 32, 8)))
                        continue
                    a = (addr_chunk_major if layout == "chunk" else addr_texel_major)(y, 0, x)   # any chunk: same pattern
                    for w0 in range(0, 128, 32):
                        aw = a[w0:w0 + 32][ok[w0:w0 + 32]]
                        if aw.size == 0:
                            continue
                        lines_w.append(len(np.unique(aw // 128)))
                        sect_w.append(len(np.unique(aw // 32)))
                        lines_q.append(sum(len(np.unique(a[w0 + q:w0 + q + 8] // 128)) for q in range(0, 32, 8)))
    per_row = {"coop": 6 / 5.333}.get(layout, 36 / 32)   # load instructions per row
    print(f"{name:44s} lines/warp-instr {np.mean(lines_w):5.2f}  sum of per-quarter lines {np.mean(lines_q):5.2f}  "
          f"sectors/warp-instr {np.mean(sect_w):5.2f}  | per row: lines {np.mean(lines_w) * per_row:5.2f}, quarter-lines {np.mean(lines_q) * per_row:5.2f}")


for nm, z in (("coarse", zc), ("fine", zf)):
    stats(z, "rays8", "chunk", f"{nm}: current (8 rays / quarter, chunk-major)")
    stats(z, "r2s4", "chunk", f"{nm}: 2 rays x 4 samples / quarter")
    stats(z, "r4s2", "chunk", f"{nm}: 4 rays x 2 samples / quarter")
    stats(z, "s8", "chunk", f"{nm}: 8 samples of one ray / quarter")
    stats(z, "rays8", "texel", f"{nm}: 8 rays / quarter, texel-major (1 chunk)")
    stats(z, "rays8", "coop", f"{nm}: texel-major, lanes = (row, chunk)")
