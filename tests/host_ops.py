"""Host stand-ins for the C-ABI calls of nvsr_b200.ops (TEST INFRASTRUCTURE): forward = torch ops of the oracle,
backward = the host build of the kernels' own per-element bodies (tests/hostcheck/hostcheck.cpp).  With them
nvsr_b200.autograd — the autograd.Functions, the channels-last <-> NCHW gradient views, None gradients, the decoder
under torch autograd — runs on the CPU as the product code it is.  Used by tests/test_backward_bodies.py (pytest
fixture) and tests/golden/check_live_backward.py (against the live reference's model classes)."""
import ctypes as C
import os
import subprocess
import tempfile

import torch
import torch.nn.functional as F

from oracle import nvsr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "neural-volume-super-resolution_b200", "csrc")


def build_hostcheck(out_dir=None):
    out = os.path.join(out_dir or tempfile.mkdtemp(prefix="hostcheck"), "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", out,
                    os.path.join(HERE, "hostcheck", "hostcheck.cpp")], check=True)
    return C.CDLL(out)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def standins(hc):
    """name -> host implementation, for every nvsr_b200.ops entry the autograd module calls"""
    def pack_plane(p, dtype=0):
        p = p.detach()
        return (p[0] if p.dim() == 4 else p).permute(1, 2, 0).contiguous()

    def sample_gather(ro, rd, near, far, packed, layout, z_in=None, **kw):
        pts = (ro[:, None, :] + rd[:, None, :] * z_in[..., None]).reshape(-1, 3)
        cn = 2 * (pts - torch.tensor(packed.box_lo)) / torch.tensor(packed.box_rng) - 1
        feats = []
        for d in range(3):
            grid = (cn @ torch.tensor(packed.proj[d])).reshape(1, -1, 1, 2)
            img = packed.planes[d].permute(2, 0, 1)[None]
            feats.append(F.grid_sample(img, grid, mode="bilinear", align_corners=True, padding_mode="border")[0, :, :, 0].t())
        return torch.cat(feats, 1).contiguous(), torch.stack(feats, 0).mean(0).contiguous(), z_in

    def sample_gather_bwd(ro, rd, z, packed, gp, gm, acc=None):
        if acc is None:
            acc = [torch.zeros(tuple(p.shape[:2]) + (packed.channels,)) for p in packed.planes]
        n, S = z.shape
        rh = (C.c_int * 3)(*[a.shape[0] for a in acc])
        rw = (C.c_int * 3)(*[a.shape[1] for a in acc])
        # every buffer handed to the C side is held in a local until the call returns
        proj, lo, rng = torch.tensor(packed.proj).contiguous(), torch.tensor(packed.box_lo), torch.tensor(packed.box_rng)
        ro_c, rd_c, z_c = ro.contiguous(), rd.contiguous(), z.contiguous()
        gp_c, gm_c = (None if gp is None else gp.contiguous()), (None if gm is None else gm.contiguous())
        hc.hc_gather_bwd(rh, rw, packed.channels, _p(lo), _p(rng), _p(proj), _p(ro_c), _p(rd_c), _p(z_c), C.c_int64(n), S,
                         _p(gp_c), _p(gm_c), _p(acc[0]), _p(acc[1]), _p(acc[2]))
        return acc

    def viewdir_gather(vd, packed):
        az_lo, az_rng, el_lo, el_rng = packed.view_lo_rng
        ae = O.cart2az_el(vd)
        g = torch.stack([2 * (ae[:, 0] - az_lo) / az_rng - 1, 2 * (ae[:, 1] - el_lo) / el_rng - 1], -1).reshape(1, -1, 1, 2)
        img = packed.vplane.permute(2, 0, 1)[None]
        return F.grid_sample(img, g, mode="bilinear", align_corners=True, padding_mode="border")[0, :, :, 0].t().contiguous()

    def viewdir_gather_bwd(vd, packed, g, acc):
        az_lo, az_rng, el_lo, el_rng = packed.view_lo_rng
        vd_c, g_c = vd.contiguous(), g.contiguous()
        hc.hc_viewdir_gather_bwd(_p(vd_c), C.c_int64(vd.shape[0]), acc.shape[0], acc.shape[1], acc.shape[2],
                                 C.c_float(az_lo), C.c_float(az_rng), C.c_float(el_lo), C.c_float(el_rng), _p(g_c), _p(acc))
        return acc

    def composite(raw_planar, z, rd, S, noise=None, white_background=False, mip=False, want_weights=False, **kw):
        rf = raw_planar.t().reshape(-1, S, 4)
        rgb, disp, acc, w, depth = O.volume_render_radiance_field(rf, z, rd, 1.0 if noise is not None else 0.0, white_background,
                                                                  mip_nerf=mip, noise=noise)
        return {"rgb": rgb, "disp": disp, "acc": acc, "weights": w, "depth": depth}

    def composite_bwd(rf, z, rd, d_rgb, d_acc=None, d_depth=None, d_weights=None, noise=None, white_background=False, mip=False):
        out = torch.empty_like(rf)
        c = [None if t is None else t.contiguous() for t in (noise, d_rgb, d_acc, d_depth, d_weights)]
        rf_c, z_c, rd_c = rf.contiguous(), z.contiguous(), rd.contiguous()
        hc.hc_composite_bwd(_p(rf_c), _p(z_c), _p(rd_c), _p(c[0]), C.c_int64(rf.shape[0]),
                            rf.shape[1], int(white_background), int(mip), _p(c[1]), _p(c[2]), _p(c[3]), _p(c[4]), _p(out))
        return out

    def prepare_rays(ro_in, rd_in, use_ndc, H, W, focal, near):
        assert not use_ndc
        rd_f = rd_in.reshape(-1, 3)
        return ro_in.reshape(-1, 3).contiguous(), rd_f.contiguous(), (rd_f / rd_f.norm(p=2, dim=-1).unsqueeze(-1)).contiguous()

    def sample_pdf(bins, weights, num_samples, det=False, u=None, **kw):
        return O.sample_pdf(bins, weights, num_samples, det=det, u=u)

    def ipe(z_edges, ro, rd, radius, n_freqs, **kw):
        means, covs = O.cast_rays(z_edges, ro, rd, torch.full((ro.shape[0], 1), float(radius)))
        e = O.integrated_pos_enc(means, covs, n_freqs + 1)
        return e.reshape(-1, e.shape[-1]).contiguous()

    def dir_encoding(dirs, n_freqs, include_input=True):
        return O.positional_encoding(dirs, n_freqs, include_input)

    def sort_cat(a, b):
        return torch.sort(torch.cat((a, b), -1), -1).values.contiguous()

    return dict(sort_cat=sort_cat, pack_plane=pack_plane, sample_gather=sample_gather, sample_gather_bwd=sample_gather_bwd,
                viewdir_gather=viewdir_gather, viewdir_gather_bwd=viewdir_gather_bwd, composite=composite,
                composite_bwd=composite_bwd, prepare_rays=prepare_rays, sample_pdf=sample_pdf, ipe=ipe,
                dir_encoding=dir_encoding)
