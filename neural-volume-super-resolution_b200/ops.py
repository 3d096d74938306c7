"""Stage-level host wrappers over the C-ABI (include/nvsr.h).

Every function takes/returns torch CUDA tensors, launches on torch's current stream and raises
`NvsrError` on a non-zero status.  torch is plumbing here (device memory + streams); all arithmetic
happens in libnvsr_b200.so.  Signatures mirror the reference functions they replace; the reference
file:line for each is in include/nvsr.h.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (BLK_RAYS, BLK_SAMPLES, FEAT_ROWMAJOR_F32, FEAT_TILE_BF16, FEAT_TILE_F16, NVSR_BF16, NVSR_F16, NVSR_F32,
                   ROWS_BLOCKED, ROWS_RAY_MAJOR, TILE_ROWS)

TORCH_DTYPE = {NVSR_F32: torch.float32, NVSR_BF16: torch.bfloat16, NVSR_F16: torch.float16}
FEAT_LAYOUT = {NVSR_F32: FEAT_ROWMAJOR_F32, NVSR_BF16: FEAT_TILE_BF16, NVSR_F16: FEAT_TILE_F16}
LAYOUT_DTYPE = {FEAT_TILE_BF16: torch.bfloat16, FEAT_TILE_F16: torch.float16}
# row order of the rows a feature layout carries (include/nvsr.h): the 16-bit tile images written by the
# gather are BLOCKED (8 adjacent rays x 16 samples per 128-row tile), everything fp32 is ray-major
LAYOUT_ROWS = {FEAT_ROWMAJOR_F32: ROWS_RAY_MAJOR, FEAT_TILE_BF16: ROWS_BLOCKED, FEAT_TILE_F16: ROWS_BLOCKED}


def rows_padded(n_rays, n_samples, row_order):
    """rows a feature/raw buffer holds for n_rays x n_samples points in `row_order` (nvsr_rows_padded)."""
    if row_order == ROWS_BLOCKED:
        return -(-n_rays // BLK_RAYS) * -(-n_samples // BLK_SAMPLES) * TILE_ROWS
    return n_rays * n_samples


def raw_buffer(n_rays, n_samples, row_order, device):
    """planar raw [4, stride] for the decoder heads (stride padded to whole tiles)"""
    rows = rows_padded(n_rays, n_samples, row_order)
    stride = (rows + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS
    return torch.empty((4, stride), dtype=torch.float32, device=device)


def raw_to_nsc(raw, n_rays, n_samples, row_order):
    """planar raw [4, stride] in `row_order` -> the reference's radiance_field layout [N, S, 4]"""
    if row_order == ROWS_RAY_MAJOR:
        return raw[:, :n_rays * n_samples].t().reshape(n_rays, n_samples, 4)
    nb, ts = -(-n_rays // BLK_RAYS), -(-n_samples // BLK_SAMPLES)
    r = raw[:, :nb * ts * TILE_ROWS].reshape(4, nb, ts, BLK_SAMPLES, BLK_RAYS)   # [ch, block, sblock, s%16, ray%8]
    r = r.permute(1, 4, 2, 3, 0).reshape(nb * BLK_RAYS, ts * BLK_SAMPLES, 4)
    return r[:n_rays, :n_samples]


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """the current stream of the current device as a raw handle.  torch.cuda.current_stream() costs ~20 us of Python
    per call (device-index bookkeeping); the training step makes ~60 kernel calls, so the raw accessor is used when the
    build has it."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _OnDevice:
    """`with _OnDevice(d)` only when d is not already the current device (the guard costs ~10 us per call)."""
    __slots__ = ("guard",)

    def __init__(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.guard = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f32c(t, device=None):
    """contiguous fp32 CUDA tensor (no copy when already so)"""
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.NvsrError(f"{name} must be a CUDA tensor: nvsr_b200 has no CPU path")


# Launch accounting: every C-ABI call below launches exactly one kernel of ours.  `LAUNCHES` counts
# them (bench.py reports it as gpu_launches); when `PROFILE` is a list, each call is bracketed by CUDA
# events on the launching stream and appended as (name, start, end, meta) for the roofline numbers.
LAUNCHES = {}
PROFILE = None


def _call(name, fn, *args, **meta):
    LAUNCHES[name] = LAUNCHES.get(name, 0) + 1
    if PROFILE is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = fn(*args)
    e1.record()
    PROFILE.append((name, e0, e1, meta))
    return st


# ---------------------------------------------------------------------------------------------
# a1  get_ray_bundle
def get_ray_bundle(height, width, focal_length, tform_cam2world, padding_size=0, downsampling_offset=0,
                   row_range=None):
    """Drop-in for nerf_helpers.get_ray_bundle (nerf_helpers.py:507-549).

    Returns (ray_origins, ray_directions), each [H+2p, W+2p, 3] (or the [row_begin,row_end) band when
    `row_range` is given — the multi-GPU row-band sharding).  `focal_length` may be a scalar or an
    [fx?, fy?] list; the reference divides x by get_focal(f,'H') and y by get_focal(f,'W')
    (nerf_helpers.py:432-437, 540-541), which is mirrored here.
    """
    lib = _lib.load()
    _require_cuda(tform_cam2world, "tform_cam2world")
    if isinstance(focal_length, (list, tuple)):
        fx, fy = float(focal_length[1]), float(focal_length[0])  # get_focal(.,'H') / get_focal(.,'W')
    else:
        fx = fy = float(focal_length)
    hp, wp = height + 2 * padding_size, width + 2 * padding_size
    r0, r1 = (0, hp) if row_range is None else row_range
    if tform_cam2world.numel() != 16:
        raise _lib.NvsrError("tform_cam2world must be 4x4")
    # the pose stays on the device: the kernel reads it there (no .cpu() round trip, no sync per frame)
    c2w = _f32c(tform_cam2world.detach()).reshape(-1)
    dev = tform_cam2world.device
    ro = torch.empty((r1 - r0, wp, 3), dtype=torch.float32, device=dev)
    rd = torch.empty_like(ro)
    with _OnDevice(dev):
        st = _call("nvsr_ray_bundle", lib.nvsr_ray_bundle_dev, height, width, fx, fy, _ptr(c2w), padding_size,
                   float(downsampling_offset), r0, r1, _ptr(ro), _ptr(rd), _stream())
    _lib.check(st, "nvsr_ray_bundle_dev")
    return ro, rd


# a2/a3
def prepare_rays(ray_origins, ray_directions, use_ndc=False, height=0, width=0, focal=1.0, ndc_near=1.0,
                 want_viewdirs=True):
    """viewdirs = rd/|rd| and optional ndc_rays (train_utils.py:210-221, nerf_helpers.py:578-605)."""
    lib = _lib.load()
    ro = _f32c(ray_origins.reshape(-1, 3))
    rd = _f32c(ray_directions.reshape(-1, 3))
    _require_cuda(ro, "ray_origins")
    n = ro.shape[0]
    ro_o = torch.empty_like(ro)
    rd_o = torch.empty_like(rd)
    vd = torch.empty_like(rd) if want_viewdirs else None
    if isinstance(focal, (list, tuple)):
        raise _lib.NvsrError("ndc_rays needs a scalar focal (as in the reference)")
    with _OnDevice(ro.device):
        st = _call("nvsr_prepare_rays", lib.nvsr_prepare_rays, _ptr(ro), _ptr(rd), n, int(bool(use_ndc)), int(height), int(width), float(focal),
                                   float(ndc_near), _ptr(ro_o), _ptr(rd_o), _ptr(vd), _stream())
    _lib.check(st, "nvsr_prepare_rays")
    return ro_o, rd_o, vd


# ---------------------------------------------------------------------------------------------
F16_MAX = 65504.0


def _check_f16_range(t, what):
    """fp16 images saturate (cvt.satfinite) instead of overflowing to inf — silently.  Packing happens once per scene /
    weight update, so one reduction + host read here is free; a tensor that does not fit is refused loudly."""
    if t.numel() and float(t.abs().max()) > F16_MAX:
        raise _lib.NvsrError(f"{what}: |value| exceeds the fp16 range ({F16_MAX:g}); values would saturate silently - "
                             "use set_precision('bf16') (fp32 range) or 'fp32' for this scene")


class DeferredRangeCheck:
    """fp16 range check WITHOUT a host synchronisation, for paths that re-pack every step (training): the maximum of the
    tensors packed in one step is reduced on the device and copied to pinned memory asynchronously; the NEXT step (or
    `flush()`) reads it and raises — one step late, but loudly, and the step itself never waits for the GPU."""

    def __init__(self):
        self.pending = None      # (pinned host tensor, event, description)
        self.cur = []
        self.graph_max = None    # running maximum kept on the device by steps replayed from a CUDA graph
        self.graph_what = ""
        self.dev = {}            # device -> fp32 scalar the pack kernels fold max |value| into (`absmax` arguments)
        self.dev_used = []

    def device_max(self, device, what):
        """the device scalar a C entry with an `absmax` argument accumulates into; counts as an `add` of this step"""
        device = torch.device(device)
        t = self.dev.get(device)
        if t is None:
            t = self.dev[device] = torch.zeros((), dtype=torch.float32, device=device)
        self.dev_used.append((t, what))
        return t

    def add(self, t, what):
        self.cur.append((t.detach().abs().max(), what))

    def commit(self):
        if torch.cuda.is_current_stream_capturing():
            # a step being captured into a CUDA graph: no events, no host copies — fold the maximum into a device scalar
            # that `check_graph()` reads between replays (the pack kernels' own accumulators just keep accumulating)
            if self.dev_used:
                self.graph_what = ", ".join(sorted({w for _, w in self.dev_used} | ({self.graph_what} if self.graph_what else set())))
                self.dev_used = []
            if self.cur:
                m = torch.stack([v.float() for v, _ in self.cur]).max()
                self.graph_what = ", ".join(sorted({w for _, w in self.cur}))
                self.cur = []
                if self.graph_max is None:
                    self.graph_max = torch.zeros_like(m)
                self.graph_max.copy_(torch.maximum(self.graph_max, m))
            return
        self.poll()
        if not self.cur and not self.dev_used:
            return
        vals = [v.float() for v, _ in self.cur] + [t for t, _ in {id(t): (t, w) for t, w in self.dev_used}.values()]
        m = vals[0] if len(vals) == 1 else torch.stack(vals).max()
        what = ", ".join(sorted({w for _, w in self.cur} | {w for _, w in self.dev_used}))
        used, self.cur, self.dev_used = self.dev_used, [], []
        if self.pending is None:         # at most one check in flight; a step whose check is skipped is covered by the next
            host = torch.empty((), dtype=torch.float32).pin_memory()
            host.copy_(m, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self.pending = (host, ev, what)
            for t in {id(t): t for t, _ in used}.values():
                t.zero_()        # a skipped copy leaves the accumulator running: the next check still sees its maximum

    def poll(self, wait=False):
        if self.pending is None:
            return
        host, ev, what = self.pending
        if wait:
            ev.synchronize()
        if ev.query():
            self.pending = None
            if not (float(host) <= F16_MAX):
                raise _lib.NvsrError(f"{what}: |value| exceeded the fp16 range ({F16_MAX:g}) in the previous step; values "
                                     "saturated silently - use the fp32 decoder mode for this model")

    def flush(self):
        self.poll(wait=True)

    def check_graph(self):
        """host read (synchronises) of the maximum accumulated by graph replays; raises like `poll`"""
        vals = ([self.graph_max] if self.graph_max is not None else []) + list(self.dev.values())
        if any(not (float(v) <= F16_MAX) for v in vals):
            raise _lib.NvsrError(f"{self.graph_what}: |value| exceeded the fp16 range ({F16_MAX:g}) in a replayed step; values "
                                 "saturated silently - use the fp32 decoder mode for this model")


def pack_plane(plane_nchw, dtype=NVSR_F32, range_check=None):
    """[1,C,Rh,Rw] fp32 (models.py:436-439) -> device plane image: fp32 channels-last [Rh,Rw,C], or 16-bit
    x-pair records [Rh,C/8,Rw,2,8] (nvsr.h: 8-channel chunk of a texel followed by its right neighbour's).
    `range_check`: a DeferredRangeCheck to use instead of the synchronous fp16 range check."""
    lib = _lib.load()
    p = _f32c(plane_nchw.detach())
    _require_cuda(p, "plane")
    if dtype == NVSR_F16:
        if range_check is None:
            _check_f16_range(p, "pack_plane")
        else:
            range_check.add(p, "pack_plane")
    if p.dim() == 4:
        assert p.shape[0] == 1
        p = p[0]
    c, rh, rw = p.shape
    if dtype == NVSR_F32:
        out = torch.empty((rh, rw, c), dtype=torch.float32, device=p.device)
    else:
        if c % 8:
            raise ValueError("16-bit planes need a channel count that is a multiple of 8")
        out = torch.empty((rh, c // 8, rw, 2, 8), dtype=TORCH_DTYPE[dtype], device=p.device)
    with _OnDevice(p.device):
        st = _call("nvsr_pack_plane", lib.nvsr_pack_plane, _ptr(p), c, rh, rw, _ptr(out), dtype, _stream())
    _lib.check(st, "nvsr_pack_plane")
    return out


def pack_weight16(weight, k_pad=None, dtype=NVSR_BF16, range_check=None):
    """nn.Linear weight [n_out,k] (may be a column-slice view) -> UMMA image [k_pad/8, n_out, 8] bf16|fp16.
    `range_check`: a DeferredRangeCheck to use instead of the synchronous fp16 range check."""
    lib = _lib.load()
    w = weight.detach()
    _require_cuda(w, "weight")
    if w.dtype != torch.float32 or w.stride(1) != 1:
        w = w.float().contiguous()
    if dtype == NVSR_F16:
        if range_check is None:
            _check_f16_range(w, "pack_weight16")
        else:
            range_check.add(w, "pack_weight16")
    n_out, k = w.shape
    ldw = w.stride(0)
    if k_pad is None:
        k_pad = (k + 15) // 16 * 16
    out = torch.empty((k_pad // 8, n_out, 8), dtype=TORCH_DTYPE[dtype], device=w.device)
    with _OnDevice(w.device):
        st = _call("nvsr_pack_weight16", lib.nvsr_pack_weight16, _ptr(w), n_out, k, ldw, k_pad, _ptr(out), dtype, _stream())
    _lib.check(st, "nvsr_pack_weight16")
    return out


# ---------------------------------------------------------------------------------------------
class PackedPlanes:
    """Device-resident packed position planes of one scene (nvsr_pack_plane images) + box + projection matrices."""

    def __init__(self, planes, dtype, box_lo, box_rng, proj, vplane=None, view_lo_rng=None, combine="avg"):
        if combine not in ("avg", "sum"):
            raise NotImplementedError(f"nvsr_b200: proj_combination={combine!r} (supported: 'avg', 'sum')")
        self.combine = combine          # combine_pos_planes for the density features (models.py:355-361)
        self.planes = planes            # list of 3 tensors: fp32 [Rh,Rw,C] or 16-bit [Rh,C/8,Rw,2,8]
        self.dtype = dtype
        self.box_lo = [float(v) for v in box_lo]
        self.box_rng = [float(v) for v in box_rng]
        self.proj = proj                # 3 x [3][2] nested lists
        self.vplane = vplane            # [Rh,Rw,C] fp32 view-direction plane or None
        self.view_lo_rng = view_lo_rng  # (az_lo, az_rng, el_lo, el_rng)
        self.channels = planes[0].shape[-1] if planes[0].dim() == 3 else planes[0].shape[1] * 8

    def cstruct(self):
        s = _lib.Planes()
        for d in range(3):
            s.plane[d] = self.planes[d].data_ptr()
            s.rh[d], s.rw[d] = self.planes[d].shape[0], self.planes[d].shape[1 if self.planes[d].dim() == 3 else 2]
            s.box_lo[d], s.box_rng[d] = self.box_lo[d], self.box_rng[d]
            for i in range(3):
                for j in range(2):
                    s.proj[d][i * 2 + j] = float(self.proj[d][i][j])
        s.channels = self.channels
        s.dtype = self.dtype
        s.combine = 1 if self.combine == "sum" else 0
        return s


def feature_buffers(n_rays, n_samples, channels, layout, device, density_only=False):
    """Allocate (featP, featM) for n_rays x n_samples points in the given layout."""
    rows = rows_padded(n_rays, n_samples, LAYOUT_ROWS[layout])
    if density_only and layout != FEAT_ROWMAJOR_F32:
        return (None, torch.empty((rows // TILE_ROWS, channels // 8, TILE_ROWS, 8), dtype=LAYOUT_DTYPE[layout], device=device))
    if layout == FEAT_ROWMAJOR_F32:
        return (None if density_only else torch.empty((rows, 3 * channels), dtype=torch.float32, device=device),
                torch.empty((rows, channels), dtype=torch.float32, device=device))
    tiles = rows // TILE_ROWS
    dt = LAYOUT_DTYPE[layout]
    return (torch.empty((tiles, 3 * channels // 8, TILE_ROWS, 8), dtype=dt, device=device),
            torch.empty((tiles, channels // 8, TILE_ROWS, 8), dtype=dt, device=device))


def sample_gather(ro, rd, near, far, packed, layout, t_vals=None, z_in=None, t_rand=None, lindisp=False,
                  n_samples=None, out=None, want_z=True, density_only=False):
    """Fused stratified sampler + tri-plane gather (train_utils.py:95-111 + models.py:381-391 xyz half).

    Returns (featP, featM, z_vals[n,S] or None).  `density_only` (tile layouts): only the mean features featM are
    written (featP is None) — the sparse colour path gathers the 3-plane features of the contributing rows later."""
    lib = _lib.load()
    n = ro.shape[0]
    if z_in is not None:
        S = z_in.shape[1]
        z_in = _f32c(z_in)
    else:
        S = t_vals.numel() if n_samples is None else n_samples
        t_vals = _f32c(t_vals)
    rows = n * S
    if out is None:
        out = feature_buffers(n, S, packed.channels, layout, ro.device, density_only=density_only)
    feat_p, feat_m = out
    z_out = torch.empty((n, S), dtype=torch.float32, device=ro.device) if (want_z and z_in is None) else None
    s = _lib.Sampler()
    s.n_rays, s.n_samples = n, S
    s.ro, s.rd = ro.data_ptr(), rd.data_ptr()
    s.near_, s.far_, s.lindisp = float(near), float(far), int(bool(lindisp))
    s.t_vals = 0 if t_vals is None else t_vals.data_ptr()
    s.t_rand = 0 if t_rand is None else t_rand.data_ptr()
    s.z_in = 0 if z_in is None else z_in.data_ptr()
    pl = packed.cstruct()
    with _OnDevice(ro.device):
        st = _call("nvsr_sample_gather", lib.nvsr_sample_gather, C.byref(s), C.byref(pl), layout, _ptr(feat_p),
                   _ptr(feat_m), _ptr(z_out), _stream(),
                   # algorithmic HBM bytes: feature write 4C*e per row + z (4 B) + rays (24 B/ray)
                   bytes=rows * ((1 if feat_p is None else 4) * packed.channels * (4 if layout == FEAT_ROWMAJOR_F32 else 2) + 4)
                   + n * 24, rows=rows)
    _lib.check(st, "nvsr_sample_gather")
    return feat_p, feat_m, (z_out if z_in is None else z_in)


def sample_gather_hilo(ro, rd, near, far, packed, lo_planes, t_vals=None, z_in=None, t_rand=None, lindisp=False,
                       n_samples=None, density_only=False, want_m16=False):
    """The fp16-split mode's gather (nvsr_sample_gather_hilo): `packed` holds the fp16 x-pair images of fp16(p),
    `lo_planes` those of p - fp16(p).  Returns (featP fp16 tile image or None, featM32 fp32 tile image
    [tiles, C/4, 128, 4] interpolated from hi + lo, z [n,S]); with want_m16 a 4-tuple whose third entry is the combined
    feature as an fp16 tile image [tiles, C/8, 128, 8] as well."""
    lib = _lib.load()
    n = ro.shape[0]
    if z_in is not None:
        S = z_in.shape[1]
        z_in = _f32c(z_in)
    else:
        S = t_vals.numel() if n_samples is None else n_samples
        t_vals = _f32c(t_vals)
    rows = rows_padded(n, S, ROWS_BLOCKED)
    tiles, Cc = rows // TILE_ROWS, packed.channels
    feat_p = None if density_only else torch.empty((tiles, 3 * Cc // 8, TILE_ROWS, 8), dtype=torch.float16, device=ro.device)
    feat_m = torch.empty((tiles, Cc // 4, TILE_ROWS, 4), dtype=torch.float32, device=ro.device)
    feat_m16 = torch.empty((tiles, Cc // 8, TILE_ROWS, 8), dtype=torch.float16, device=ro.device) if want_m16 else None
    z_out = torch.empty((n, S), dtype=torch.float32, device=ro.device) if z_in is None else None
    s = _lib.Sampler()
    s.n_rays, s.n_samples = n, S
    s.ro, s.rd = ro.data_ptr(), rd.data_ptr()
    s.near_, s.far_, s.lindisp = float(near), float(far), int(bool(lindisp))
    s.t_vals = 0 if t_vals is None else t_vals.data_ptr()
    s.t_rand = 0 if t_rand is None else t_rand.data_ptr()
    s.z_in = 0 if z_in is None else z_in.data_ptr()
    pl = packed.cstruct()
    lo = (C.c_void_p * 3)(*[t.data_ptr() for t in lo_planes])
    with _OnDevice(ro.device):
        st = _call("nvsr_sample_gather", lib.nvsr_sample_gather_hilo, C.byref(s), C.byref(pl), lo, _ptr(feat_p), _ptr(feat_m),
                   _ptr(feat_m16), _ptr(z_out), _stream(), rows=n * S,
                   bytes=n * S * ((0 if feat_p is None else 3) * Cc * 2 + Cc * 4 + 4) + n * 24)
    _lib.check(st, "nvsr_sample_gather_hilo")
    if want_m16:
        return feat_p, feat_m, feat_m16, (z_out if z_in is None else z_in)
    return feat_p, feat_m, (z_out if z_in is None else z_in)


def keep_rows(raw, n_rays, n_samples, noise=None):
    """Rows of a BLOCKED raw buffer that can contribute to the maps: sigma (+ noise) > 0 (or NaN) — every other
    sample has alpha = 0 and weight exactly 0 (volume_rendering_utils.py:29-44).  Returns (row_ids int32
    [rows_padded], count int32 [1]) on the device; the list order is unspecified."""
    lib = _lib.load()
    rows = rows_padded(n_rays, n_samples, ROWS_BLOCKED)
    keep = torch.empty((rows,), dtype=torch.int32, device=raw.device)
    count = torch.zeros((1,), dtype=torch.int32, device=raw.device)
    noise = None if noise is None else _f32c(noise)
    with _OnDevice(raw.device):
        st = _call("nvsr_keep_rows", lib.nvsr_keep_rows, _ptr(raw[3]), _ptr(noise), n_rays, n_samples, _ptr(keep),
                   _ptr(count), _stream(), bytes=rows * 4, rows=rows)
    _lib.check(st, "nvsr_keep_rows")
    return keep, count


def sample_gather_rows(ro, rd, packed, layout, z, keep, count, out=None):
    """3-plane features (featP) of the listed rows only, densely packed in list order (nvsr_sample_gather_rows)."""
    lib = _lib.load()
    n, S = z.shape
    max_rows = keep.numel()
    if out is None:
        out = torch.empty((max_rows // TILE_ROWS, 3 * packed.channels // 8, TILE_ROWS, 8), dtype=LAYOUT_DTYPE[layout],
                          device=ro.device)
    s = _lib.Sampler()
    s.n_rays, s.n_samples = n, S
    s.ro, s.rd = ro.data_ptr(), rd.data_ptr()
    s.near_, s.far_, s.lindisp = 0.0, 1.0, 0
    s.t_vals, s.t_rand, s.z_in = 0, 0, z.data_ptr()
    pl = packed.cstruct()
    with _OnDevice(ro.device):
        st = _call("nvsr_sample_gather_rows", lib.nvsr_sample_gather_rows, C.byref(s), C.byref(pl), layout, _ptr(keep),
                   _ptr(count), max_rows, _ptr(out), _stream(), count=count, bytes_per_row=3 * packed.channels * 2)
    _lib.check(st, "nvsr_sample_gather_rows")
    return out


def viewdir_gather(viewdirs, packed):
    """cart2az_el + normalize + project_viewdir per ray (nerf_helpers.py:492-496, models.py:312-326)."""
    lib = _lib.load()
    vd = _f32c(viewdirs)
    n = vd.shape[0]
    vp = packed.vplane
    out = torch.empty((n, vp.shape[-1]), dtype=torch.float32, device=vd.device)
    az_lo, az_rng, el_lo, el_rng = packed.view_lo_rng
    with _OnDevice(vd.device):
        st = _call("nvsr_viewdir_gather", lib.nvsr_viewdir_gather, _ptr(vd), n, _ptr(vp), vp.shape[0], vp.shape[1], vp.shape[2], az_lo, az_rng,
                                     el_lo, el_rng, _ptr(out), _stream())
    _lib.check(st, "nvsr_viewdir_gather")
    return out


def row_bias(vin, weight_cols, bias):
    """out[ray] = bias + weight_cols @ vin[ray]; weight_cols may be a column-slice view of a Linear weight."""
    lib = _lib.load()
    vin = _f32c(vin)
    w = weight_cols.detach()
    if w.dtype != torch.float32 or w.stride(1) != 1:
        w = w.float().contiguous()
    n, k = vin.shape
    n_out = w.shape[0]
    assert w.shape[1] == k
    b = None if bias is None else _f32c(bias.detach())
    out = torch.empty((n, n_out), dtype=torch.float32, device=vin.device)
    with _OnDevice(vin.device):
        st = _call("nvsr_row_bias", lib.nvsr_row_bias, _ptr(vin), n, k, _ptr(w), w.stride(0), _ptr(b), n_out, _ptr(out), _stream())
    _lib.check(st, "nvsr_row_bias")
    return out


# ---------------------------------------------------------------------------------------------
class ChainLayer:
    """One dense layer of a decoder chain (+ optional fp32 output head tapped on its output)."""

    def __init__(self, w, bias, k, n_out, relu, row_bias=None, head_w=None, head_b=None, head_ch=0):
        self.w, self.bias, self.k, self.n_out, self.relu = w, bias, k, n_out, relu
        self.row_bias, self.head_w, self.head_b, self.head_ch = row_bias, head_w, head_b, head_ch


def _mlp_struct(inp, layers, rows, raw, precision, samples_per_ray, n_rays, row_order, row_ids=None, row_count=None):
    m = _lib.Mlp()
    m.precision = precision
    m.n_layers = len(layers)
    for i, ly in enumerate(layers):
        c = m.layer[i]
        c.w = ly.w.data_ptr()
        c.bias = 0 if ly.bias is None else ly.bias.data_ptr()
        c.row_bias = 0 if ly.row_bias is None else ly.row_bias.data_ptr()
        c.head_w = 0 if ly.head_w is None else ly.head_w.data_ptr()
        c.head_b = 0 if ly.head_b is None else ly.head_b.data_ptr()
        c.k, c.n_out, c.relu = ly.k, ly.n_out, int(bool(ly.relu))
        c.head_n = 0 if ly.head_w is None else ly.head_w.shape[0]
        c.head_ch = ly.head_ch
    m.in_ = inp.data_ptr()
    m.rows = rows
    m.samples_per_ray = samples_per_ray
    m.n_rays = n_rays
    m.raw = raw.data_ptr()
    m.raw_stride = raw.stride(0)
    m.row_order = row_order
    m.row_ids = 0 if row_ids is None else row_ids.data_ptr()
    m.row_count = 0 if row_count is None else row_count.data_ptr()
    # true MACs x2 only (no padding): layers + heads
    fpr = 2 * sum(ly.k * ly.n_out + (0 if ly.head_w is None else ly.head_w.shape[0] * ly.n_out) for ly in layers)
    bpr = layers[0].k * (4 if precision == NVSR_F32 else 2) + 4 * sum(
        0 if ly.head_w is None else ly.head_w.shape[0] for ly in layers)
    return m, fpr, bpr


def mlp_chain(inp, layers, rows, raw, precision, samples_per_ray=1, n_rays=1, row_order=ROWS_RAY_MAJOR,
              row_ids=None, row_count=None):
    """Evaluate one decoder chain (models.py:393-421 / :85-108) over `rows` rows into planar raw [4,stride].
    `rows` counts the rows of the input buffer (padded rows included for ROWS_BLOCKED).  With `row_ids` /
    `row_count` (sparse colour path) input row i stands for BLOCKED row row_ids[i] and row_count[0] rows are
    evaluated; `rows` is then the capacity of the input buffer."""
    lib = _lib.load()
    m, fpr, bpr = _mlp_struct(inp, layers, rows, raw, precision, samples_per_ray, n_rays, row_order, row_ids, row_count)
    with _OnDevice(raw.device):
        # `count` (sparse): rows actually evaluated, on the device
        st = _call("nvsr_mlp_chain", lib.nvsr_mlp_chain, C.byref(m), _stream(), rows=rows, count=row_count,
                   flops_per_row=fpr, bytes_per_row=bpr, flops=rows * fpr, bytes=rows * bpr)
    _lib.check(st, "nvsr_mlp_chain")
    return raw


def mlp_chain_split(feat, w_hi, w_lo, biases, head_w, head_b, head_ch, n_rays, n_samples, raw):
    """One tri-plane decoder chain with split fp16 operands (three tcgen05 passes per layer: fp32-grade accuracy).
    feat: fp32 [n_rays*n_samples, k0] ray-major, or the fp32 tile image [tiles, k0/4, 128, 4] of `sample_gather_hilo`;
    w_hi / w_lo: lists of 4 fp16 weight images; raw: BLOCKED planar buffer."""
    lib = _lib.load()
    tiled = feat.dim() == 4
    k0 = feat.shape[1] * 4 if tiled else feat.shape[1]
    P = C.c_void_p * 4
    wh, wl, bs = P(*[w.data_ptr() for w in w_hi]), P(*[w.data_ptr() for w in w_lo]), P(*[b.data_ptr() for b in biases])
    rows = n_rays * n_samples
    with _OnDevice(raw.device):
        fpr = 2 * (k0 * 128 + 3 * 128 * 128 + head_w.shape[0] * 128)
        st = _call("nvsr_mlp_chain_split", lib.nvsr_mlp_chain_split_tiled if tiled else lib.nvsr_mlp_chain_split, _ptr(feat), k0,
                   wh, wl, bs, _ptr(head_w), _ptr(head_b),
                   head_w.shape[0], head_ch, n_rays, n_samples, _ptr(raw), raw.stride(0), _stream(), rows=rows,
                   flops_per_row=fpr, flops=rows * fpr, bytes=rows * (4 * k0 + 4 * head_w.shape[0]))
    _lib.check(st, "nvsr_mlp_chain_split")
    return raw


# ---- decoder training path on the tensor cores (csrc/train_tc.cu) ------------------------
ACT_TILE_ELEMS = TILE_ROWS * 128


def mlp_chain_train(inp, layers, rows, raw, samples_per_ray, n_rays, row_ids=None, row_count=None):
    """`mlp_chain` (fp16, BLOCKED rows, one of the two tri-plane chains) that also returns the four activation images
    x_1..x_4 [tiles,16,128,8] fp16 the backward needs.  row_ids / row_count (rgb chain): the sparse colour path — `inp`
    and the returned images are in LIST order, the heads go to the listed rows of `raw`."""
    lib = _lib.load()
    m, fpr, bpr = _mlp_struct(inp, layers, rows, raw, NVSR_F16, samples_per_ray, n_rays, ROWS_BLOCKED, row_ids, row_count)
    tiles = rows // TILE_ROWS
    acts = [torch.empty((tiles, 16, TILE_ROWS, 8), dtype=torch.float16, device=raw.device) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[a.data_ptr() for a in acts])
    with _OnDevice(raw.device):
        st = _call("nvsr_mlp_chain_train", lib.nvsr_mlp_chain_train, C.byref(m), ptrs, _stream(), rows=rows,
                   flops=rows * fpr, bytes=rows * (bpr + 4 * 256))
    _lib.check(st, "nvsr_mlp_chain_train")
    return acts


def mlp_dgrad(w_imgs, k0, head_w, head_ch, d_raw, scale, acts, n_rays, n_samples, row_count=None, row_ids=None, x0_img=None,
              acts_listed=False):
    """Data-gradient chain of one tri-plane decoder chain.  w_imgs: the 4 forward weight images (fp16); head_w [h,128]
    fp32; d_raw planar [4,stride] (BLOCKED rows, padding rows 0); acts: x_1..x_4 images.
    -> (g images g_0..g_3, d_out image [tiles,2,128,8], d_x0 fp32 [n_rays*n_samples, k0]).
    row_count (device int32 [1]): row-list mode — the outputs are LIST-ordered (d_x0 [tiles*128, k0]) and only the tiles
    the list fills are touched.  Without row_ids, d_raw / acts are the LIST-ordered copies of `compact_rows`; with
    row_ids (+ x0_img, the feature image) they are the forward's own buffers, gathered through the list by the kernel,
    which then also returns the LIST-ordered x_1..x_4 and feature images: (g, d_out, d_x0, acts_list, x0_list).
    acts_listed (with row_ids): `acts` are LIST-ordered already (the sparse training forward over the same list) — only
    d_raw is read through the list; acts_list / x0_list come back as None."""
    lib = _lib.load()
    dev = d_raw.device
    tiles = rows_padded(n_rays, n_samples, ROWS_BLOCKED) // TILE_ROWS
    g = [torch.empty((tiles, 16, TILE_ROWS, 8), dtype=torch.float16, device=dev) for _ in range(4)]
    dout = torch.empty((tiles, 2, TILE_ROWS, 8), dtype=torch.float16, device=dev)
    d_x0 = torch.empty((n_rays * n_samples if row_count is None else tiles * TILE_ROWS, k0), dtype=torch.float32, device=dev)
    head_w = _f32c(head_w.detach())
    a = _lib.Dgrad()
    for l in range(4):
        a.w[l], a.act[l], a.g[l] = w_imgs[l].data_ptr(), acts[l].data_ptr(), g[l].data_ptr()
    a.k0, a.head_w, a.head_n, a.head_ch = k0, head_w.data_ptr(), head_w.shape[0], head_ch
    a.d_raw, a.raw_stride, a.scale = d_raw.data_ptr(), d_raw.stride(0), float(scale)
    a.dout_img, a.d_x0, a.n_rays, a.n_samples = dout.data_ptr(), d_x0.data_ptr(), n_rays, n_samples
    a.row_count = _ptr(row_count)
    acts_list = x0_list = None
    if row_ids is not None and acts_listed:
        a.row_ids, a.acts_listed = row_ids.data_ptr(), 1
    elif row_ids is not None:
        acts_list = [torch.empty_like(t) for t in acts]
        a.row_ids = row_ids.data_ptr()
        for l in range(4):
            a.act_list[l] = acts_list[l].data_ptr()
        if x0_img is not None:
            x0_list = torch.empty_like(x0_img)
            a.x0_img, a.x0_list = x0_img.data_ptr(), x0_list.data_ptr()
    with _OnDevice(dev):
        rows = tiles * TILE_ROWS
        st = _call("nvsr_mlp_dgrad", lib.nvsr_mlp_dgrad, C.byref(a), _stream(), rows=rows,
                   flops=rows * 2 * (3 * 128 * 128 + k0 * 128 + head_w.shape[0] * 128), bytes=rows * (8 * 256 + 4 * k0 + 16))
    _lib.check(st, "nvsr_mlp_dgrad")
    if row_ids is not None:
        return g, dout, d_x0, acts_list, x0_list
    return g, dout, d_x0


def mlp_wgrad(a_img, b_img, n_b, inv_scale, dw, db=None):
    """dw[128, n_b] += inv_scale * a^T b over every row of the two tile images; db[128] += inv_scale * column sums of a."""
    lib = _lib.load()
    tiles = a_img.shape[0]
    assert b_img.shape[0] == tiles and dw.dtype == torch.float32 and dw.stride(1) == 1
    with _OnDevice(dw.device):
        st = _call("nvsr_mlp_wgrad", lib.nvsr_mlp_wgrad, _ptr(a_img), _ptr(b_img), n_b, tiles, float(inv_scale), _ptr(dw),
                   dw.stride(0), _ptr(db), _stream(), rows=tiles * TILE_ROWS, flops=tiles * TILE_ROWS * 2 * 128 * n_b,
                   bytes=tiles * TILE_ROWS * 2 * (128 + n_b))
    _lib.check(st, "nvsr_mlp_wgrad")
    return dw


def mlp_wgrad_chain(g, x0_img, k0, acts, dout, inv_scale, dws, dbs, dw_head, row_count=None):
    """the five weight gradients of one chain in ONE C call (nvsr_mlp_wgrad_chain[_rows]): layers 0..3 into dws[l]
    [128, k] / dbs[l] [128], the head into dw_head [128, 16]; all accumulated into (zero them first).
    row_count (device int32 [1]): the images are LIST-ordered (`compact_rows`); only the listed tiles are read."""
    lib = _lib.load()
    tiles = g[0].shape[0]
    P = C.c_void_p * 4
    with _OnDevice(dw_head.device):
        st = _call("nvsr_mlp_wgrad_chain", lib.nvsr_mlp_wgrad_chain_rows, P(*[t.data_ptr() for t in g]), _ptr(x0_img), k0,
                   P(*[t.data_ptr() for t in acts]), _ptr(dout), tiles, _ptr(row_count), float(inv_scale),
                   P(*[t.data_ptr() for t in dws]),
                   (C.c_int64 * 4)(*[t.stride(0) for t in dws]), P(*[t.data_ptr() for t in dbs]), _ptr(dw_head), _stream(),
                   rows=tiles * TILE_ROWS, flops=tiles * TILE_ROWS * 2 * 128 * (k0 + 3 * 128 + 16),
                   bytes=tiles * TILE_ROWS * 2 * (9 * 128 + k0 + 16))
    _lib.check(st, "nvsr_mlp_wgrad_chain")


def pack_weights16(weights, dtype=NVSR_F16, range_check=None):
    """several nn.Linear weights [n_out, k] (column-slice views allowed) -> UMMA images, ONE C call"""
    lib = _lib.load()
    ws, outs = [], []
    for w in weights:
        w = w.detach()
        _require_cuda(w, "weight")
        if w.dtype != torch.float32 or w.stride(1) != 1:
            w = w.float().contiguous()
        if dtype == NVSR_F16 and range_check is None:
            _check_f16_range(w, "pack_weights16")
        ws.append(w)
        outs.append(torch.empty(((w.shape[1] + 15) // 16 * 2, w.shape[0], 8), dtype=TORCH_DTYPE[dtype], device=w.device))
    n = len(ws)
    I = C.c_int32 * n
    # the deferred range check rides in the pack kernel (one atomic max per warp) instead of an abs().max() per weight
    absmax = range_check.device_max(ws[0].device, "pack_weights16") if (dtype == NVSR_F16 and range_check is not None) else None
    with _OnDevice(ws[0].device):
        st = _call("nvsr_pack_weights16", lib.nvsr_pack_weights16, n, (C.c_void_p * n)(*[w.data_ptr() for w in ws]),
                   I(*[w.shape[0] for w in ws]), I(*[w.shape[1] for w in ws]), I(*[w.stride(0) for w in ws]),
                   I(*[(w.shape[1] + 15) // 16 * 16 for w in ws]), (C.c_void_p * n)(*[o.data_ptr() for o in outs]), dtype,
                   _ptr(absmax), _stream())
    _lib.check(st, "nvsr_pack_weights16")
    return outs


def ray_sum(img, n_rays, n_samples, inv_scale=1.0):
    """[n_rays, 128] fp32 = inv_scale * per-ray sum over the samples of a 128-channel tile image (BLOCKED rows)."""
    lib = _lib.load()
    out = torch.empty((n_rays, 128), dtype=torch.float32, device=img.device)
    with _OnDevice(img.device):
        st = _call("nvsr_ray_sum", lib.nvsr_ray_sum, _ptr(img), n_rays, n_samples, float(inv_scale), _ptr(out), _stream())
    _lib.check(st, "nvsr_ray_sum")
    return out


def sort_cat(a, b):
    """sort(cat((a, b), -1), -1).values for [n, Sa] / [n, Sb] fp32 depths (train_utils.py:144-156) in one launch
    (nvsr_sort_cat, Sa + Sb <= 512; longer rows go through torch.sort on the device)."""
    a, b = _f32c(a), _f32c(b, a.device)
    _require_cuda(a, "depths")
    n, sa = a.shape
    sb = b.shape[1]
    if sa + sb > 512:
        return torch.sort(torch.cat((a, b), -1), -1).values.contiguous()
    lib = _lib.load()
    out = torch.empty((n, sa + sb), dtype=torch.float32, device=a.device)
    with _OnDevice(a.device):
        st = _call("nvsr_sort_cat", lib.nvsr_sort_cat, _ptr(a), sa, _ptr(b), sb, n, _ptr(out), _stream(), rows=n * (sa + sb))
    _lib.check(st, "nvsr_sort_cat")
    return out


def nonzero_rows(d_raw):
    """BLOCKED row ids whose raw gradient (planar [4, stride]) is not identically zero -> (row_ids int32 [stride],
    count int32 [1]) on the device, order unspecified (nvsr_nonzero_rows)."""
    lib = _lib.load()
    rows = d_raw.shape[1]
    ids = torch.empty((rows,), dtype=torch.int32, device=d_raw.device)
    count = torch.empty((1,), dtype=torch.int32, device=d_raw.device)
    with _OnDevice(d_raw.device):
        st = _call("nvsr_nonzero_rows", lib.nvsr_nonzero_rows, _ptr(d_raw), d_raw.stride(0), rows, _ptr(ids), _ptr(count),
                   _stream(), rows=rows, bytes=rows * 16)
    _lib.check(st, "nvsr_nonzero_rows")
    return ids, count


def compact_rows(images, d_raw, ids, count):
    """The listed rows of tile images [tiles, C/8, 128, 8] (16-bit) and (unless None) of d_raw [4, stride], packed densely
    in LIST order (nvsr_compact_rows; the last tile's tail is zero-filled).  -> (list of compact images, compact d_raw or
    None); tiles past the list are left uninitialised."""
    lib = _lib.load()
    tiles = images[0].shape[0] if d_raw is None else d_raw.shape[1] // TILE_ROWS
    out = [torch.empty_like(t) for t in images]
    d_out = None if d_raw is None else torch.empty_like(d_raw)
    k = len(images)
    P = C.c_void_p * k
    with _OnDevice(images[0].device):
        st = _call("nvsr_compact_rows", lib.nvsr_compact_rows, P(*[t.data_ptr() for t in images]), P(*[t.data_ptr() for t in out]),
                   (C.c_int32 * k)(*[t.shape[1] * 8 for t in images]), k, _ptr(d_raw), 0 if d_raw is None else d_raw.stride(0),
                   _ptr(d_out), 0 if d_out is None else d_out.stride(0), _ptr(ids), _ptr(count), tiles, _stream(),
                   rows=tiles * TILE_ROWS)
    _lib.check(st, "nvsr_compact_rows")
    return out, d_out


def ray_sum_rows(img, ids, count, n_rays, n_samples, inv_scale=1.0, out=None):
    """per-ray sums of a LIST-ordered 128-channel image -> [n_rays, 128] fp32 (accumulated into `out` when given)."""
    lib = _lib.load()
    if out is None:
        out = torch.zeros((n_rays, 128), dtype=torch.float32, device=img.device)
    with _OnDevice(img.device):
        st = _call("nvsr_ray_sum_rows", lib.nvsr_ray_sum_rows, _ptr(img), _ptr(ids), _ptr(count), ids.numel(), n_samples,
                   float(inv_scale), _ptr(out), _stream())
    _lib.check(st, "nvsr_ray_sum_rows")
    return out


def nsc_to_planar_blocked(x, n_rays, n_samples):
    """[N, S, 4] -> planar [4, stride] in the BLOCKED row order, padding rows 0 (inverse of raw_to_nsc)."""
    nb, ts = -(-n_rays // BLK_RAYS), -(-n_samples // BLK_SAMPLES)
    buf = torch.zeros((nb * BLK_RAYS, ts * BLK_SAMPLES, 4), dtype=torch.float32, device=x.device)
    buf[:n_rays, :n_samples] = x
    r = buf.reshape(nb, BLK_RAYS, ts, BLK_SAMPLES, 4).permute(4, 0, 2, 3, 1)     # [ch, block, sblock, s%16, ray%8]
    return r.reshape(4, nb * ts * TILE_ROWS).contiguous()


# ---------------------------------------------------------------------------------------------
def composite(raw, z, rd, n_samples, noise=None, white_background=False, mip=False, n_fine=0, u=None,
              want_weights=False, want_inds=False, want_samples=False, row_order=ROWS_RAY_MAJOR):
    """volume_render_radiance_field (+ sample_pdf and the sort-merge on the coarse pass).

    raw: planar [4, stride] (r,g,b,sigma).  Returns a dict with rgb/disp/acc/depth and, optionally,
    weights / inds / z_samples / z_merged."""
    lib = _lib.load()
    n = rd.shape[0]
    dev = rd.device
    c = _lib.Composite()
    c.n_rays, c.n_samples = n, n_samples
    c.raw, c.raw_stride, c.row_order = raw.data_ptr(), raw.stride(0), row_order
    z = _f32c(z)
    rd = _f32c(rd)
    c.z, c.rd = z.data_ptr(), rd.data_ptr()
    noise = None if noise is None else _f32c(noise)
    c.noise = 0 if noise is None else noise.data_ptr()
    c.white_bkgd, c.mip = int(bool(white_background)), int(bool(mip))
    out = {
        "rgb": torch.empty((n, 3), dtype=torch.float32, device=dev),
        "disp": torch.empty((n,), dtype=torch.float32, device=dev),
        "acc": torch.empty((n,), dtype=torch.float32, device=dev),
        "depth": torch.empty((n,), dtype=torch.float32, device=dev),
    }
    if want_weights:
        out["weights"] = torch.empty((n, n_samples), dtype=torch.float32, device=dev)
    c.rgb, c.disp, c.acc, c.depth = (out[k].data_ptr() for k in ("rgb", "disp", "acc", "depth"))
    c.weights = out["weights"].data_ptr() if want_weights else 0
    c.n_fine = int(n_fine)
    if n_fine > 0:
        u = _f32c(u)
        c.u = u.data_ptr()
        c.u_per_ray = int(u.dim() == 2)
        s1 = n_samples + (1 if mip else 0)
        out["z_merged"] = torch.empty((n, s1 + n_fine), dtype=torch.float32, device=dev)
        c.z_merged = out["z_merged"].data_ptr()
        if want_inds:
            out["inds"] = torch.empty((n, n_fine), dtype=torch.int64, device=dev)
            c.inds = out["inds"].data_ptr()
        if want_samples:
            out["z_samples"] = torch.empty((n, n_fine), dtype=torch.float32, device=dev)
            c.z_samples = out["z_samples"].data_ptr()
    with _OnDevice(dev):
        st = _call("nvsr_composite", lib.nvsr_composite, C.byref(c), _stream(), rows=n * n_samples,
                   # per sample 16 B raw + 4 B z; per ray 12 B rd + 24 B out; coarse pass: merged depths written
                   bytes=n * n_samples * 20 + n * 36 + (n * (n_samples + n_fine) * 4 if n_fine > 0 else 0))
    _lib.check(st, "nvsr_composite")
    return out


def _fill_layers(dst, layers):
    for i, ly in enumerate(layers):
        c = dst[i]
        c.w = ly.w.data_ptr()
        c.bias = 0 if ly.bias is None else ly.bias.data_ptr()
        c.row_bias = 0
        c.head_w = 0 if ly.head_w is None else ly.head_w.data_ptr()
        c.head_b = 0 if ly.head_b is None else ly.head_b.data_ptr()
        c.k, c.n_out, c.relu = ly.k, ly.n_out, int(bool(ly.relu))
        c.head_n = 0 if ly.head_w is None else ly.head_w.shape[0]
        c.head_ch = ly.head_ch


def decoder_struct(dec):
    """nvsr_decoder_t of a scene.PackedPlanesDecoder (the per-ray bias of rgb[0] is produced inside nvsr_render_rays)"""
    d = _lib.Decoder()
    _fill_layers(d.density, dec.density)
    d.n_density = len(dec.density)
    _fill_layers(d.rgb, dec.rgb)
    d.n_rgb = len(dec.rgb)
    vw = dec.view_w
    if vw.dtype != torch.float32 or vw.stride(1) != 1:
        raise _lib.NvsrError("decoder view weights must be fp32 with unit column stride")
    d.view_w, d.view_ldw, d.view_b = vw.data_ptr(), vw.stride(0), dec.view_b.data_ptr()
    return d


_workspaces = {}   # (device, bytes rounded up) -> uint8 tensor, reused by every later frame of the same shape


def _workspace(nbytes, device):
    key = (str(device), (nbytes + (1 << 20) - 1) >> 20)
    ws = _workspaces.get(key)
    if ws is None:
        for k in [k for k in _workspaces if k[0] == key[0] and k[1] < key[1]]:     # keep only the largest per device
            del _workspaces[k]
        big = [v for k, v in _workspaces.items() if k[0] == key[0] and k[1] >= key[1]]
        ws = big[0] if big else torch.empty((key[1] << 20,), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def render_rays(ro, rd, viewdirs, near, far, packed_c, dec_c, packed_f, dec_f, precision, n_coarse, n_fine, t_vals, u=None,
                t_rand=None, noise_c=None, noise_f=None, lindisp=False, white_background=False):
    """nvsr_render_rays: coarse -> fine for one batch of prepared rays with ONE host call out of a cached workspace
    (no per-frame allocation besides the eight result maps).  Returns (coarse dict, fine dict or None) with rgb / disp /
    acc / depth, bit-identical to issuing the stage calls one by one."""
    lib = _lib.load()
    n = ro.shape[0]
    dev = ro.device
    r = _lib.Render()
    r.precision, r.n_rays, r.n_coarse, r.n_fine = precision, n, int(n_coarse), int(n_fine)
    r.ro, r.rd, r.viewdirs = ro.data_ptr(), rd.data_ptr(), viewdirs.data_ptr()
    r.near_, r.far_, r.lindisp, r.white_bkgd = float(near), float(far), int(bool(lindisp)), int(bool(white_background))
    keep = [_f32c(t_vals)]
    r.t_vals = keep[0].data_ptr()
    for name, t in (("t_rand", t_rand), ("u", u), ("noise_c", noise_c), ("noise_f", noise_f)):
        if t is not None:
            t = _f32c(t)
            keep.append(t)
            setattr(r, name, t.data_ptr())
    r.u_per_ray = int(u is not None and u.dim() == 2)
    pc = packed_c.cstruct()
    pf = pc if (packed_f is None or packed_f is packed_c) else packed_f.cstruct()
    dc = decoder_struct(dec_c)
    df = dc if (dec_f is None or dec_f is dec_c) else decoder_struct(dec_f)
    r.planes_coarse, r.planes_fine = C.pointer(pc), C.pointer(pf)
    r.dec_coarse, r.dec_fine = C.pointer(dc), C.pointer(df)
    vp_c = packed_c.vplane
    vp_f = vp_c if packed_f is None else packed_f.vplane
    r.vplane_coarse, r.vplane_fine = vp_c.data_ptr(), vp_f.data_ptr()
    r.vrh, r.vrw = vp_c.shape[0], vp_c.shape[1]
    r.az_lo, r.az_rng, r.el_lo, r.el_rng = packed_c.view_lo_rng

    def maps():
        return {"rgb": torch.empty((n, 3), dtype=torch.float32, device=dev), "disp": torch.empty((n,), dtype=torch.float32, device=dev),
                "acc": torch.empty((n,), dtype=torch.float32, device=dev), "depth": torch.empty((n,), dtype=torch.float32, device=dev)}

    co = maps()
    r.rgb_c, r.disp_c, r.acc_c, r.depth_c = (co[k].data_ptr() for k in ("rgb", "disp", "acc", "depth"))
    fo = None
    if n_fine > 0:
        fo = maps()
        r.rgb_f, r.disp_f, r.acc_f, r.depth_f = (fo[k].data_ptr() for k in ("rgb", "disp", "acc", "depth"))
    need = lib.nvsr_workspace_bytes(C.byref(r))
    if need < 0:
        raise _lib.NvsrError("nvsr_workspace_bytes: invalid render request")
    ws = _workspace(int(need), dev)
    r.workspace, r.workspace_bytes = ws.data_ptr(), ws.numel()
    evals = n * (n_coarse + ((n_coarse + n_fine) if n_fine > 0 else 0))
    LAUNCHES["nvsr_render_rays(stage launches)"] = LAUNCHES.get("nvsr_render_rays(stage launches)", 0) + \
        (6 if n_fine == 0 else (11 if vp_f is vp_c else 12))
    with _OnDevice(dev):
        st = lib.nvsr_render_rays(C.byref(r), _stream())
    _lib.check(st, "nvsr_render_rays")
    return co, fo


def raw_to_planar(radiance_field):
    """[N,S,4] (reference layout, train_utils.py:57-60) -> planar [4, N*S]."""
    n, s, _ = radiance_field.shape
    return radiance_field.reshape(n * s, 4).t().contiguous()


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, mip_nerf=False, noise=None):
    """Drop-in for volume_rendering_utils.volume_render_radiance_field (:6-51).

    When radiance_field_noise_std > 0 the reference draws CPU randn (:32); pass the same draw as
    `noise` ([N,S], unscaled) for parity, otherwise it is drawn here with torch.randn on CPU."""
    rf = _f32c(radiance_field)
    n, s, _ = rf.shape
    nz = None
    if radiance_field_noise_std > 0.0:
        if noise is None:
            noise = torch.randn(rf[..., 3].shape)
        nz = (noise * radiance_field_noise_std).to(rf)
    o = composite(raw_to_planar(rf), depth_values, ray_directions, s, noise=nz, white_background=white_background,
                  mip=mip_nerf, want_weights=True)
    return o["rgb"], o["disp"], o["acc"], o["weights"], o["depth"]


def sample_pdf(bins, weights, num_samples, det=False, u=None, cdf=None, return_all=False):
    """Drop-in for nerf_helpers.sample_pdf_2 (:668-702), bound as sample_pdf at train_utils.py:4.

    `u` overrides the uniform draws (the reference draws them on the CPU); `cdf` bypasses the
    pdf/cdf construction (stage test: identical cdf,u => bit-exact indices)."""
    lib = _lib.load()
    bins = _f32c(bins)
    _require_cuda(bins, "bins")
    n, nb = bins.shape
    dev = bins.device
    if u is None:
        u = torch.linspace(0.0, 1.0, steps=num_samples) if det else torch.rand([n, num_samples])
    u = _f32c(u, dev)
    w = None if weights is None else _f32c(weights)
    cdf = None if cdf is None else _f32c(cdf)
    inds = torch.empty((n, num_samples), dtype=torch.int64, device=dev)
    samples = torch.empty((n, num_samples), dtype=torch.float32, device=dev)
    cdf_out = torch.empty((n, nb), dtype=torch.float32, device=dev)
    with _OnDevice(dev):
        st = _call("nvsr_sample_pdf", lib.nvsr_sample_pdf, _ptr(bins), _ptr(w), _ptr(cdf), n, nb, _ptr(u), int(u.dim() == 2), num_samples,
                                 _ptr(inds), _ptr(samples), _ptr(cdf_out), _stream())
    _lib.check(st, "nvsr_sample_pdf")
    if return_all:
        return samples, inds, cdf_out
    return samples


# ---------------------------------------------------------------------------------------------
def ipe(z_edges, ro, rd, radius, n_freqs, layout=FEAT_ROWMAJOR_F32, k_pad=None):
    """cast_rays + IntegratedPositionalEncoding (mip.py:9-43,154-199): z_edges [n,S+1] -> [n*S, 6*n_freqs]."""
    lib = _lib.load()
    z = _f32c(z_edges)
    n, s1 = z.shape
    S = s1 - 1
    dev = z.device
    if layout == FEAT_ROWMAJOR_F32:
        out = torch.empty((n * S, 6 * n_freqs), dtype=torch.float32, device=dev)
        k_pad = 0
    else:
        k_pad = k_pad or (6 * n_freqs + 15) // 16 * 16
        tiles = (n * S + TILE_ROWS - 1) // TILE_ROWS
        out = torch.empty((tiles, k_pad // 8, TILE_ROWS, 8), dtype=LAYOUT_DTYPE[layout], device=dev)
    with _OnDevice(dev):
        st = _call("nvsr_ipe", lib.nvsr_ipe, _ptr(z), _ptr(_f32c(ro)), _ptr(_f32c(rd)), n, S, float(radius), n_freqs, layout, k_pad,
                          _ptr(out), _stream())
    _lib.check(st, "nvsr_ipe")
    return out


def cast_rays(t_vals, origins, directions, radii, ray_shape=None):
    """Drop-in for mip.cast_rays (mip.py:9-18): t_vals [n,S+1] interval edges, origins / directions [n,3], radii a
    scalar or a per-ray tensor [n] / [n,1] (train_utils.py:21-24 passes a constant column) ->
    (means [n,S,3], covs [n,S,3]) — the diagonal-covariance conical-frustum Gaussians.  `ray_shape` is unused in the
    reference as well."""
    lib = _lib.load()
    z = _f32c(t_vals)
    _require_cuda(z, "t_vals")
    n, s1 = z.shape
    S = s1 - 1
    ro, rd = _f32c(origins, z.device), _f32c(directions, z.device)
    rad_t, rad = None, 0.0
    if torch.is_tensor(radii):
        rad_t = _f32c(radii.reshape(-1), z.device)
        if rad_t.numel() == 1:
            rad_t = rad_t.expand(n).contiguous()
        if rad_t.numel() != n:
            raise _lib.NvsrError("cast_rays: radii must be a scalar or one value per ray")
    else:
        rad = float(radii)
    means = torch.empty((n, S, 3), dtype=torch.float32, device=z.device)
    covs = torch.empty_like(means)
    with _OnDevice(z.device):
        st = _call("nvsr_cast_rays", lib.nvsr_cast_rays, _ptr(z), _ptr(ro), _ptr(rd), _ptr(rad_t), rad, n, S, _ptr(means),
                   _ptr(covs), _stream())
    _lib.check(st, "nvsr_cast_rays")
    return means, covs


def ipe_encode(means, covs, n_freqs):
    """IntegratedPositionalEncoding.forward((means, covs)) (mip.py:164-191): [..., 3] x2 -> [..., 6*n_freqs]."""
    lib = _lib.load()
    m, c = _f32c(means), _f32c(covs)
    _require_cuda(m, "means")
    if m.shape != c.shape or m.shape[-1] != 3:
        raise _lib.NvsrError("ipe_encode: means and covs must both be [..., 3]")
    rows = m.numel() // 3
    out = torch.empty(tuple(m.shape[:-1]) + (6 * n_freqs,), dtype=torch.float32, device=m.device)
    with _OnDevice(m.device):
        st = _call("nvsr_ipe_encode", lib.nvsr_ipe_encode, _ptr(m), _ptr(c), rows, int(n_freqs), _ptr(out), _stream())
    _lib.check(st, "nvsr_ipe_encode")
    return out


def dir_encoding(dirs, n_freqs, include_input=True):
    """positional_encoding(dirs, n_freqs, include_input) per ray (nerf_helpers.py:552-575)."""
    lib = _lib.load()
    d = _f32c(dirs)
    n = d.shape[0]
    out = torch.empty((n, (3 if include_input else 0) + 6 * n_freqs), dtype=torch.float32, device=d.device)
    with _OnDevice(d.device):
        st = _call("nvsr_dir_encoding", lib.nvsr_dir_encoding, _ptr(d), n, n_freqs, int(bool(include_input)), _ptr(out), _stream())
    _lib.check(st, "nvsr_dir_encoding")
    return out


def mip_radius(scene_id):
    """radii of train_utils.py:21-23: DS parsed from the scene id suffix `_DS<k>`."""
    import re
    m = re.search(r"(?<=_DS)(\d)+(?=$)", scene_id)
    if m is None:
        raise _lib.NvsrError(f"mip path needs a scene id ending in _DS<k>, got {scene_id!r}")
    return int(m.group(0)) * 0.00135 * 2 / math.sqrt(12.0)


# ---------------------------------------------------------------------------------------------
# backward of the memory-bound stages (include/nvsr.h "BACKWARD")
def sample_gather_bwd(ro, rd, z, packed, d_feat_p, d_feat_m, d_planes=None, rows=None):
    """Scatter-add of the row-major fp32 feature gradients into channels-last plane gradients [Rh,Rw,C] (x3).

    `packed` supplies the geometry (box, projections, plane sizes); its plane images are not read.  `d_planes`
    (list of 3 tensors) are accumulated into when given, else allocated zeroed.  Returns the list.
    rows=(row_ids, count): the feature gradients are LIST-ordered ([len(row_ids), .]), row i belongs to BLOCKED row
    row_ids[i] (nvsr_sample_gather_bwd_rows)."""
    lib = _lib.load()
    ro, rd, z = _f32c(ro), _f32c(rd), _f32c(z)
    _require_cuda(ro, "ray_origins")
    n, S = z.shape
    pl = packed.cstruct()
    Cc = packed.channels
    if d_planes is None:
        d_planes = [torch.zeros((pl.rh[d], pl.rw[d], Cc), dtype=torch.float32, device=ro.device) for d in range(3)]
    gp = None if d_feat_p is None else _f32c(d_feat_p)
    gm = None if d_feat_m is None else _f32c(d_feat_m)
    nrow = n * S if rows is None else rows[0].numel()
    if gp is not None and tuple(gp.shape) != (nrow, 3 * Cc) or gm is not None and tuple(gm.shape) != (nrow, Cc):
        raise _lib.NvsrError("sample_gather_bwd: feature gradients must be [rows, 3C] / [rows, C]")
    s = _lib.Sampler()
    s.n_rays, s.n_samples = n, S
    s.ro, s.rd, s.z_in = ro.data_ptr(), rd.data_ptr(), z.data_ptr()
    ptrs = (C.c_void_p * 3)(*[t.data_ptr() for t in d_planes])
    with _OnDevice(ro.device):
        if rows is None:
            st = _call("nvsr_sample_gather_bwd", lib.nvsr_sample_gather_bwd, C.byref(s), C.byref(pl), _ptr(gp), _ptr(gm), ptrs,
                       _stream(), rows=n * S, bytes=n * S * (4 * Cc * 4 + 4))
        else:
            st = _call("nvsr_sample_gather_bwd", lib.nvsr_sample_gather_bwd_rows, C.byref(s), C.byref(pl), _ptr(gp), _ptr(gm),
                       _ptr(rows[0]), _ptr(rows[1]), nrow, ptrs, _stream(), rows=nrow)
    _lib.check(st, "nvsr_sample_gather_bwd")
    return d_planes


def viewdir_gather_bwd(viewdirs, packed, d_vfeat, d_vplane=None):
    """Scatter-add of the per-ray view-feature gradient [n,C] into the channels-last view-plane gradient."""
    lib = _lib.load()
    vd, g = _f32c(viewdirs), _f32c(d_vfeat)
    _require_cuda(vd, "viewdirs")
    n = vd.shape[0]
    rh, rw, Cc = packed.vplane.shape
    if d_vplane is None:
        d_vplane = torch.zeros((rh, rw, Cc), dtype=torch.float32, device=vd.device)
    az_lo, az_rng, el_lo, el_rng = packed.view_lo_rng
    with _OnDevice(vd.device):
        st = _call("nvsr_viewdir_gather_bwd", lib.nvsr_viewdir_gather_bwd, _ptr(vd), n, rh, rw, Cc, az_lo, az_rng, el_lo,
                   el_rng, _ptr(g), _ptr(d_vplane), _stream())
    _lib.check(st, "nvsr_viewdir_gather_bwd")
    return d_vplane


def composite_bwd(radiance_field, depth_values, ray_directions, d_rgb, d_acc=None, d_depth=None, d_weights=None,
                  noise=None, white_background=False, mip=False):
    """d radiance_field [N,S,4] of volume_render_radiance_field (volume_rendering_utils.py:15-51) from the upstream
    gradients of rgb_map (required), acc_map, depth_map and weights.  `noise`: [N,S] ALREADY scaled by the std."""
    lib = _lib.load()
    rf, z, rd = _f32c(radiance_field), _f32c(depth_values), _f32c(ray_directions)
    _require_cuda(rf, "radiance_field")
    n, S, _ = rf.shape
    if z.shape[1] != S + (1 if mip else 0):
        raise _lib.NvsrError("composite_bwd: depth_values must be [N,S] ([N,S+1] interval edges for mip)")
    out = torch.empty_like(rf)
    args = [None if t is None else _f32c(t) for t in (noise, d_rgb, d_acc, d_depth, d_weights)]
    with _OnDevice(rf.device):
        st = _call("nvsr_composite_bwd", lib.nvsr_composite_bwd, _ptr(rf), _ptr(z), _ptr(rd), _ptr(args[0]), n, S,
                   int(bool(white_background)), int(bool(mip)), _ptr(args[1]), _ptr(args[2]), _ptr(args[3]), _ptr(args[4]),
                   _ptr(out), _stream(), rows=n * S, bytes=n * S * 36 + n * 24)
    _lib.check(st, "nvsr_composite_bwd")
    return out
