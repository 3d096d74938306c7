"""Plane super-resolution inference, device-resident (`PlanesSR.forward`, models.py:884-926, with
the EDSR of models.py:792-822 as `inner_model`).

What the reference does per SR plane: run EDSR on the replicate-padded LR plane, crop, add the bilinearly up-sampled
LR plane, cache the result ON THE CPU (:925) and re-upload it for every network chunk (:893) — 123 MB of H2D traffic per
plane and chunk at 800^2 — after which `grid_sample` reads it channel-strided.  Here:
  * the conv chain runs once per plane version, channels-last, on the tensor cores through cuDNN — stock PyTorch
    plumbing, as BASELINE.json's north_star prescribes for the SR convolutions (fp16 / bf16 operands in the 16-bit
    precision modes, fp32 in the fp32 parity mode);
  * everything after it — crop, bilinear x`scale` up-sampling of the LR plane, the add — is ONE hand-written kernel
    (`nvsr_sr_finalize`, csrc/sr.cu) that writes the plane directly as the gather's packed image (16-bit x-pair records or
    fp32 channels-last): the SR plane is born in its final layout and never leaves the device;
  * the packed image is cached by (LR plane identity, version, conv weights) and seeds the render path's plane cache.
Unsupported PlanesSR options raise (input/output noise are training-only; region-of-interest inference is the training
path's crop and is not needed for full-plane rendering).
"""
import ctypes as C
import math
import weakref

import torch
import torch.nn.functional as F

from . import _lib, ops
from ._lib import NVSR_BF16, NVSR_F16, NVSR_F32

_DT = {NVSR_F32: torch.float32, NVSR_F16: torch.float16, NVSR_BF16: torch.bfloat16}
_CODE = {torch.float32: NVSR_F32, torch.float16: NVSR_F16, torch.bfloat16: NVSR_BF16}


def edsr_forward(net, x, compute_dtype=torch.float32):
    """EDSR.forward (models.py:817-822) with 'valid' convolutions (PlanesSR builds it with padding=0), channels-last.
    x: [1, C, H, W] already padded by `required_padding`.  Returns the network output [1, C_out, H', W'] (channels-last
    memory format) in `compute_dtype`."""
    cl = torch.channels_last

    def conv(m, t):
        if m.bias is not None or m.padding not in ((0, 0), 0) or m.stride != (1, 1):
            raise NotImplementedError("nvsr_b200.sr: EDSR convolutions are expected bias-free, unpadded, stride 1")
        return F.conv2d(t, m.weight.to(dtype=compute_dtype, memory_format=cl))

    t = x.to(dtype=compute_dtype).contiguous(memory_format=cl)
    out = conv(net.conv_input, t)
    for blk in net.residual:
        k = blk.conv1.kernel_size[0]
        m = 2 * (k // 2)                                  # the two valid convs eat 2*(k//2) pixels per side
        ident = out[..., m:out.shape[-2] - m, m:out.shape[-1] - m] if m else out
        y = torch.relu_(conv(blk.conv1, out))
        y = conv(blk.conv2, y)
        out = torch.add(ident, y, alpha=0.1)              # output *= 0.1; output += identity  (models.py:786-788)
    out = conv(net.conv_mid, out)
    for m in net.upscale:
        out = F.pixel_shuffle(out, 2) if isinstance(m, torch.nn.PixelShuffle) else conv(m, out)
    return conv(net.conv_output, out)


class PlaneSuperResolver:
    """Device-resident `PlanesSR.forward(plane_name)` for one SR model object (the reference's `models.PlanesSR` or the
    stand-in `scene.PlanesSRModel`): reads `inner_model`, `scale_factor`, `LR_planes`, `align_corners`, `HR_overpadding`,
    `inner_model.required_padding` and the optional `planes_mean_NON_LEARNED` / `planes_std_NON_LEARNED`."""

    def __init__(self, sr_model):
        self.sr = weakref.ref(sr_model)
        self.cache = {}      # (plane_name, packed_dtype) -> (signature, packed image)

    def _check(self, sr):
        if getattr(sr, "plane_interp", "bilinear") != "bilinear":
            raise NotImplementedError("nvsr_b200.sr: plane_interp must be 'bilinear'")
        if sr.training and (getattr(sr, "input_noise", 0) or getattr(sr, "output_noise", 0)):
            raise NotImplementedError("nvsr_b200.sr: SR input/output noise is a training-time option")
        if getattr(sr, "residual_planes", None):
            raise NotImplementedError("nvsr_b200.sr: pre-interpolated residual planes (save_interpolated) are not supported")

    def _signature(self, sr, lr):
        from . import scene
        return (id(lr), lr.data_ptr(), lr._version, scene._GENERATION[0]) + \
            tuple((p.data_ptr(), p._version) for p in sr.inner_model.parameters())

    @torch.no_grad()
    def super_resolve(self, plane_name, packed_dtype=None, want_nchw=False, compute_dtype=None):
        """-> (packed image or None, fp32 NCHW plane [1,C,RH,RW] or None).  packed_dtype NVSR_F32 | NVSR_F16 | NVSR_BF16
        selects the gather image written; the conv chain computes in `compute_dtype` (default: fp32 for an fp32 image,
        else the image's 16-bit type)."""
        lib = _lib.load()
        sr = self.sr()
        self._check(sr)
        lr = sr.LR_planes[plane_name]
        if not lr.is_cuda:
            lr = lr.cuda()
        lr = lr.detach()
        if lr.dim() != 4 or lr.shape[0] != 1:
            raise _lib.NvsrError("LR plane must be [1, C, R, R]")
        if compute_dtype is None:
            compute_dtype = torch.float32 if packed_dtype in (None, NVSR_F32) else _DT[packed_dtype]
        key = (plane_name, packed_dtype, want_nchw, compute_dtype)
        sig = self._signature(sr, sr.LR_planes[plane_name])
        hit = self.cache.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        _, c, rh, rw = lr.shape
        s = int(sr.scale_factor)
        x = lr.float()
        if hasattr(sr, "planes_mean_NON_LEARNED"):
            x = (x - sr.planes_mean_NON_LEARNED.to(x)) / sr.planes_std_NON_LEARNED.to(x)      # models.py:899-901
        pad = int(sr.inner_model.required_padding)
        x = F.pad(x, (pad, pad, pad, pad), mode="replicate")                                  # models.py:910-912
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False       # fp32 mode means fp32 arithmetic (the 1e-3 parity contract)
        try:
            diff = edsr_forward(sr.inner_model, x, compute_dtype)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        crop = int(sr.HR_overpadding)
        if diff.shape[1] != c or diff.shape[2] != rh * s + 2 * crop or diff.shape[3] != rw * s + 2 * crop:
            raise _lib.NvsrError(f"EDSR output {tuple(diff.shape)} does not match scale {s} / over-padding {crop}")
        d = diff.permute(0, 2, 3, 1)                  # [1, H, W, C] view of the channels-last result
        if d.stride(3) != 1:
            d = d.contiguous()
        lr32 = lr.float().contiguous()
        packed = nchw = None
        if packed_dtype is not None:
            if packed_dtype == NVSR_F32:
                packed = torch.empty((rh * s, rw * s, c), dtype=torch.float32, device=lr.device)
            else:
                packed = torch.empty((rh * s, c // 8, rw * s, 2, 8), dtype=_DT[packed_dtype], device=lr.device)
        if want_nchw:
            nchw = torch.empty((1, c, rh * s, rw * s), dtype=torch.float32, device=lr.device)
        with torch.cuda.device(lr.device):
            st = ops._call("nvsr_sr_finalize", lib.nvsr_sr_finalize, ops._ptr(d), _CODE[d.dtype], d.stride(1), d.stride(2), crop,
                           ops._ptr(lr32), c, rh, rw, s, int(bool(getattr(sr, "align_corners", True))), ops._ptr(packed),
                           NVSR_F32 if packed_dtype is None else packed_dtype, ops._ptr(nchw), ops._stream())
        _lib.check(st, "nvsr_sr_finalize")
        if packed_dtype == NVSR_F16 and packed is not None and float(packed.abs().max()) >= ops.F16_MAX:
            # the kernel saturates instead of producing inf: refuse a plane that did (once per plane, not per frame)
            raise _lib.NvsrError("super-resolved plane exceeds the fp16 range; use set_precision('bf16') or 'fp32'")
        self.cache[key] = (sig, (packed, nchw))
        return packed, nchw


_resolvers = weakref.WeakKeyDictionary()


def resolver_of(sr_model):
    r = _resolvers.get(sr_model)
    if r is None:
        r = _resolvers[sr_model] = PlaneSuperResolver(sr_model)
    return r


def is_sr_model(obj):
    """an SR model this module can run: the reference's PlanesSR (or the stand-in) around an EDSR"""
    inner = getattr(obj, "inner_model", None)
    return inner is not None and all(hasattr(inner, a) for a in ("conv_input", "residual", "conv_mid", "upscale", "conv_output")) \
        and hasattr(obj, "LR_planes") and hasattr(obj, "scale_factor")
