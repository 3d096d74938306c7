"""nvsr_b200 — B200 (sm_100a) ray-rendering hot path of Neural-Volume-Super-Resolution.

The directory is named after the reference (`neural-volume-super-resolution_b200`); because of the
hyphens import it as `nvsr_b200` (the alias module at the repo root) or through importlib.

Public surface = the reference's own call surface for this path:
    get_ray_bundle, run_one_iter_of_nerf, eval_nerf, volume_render_radiance_field, sample_pdf,
    install(train_utils)   and the stage-level ops in `ops`.
Host code is Python; every kernel is hand-written CUDA behind the C-ABI of include/nvsr.h.
There is no CPU fallback: without libnvsr_b200.so (or without a GPU) every entry point raises.
"""
from . import _lib, autograd, build, frames, ops, plane_store, render, scene, sharding  # noqa: F401
from ._lib import NVSR_BF16, NVSR_F16, NVSR_F32, NvsrError  # noqa: F401
from .ops import (get_ray_bundle, sample_pdf, volume_render_radiance_field)  # noqa: F401
from .render import (eval_nerf, get_precision, install, render_frame, run_one_iter_of_nerf, set_precision,  # noqa: F401
                     set_ray_chunk, set_sparse_rgb, uninstall)


from .ops import cast_rays  # noqa: F401,E402  (mip.cast_rays, mip.py:9-18)
from .render import planes_model_forward  # noqa: F401,E402  (TwoDimPlanesModel.forward(x[n,6]) -> [n,4], models.py:381-421)


class IntegratedPositionalEncoding:
    """Drop-in for mip.IntegratedPositionalEncoding (mip.py:154-199): same constructor, and calling it with the
    reference's `(means, covs)` tuple returns the `[..., 6*(multires-1)]` encoding (`nvsr_ipe_encode`).  Inside
    `run_one_iter_of_nerf` the encoding is fused with cast_rays into one kernel (`ops.ipe`); the render path only reads
    `.max_freq` of whatever object it is given as `encode_position_fn` (this one or the reference's own module)."""

    def __init__(self, input_dims=3, multires=10, include_input=False):
        self.out_dims = input_dims * 2 * (multires - 1)
        self.max_freq = multires - 1

    def __call__(self, x_coord):
        means, covs = x_coord
        return ops.ipe_encode(means, covs, self.max_freq)

    forward = __call__
