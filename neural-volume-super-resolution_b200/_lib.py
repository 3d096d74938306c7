"""ctypes binding of libnvsr_b200.so — the C-ABI declared in include/nvsr.h.

There is NO fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes as C
import os

from . import build as _build

NVSR_F32, NVSR_BF16, NVSR_F16 = 0, 1, 2
FEAT_ROWMAJOR_F32, FEAT_TILE_BF16, FEAT_TILE_F16 = 0, 1, 2
ROWS_RAY_MAJOR, ROWS_BLOCKED = 0, 1
BLK_RAYS, BLK_SAMPLES = 8, 16
TILE_ROWS = 128
MAX_LAYERS = 8
MAX_SAMPLES = 1024

c_f = C.c_float
c_i32 = C.c_int32
c_i64 = C.c_int64
c_p = C.c_void_p


class Planes(C.Structure):
    _fields_ = [
        ("plane", c_p * 3),
        ("rh", c_i32 * 3),
        ("rw", c_i32 * 3),
        ("channels", c_i32),
        ("dtype", c_i32),
        ("box_lo", c_f * 3),
        ("box_rng", c_f * 3),
        ("proj", (c_f * 6) * 3),
        ("combine", c_i32),
    ]


class Sampler(C.Structure):
    _fields_ = [
        ("n_rays", c_i64),
        ("n_samples", c_i32),
        ("ro", c_p),
        ("rd", c_p),
        ("near_", c_f),
        ("far_", c_f),
        ("lindisp", c_i32),
        ("t_vals", c_p),
        ("t_rand", c_p),
        ("z_in", c_p),
    ]


class Layer(C.Structure):
    _fields_ = [
        ("w", c_p),
        ("bias", c_p),
        ("row_bias", c_p),
        ("head_w", c_p),
        ("head_b", c_p),
        ("k", c_i32),
        ("n_out", c_i32),
        ("relu", c_i32),
        ("head_n", c_i32),
        ("head_ch", c_i32),
    ]


class Mlp(C.Structure):
    _fields_ = [
        ("precision", c_i32),
        ("n_layers", c_i32),
        ("layer", Layer * MAX_LAYERS),
        ("in_", c_p),
        ("rows", c_i64),
        ("samples_per_ray", c_i32),
        ("n_rays", c_i64),
        ("raw", c_p),
        ("raw_stride", c_i64),
        ("row_order", c_i32),
        ("row_ids", c_p),
        ("row_count", c_p),
    ]


class Dgrad(C.Structure):
    _fields_ = [
        ("w", c_p * 4),
        ("k0", c_i32),
        ("head_w", c_p),
        ("head_n", c_i32),
        ("head_ch", c_i32),
        ("d_raw", c_p),
        ("raw_stride", c_i64),
        ("scale", c_f),
        ("act", c_p * 4),
        ("g", c_p * 4),
        ("dout_img", c_p),
        ("d_x0", c_p),
        ("n_rays", c_i64),
        ("n_samples", c_i32),
        ("row_count", c_p),
        ("row_ids", c_p),
        ("act_list", c_p * 4),
        ("x0_img", c_p),
        ("x0_list", c_p),
        ("acts_listed", c_i32),
    ]


class Composite(C.Structure):
    _fields_ = [
        ("n_rays", c_i64),
        ("n_samples", c_i32),
        ("raw", c_p),
        ("raw_stride", c_i64),
        ("row_order", c_i32),
        ("z", c_p),
        ("rd", c_p),
        ("noise", c_p),
        ("white_bkgd", c_i32),
        ("mip", c_i32),
        ("rgb", c_p),
        ("disp", c_p),
        ("acc", c_p),
        ("depth", c_p),
        ("weights", c_p),
        ("n_fine", c_i32),
        ("u", c_p),
        ("u_per_ray", c_i32),
        ("inds", c_p),
        ("z_samples", c_p),
        ("z_merged", c_p),
    ]


class Decoder(C.Structure):
    _fields_ = [
        ("density", Layer * MAX_LAYERS),
        ("n_density", c_i32),
        ("rgb", Layer * MAX_LAYERS),
        ("n_rgb", c_i32),
        ("view_w", c_p),
        ("view_ldw", c_i32),
        ("view_b", c_p),
    ]


class Render(C.Structure):
    _fields_ = [
        ("precision", c_i32),
        ("n_rays", c_i64),
        ("n_coarse", c_i32),
        ("n_fine", c_i32),
        ("ro", c_p),
        ("rd", c_p),
        ("viewdirs", c_p),
        ("near_", c_f),
        ("far_", c_f),
        ("lindisp", c_i32),
        ("white_bkgd", c_i32),
        ("t_vals", c_p),
        ("t_rand", c_p),
        ("u", c_p),
        ("u_per_ray", c_i32),
        ("noise_c", c_p),
        ("noise_f", c_p),
        ("planes_coarse", C.POINTER(Planes)),
        ("planes_fine", C.POINTER(Planes)),
        ("vplane_coarse", c_p),
        ("vplane_fine", c_p),
        ("vrh", c_i32),
        ("vrw", c_i32),
        ("az_lo", c_f),
        ("az_rng", c_f),
        ("el_lo", c_f),
        ("el_rng", c_f),
        ("dec_coarse", C.POINTER(Decoder)),
        ("dec_fine", C.POINTER(Decoder)),
        ("rgb_c", c_p),
        ("disp_c", c_p),
        ("acc_c", c_p),
        ("depth_c", c_p),
        ("rgb_f", c_p),
        ("disp_f", c_p),
        ("acc_f", c_p),
        ("depth_f", c_p),
        ("workspace", c_p),
        ("workspace_bytes", c_i64),
    ]


# name -> (restype, argtypes); must list every symbol include/nvsr.h declares
SIGNATURES = {
    "nvsr_abi_version": (c_i32, []),
    "nvsr_status_string": (C.c_char_p, [c_i32]),
    "nvsr_rows_padded": (c_i64, [c_i64, c_i32, c_i32]),
    "nvsr_ray_bundle": (c_i32, [c_i32, c_i32, c_f, c_f, C.POINTER(c_f), c_i32, c_f, c_i32, c_i32, c_p, c_p, c_p]),
    "nvsr_ray_bundle_dev": (c_i32, [c_i32, c_i32, c_f, c_f, c_p, c_i32, c_f, c_i32, c_i32, c_p, c_p, c_p]),
    "nvsr_prepare_rays": (c_i32, [c_p, c_p, c_i64, c_i32, c_i32, c_i32, C.c_double, C.c_double, c_p, c_p, c_p, c_p]),
    "nvsr_pack_plane": (c_i32, [c_p, c_i32, c_i32, c_i32, c_p, c_i32, c_p]),
    "nvsr_pack_weight16": (c_i32, [c_p, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p]),
    "nvsr_sample_gather": (c_i32, [C.POINTER(Sampler), C.POINTER(Planes), c_i32, c_p, c_p, c_p, c_p]),
    "nvsr_keep_rows": (c_i32, [c_p, c_p, c_i64, c_i32, c_p, c_p, c_p]),
    "nvsr_sample_gather_rows": (c_i32, [C.POINTER(Sampler), C.POINTER(Planes), c_i32, c_p, c_p, c_i64, c_p, c_p]),
    "nvsr_viewdir_gather": (c_i32, [c_p, c_i64, c_p, c_i32, c_i32, c_i32, c_f, c_f, c_f, c_f, c_p, c_p]),
    "nvsr_row_bias": (c_i32, [c_p, c_i64, c_i32, c_p, c_i32, c_p, c_i32, c_p, c_p]),
    "nvsr_mlp_chain": (c_i32, [C.POINTER(Mlp), c_p]),
    "nvsr_mlp_chain_split": (c_i32, [c_p, c_i32, C.POINTER(c_p), C.POINTER(c_p), C.POINTER(c_p), c_p, c_p, c_i32, c_i32, c_i64, c_i32,
                                     c_p, c_i64, c_p]),
    "nvsr_mlp_chain_train": (c_i32, [C.POINTER(Mlp), C.POINTER(c_p), c_p]),
    "nvsr_mlp_dgrad": (c_i32, [C.POINTER(Dgrad), c_p]),
    "nvsr_mlp_wgrad": (c_i32, [c_p, c_p, c_i32, c_i64, c_f, c_p, c_i64, c_p, c_p]),
    "nvsr_ray_sum": (c_i32, [c_p, c_i64, c_i32, c_f, c_p, c_p]),
    "nvsr_mlp_wgrad_chain": (c_i32, [C.POINTER(c_p), c_p, c_i32, C.POINTER(c_p), c_p, c_i64, c_f, C.POINTER(c_p), C.POINTER(c_i64),
                                     C.POINTER(c_p), c_p, c_p]),
    "nvsr_sample_gather_hilo": (c_i32, [C.POINTER(Sampler), C.POINTER(Planes), C.POINTER(c_p), c_p, c_p, c_p, c_p, c_p]),
    "nvsr_mlp_chain_split_tiled": (c_i32, [c_p, c_i32, C.POINTER(c_p), C.POINTER(c_p), C.POINTER(c_p), c_p, c_p, c_i32, c_i32, c_i64,
                                           c_i32, c_p, c_i64, c_p]),
    "nvsr_sort_cat": (c_i32, [c_p, c_i32, c_p, c_i32, c_i64, c_p, c_p]),
    "nvsr_nonzero_rows": (c_i32, [c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "nvsr_compact_rows": (c_i32, [C.POINTER(c_p), C.POINTER(c_p), C.POINTER(c_i32), c_i32, c_p, c_i64, c_p, c_i64, c_p, c_p,
                                  c_i64, c_p]),
    "nvsr_mlp_wgrad_chain_rows": (c_i32, [C.POINTER(c_p), c_p, c_i32, C.POINTER(c_p), c_p, c_i64, c_p, c_f, C.POINTER(c_p),
                                          C.POINTER(c_i64), C.POINTER(c_p), c_p, c_p]),
    "nvsr_ray_sum_rows": (c_i32, [c_p, c_p, c_p, c_i64, c_i32, c_f, c_p, c_p]),
    "nvsr_sample_gather_bwd_rows": (c_i32, [C.POINTER(Sampler), C.POINTER(Planes), c_p, c_p, c_p, c_p, c_i64, C.POINTER(c_p),
                                            c_p]),
    "nvsr_pack_weights16": (c_i32, [c_i32, C.POINTER(c_p), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32),
                                    C.POINTER(c_p), c_i32, c_p, c_p]),
    "nvsr_composite": (c_i32, [C.POINTER(Composite), c_p]),
    "nvsr_sample_pdf": (c_i32, [c_p, c_p, c_p, c_i64, c_i32, c_p, c_i32, c_i32, c_p, c_p, c_p, c_p]),
    "nvsr_ipe": (c_i32, [c_p, c_p, c_p, c_i64, c_i32, c_f, c_i32, c_i32, c_i32, c_p, c_p]),
    "nvsr_dir_encoding": (c_i32, [c_p, c_i64, c_i32, c_i32, c_p, c_p]),
    "nvsr_sr_finalize": (c_i32, [c_p, c_i32, c_i64, c_i64, c_i32, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p, c_p]),
    "nvsr_workspace_bytes": (c_i64, [C.POINTER(Render)]),
    "nvsr_render_rays": (c_i32, [C.POINTER(Render), c_p]),
    "nvsr_cast_rays": (c_i32, [c_p, c_p, c_p, c_p, c_f, c_i64, c_i32, c_p, c_p, c_p]),
    "nvsr_ipe_encode": (c_i32, [c_p, c_p, c_i64, c_i32, c_p, c_p]),
    "nvsr_sample_gather_bwd": (c_i32, [C.POINTER(Sampler), C.POINTER(Planes), c_p, c_p, C.POINTER(c_p), c_p]),
    "nvsr_viewdir_gather_bwd": (c_i32, [c_p, c_i64, c_i32, c_i32, c_i32, c_f, c_f, c_f, c_f, c_p, c_p, c_p]),
    "nvsr_frame_to_u8": (c_i32, [c_p, c_i64, c_p, c_p]),
    "nvsr_composite_bwd": (c_i32, [c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_p, c_p]),
}

_LIB = None


class NvsrError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load the shared library (building it in-tree first when absent/stale and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if build_if_missing:
        try:
            _build.build_library()
        except FileNotFoundError:
            pass  # no nvcc on this box: use the prebuilt .so that travelled with the tree
    if not os.path.exists(path):
        raise NvsrError(
            f"{path} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU / PyTorch fallback for the nvsr_b200 render path."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.nvsr_abi_version() != 4:
        raise NvsrError("libnvsr_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().nvsr_status_string(status).decode()
        raise NvsrError(f"{what} failed: status {status} ({msg})")
