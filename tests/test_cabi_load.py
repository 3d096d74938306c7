"""CPU: the C-ABI library builds/loads and exports every symbol include/nvsr.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

import nvsr_b200
from nvsr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nvsr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nvsr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 14
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.lib_path())
    for s in declared_symbols():
        assert hasattr(raw, s), s
    assert lib.nvsr_abi_version() == 1
    assert lib.nvsr_status_string(0) == b"ok"
    assert b"invalid" in lib.nvsr_status_string(-1)


def test_struct_sizes_match_c_layout():
    # sizeof() of the ABI structs as compiled by g++/nvcc on LP64 (checked with a C program)
    sizes = [ctypes.sizeof(x) for x in (_lib.Layer, _lib.Planes, _lib.Sampler, _lib.Mlp, _lib.Composite)]
    assert sizes == [64, 152, 72, 568, 152]


def test_no_cpu_fallback():
    import torch
    with pytest.raises(nvsr_b200.NvsrError):
        nvsr_b200.get_ray_bundle(4, 4, 10.0, torch.eye(4))  # CPU pose: must refuse, not fall back
    with torch.no_grad(), pytest.raises(nvsr_b200.NvsrError):
        from nvsr_b200 import scene
        mc, mf, sid = scene.make_synthetic_scene(plane_res=8, view_res=8)
        rays = torch.zeros(2, 4, 3)
        nvsr_b200.run_one_iter_of_nerf(2, 2, 3.0, mc, mf, rays, scene.render_options(8, 8), sid, "validation",
                                       scene_config=scene.scene_cfg())


def test_forward_only_guard():
    import torch
    from nvsr_b200 import scene
    mc, mf, sid = scene.make_synthetic_scene(plane_res=8, view_res=8)
    with pytest.raises(RuntimeError, match="forward-only"):
        nvsr_b200.run_one_iter_of_nerf(2, 2, 3.0, mc, mf, torch.zeros(2, 4, 3), scene.render_options(8, 8), sid,
                                       "validation", scene_config=scene.scene_cfg())
