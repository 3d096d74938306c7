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
    assert lib.nvsr_abi_version() == 2
    assert lib.nvsr_status_string(0) == b"ok"
    assert b"invalid" in lib.nvsr_status_string(-1)


def test_struct_sizes_match_c_layout(tmp_path):
    """sizeof/offsetof of the ABI structs as gcc lays them out from include/nvsr.h == the ctypes mirror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("nvsr_layer_t", _lib.Layer, "head_ch"), ("nvsr_planes_t", _lib.Planes, "proj"),
             ("nvsr_sampler_t", _lib.Sampler, "z_in"), ("nvsr_mlp_t", _lib.Mlp, "row_order"),
             ("nvsr_composite_t", _lib.Composite, "z_merged")]
    src = "#include <stdio.h>\n#include <stddef.h>\n#include \"nvsr.h\"\nint main(void){\n"
    for cname, _, last in pairs:
        src += f'printf("%zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));\n'
    src += "return 0;}\n"
    c = tmp_path / "sz.c"
    c.write_text(src)
    exe = tmp_path / "sz"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for (cname, ct, last), line in zip(pairs, out):
        size, off = (int(v) for v in line.split())
        field = {"in": "in_"}.get(last, last)
        assert ctypes.sizeof(ct) == size, (cname, ctypes.sizeof(ct), size)
        assert getattr(ct, field).offset == off, (cname, last)


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
