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
    assert lib.nvsr_abi_version() == 4
    assert lib.nvsr_status_string(0) == b"ok"
    assert b"invalid" in lib.nvsr_status_string(-1)


def test_struct_sizes_match_c_layout(tmp_path):
    """sizeof/offsetof of the ABI structs as gcc lays them out from include/nvsr.h == the ctypes mirror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("nvsr_layer_t", _lib.Layer, "head_ch"), ("nvsr_planes_t", _lib.Planes, "combine"),
             ("nvsr_sampler_t", _lib.Sampler, "z_in"), ("nvsr_mlp_t", _lib.Mlp, "row_order"),
             ("nvsr_composite_t", _lib.Composite, "z_merged"), ("nvsr_decoder_t", _lib.Decoder, "view_b"),
             ("nvsr_render_t", _lib.Render, "workspace_bytes"), ("nvsr_dgrad_t", _lib.Dgrad, "acts_listed")]
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


def test_concurrent_builds_are_serialised(tmp_path, monkeypatch):
    """Several processes asking for a stale library (torchrun ranks on a box whose mtimes were disturbed) must build it
    once at a time and publish it atomically: with a stand-in compiler, 4 racing processes produce exactly one build
    after the first, and the library file is never observed half-written."""
    import subprocess
    import sys
    import textwrap
    fake = tmp_path / "fake_nvcc.py"
    fake.write_text(textwrap.dedent("""\
        #!%s
        import sys, time, os
        out = sys.argv[sys.argv.index('-o') + 1]
        with open(os.path.join(os.path.dirname(out), 'build_count'), 'a') as f:
            f.write('x')
        with open(out, 'w') as f:
            f.write('half'); f.flush(); time.sleep(0.5); f.write('-whole')
        """ % sys.executable))
    fake.chmod(0o755)
    work = tmp_path / "pkg"
    (work / "csrc").mkdir(parents=True)
    (work / "csrc" / "a.cu").write_text("// source\n")
    (tmp_path / "include").mkdir()
    (tmp_path / "include" / "nvsr.h").write_text("// header\n")
    script = textwrap.dedent("""\
        import importlib.util, os, sys
        spec = importlib.util.spec_from_file_location('b', %r)
        b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
        b.PKG_DIR, b.REPO_DIR, b.CSRC = %r, %r, %r
        b.LIB_PATH = os.path.join(b.PKG_DIR, 'libnvsr_b200.so'); b.LIB_OVERRIDE = None
        b.build_library()
        assert open(b.LIB_PATH).read() == 'half-whole'
        """ % (nvsr_b200.build.__file__, str(work), str(tmp_path), str(work / "csrc")))
    env = dict(os.environ, NVCC=str(fake))
    procs = [subprocess.Popen([sys.executable, "-c", script], env=env) for _ in range(4)]
    assert [p.wait(timeout=120) for p in procs] == [0, 0, 0, 0]
    assert (work / "build_count").read_text() == "x"          # built once; the three that waited found it fresh
    assert not [f for f in os.listdir(work) if ".tmp." in f]
