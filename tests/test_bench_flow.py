"""bench.py's control flow and JSON contract, exercised on the CPU with the GPU pieces replaced by stand-ins
(host logic only: the numbers are fake, the keys, the headline/companion modes and the launch accounting are real).
"""
import json
import sys
import types

import pytest
import torch

import bench


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 10.0

    def query(self):
        return True


class _Clock:
    def __init__(self, index):
        pass

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 7}


def _fake_package(state):
    ops = types.SimpleNamespace(LAUNCHES={}, PROFILE=None)
    pkg = types.ModuleType("nvsr_b200")
    pkg.ops = ops
    pkg.render = types.SimpleNamespace(_state={"ray_chunk": 327680, "sparse_rgb": True})
    pkg.set_precision = lambda p: state.__setitem__("precision", p)
    pkg.set_ray_chunk = lambda n: None
    pkg.set_sparse_rgb = lambda on: state.__setitem__("sparse", bool(on))

    def launch(name, **meta):
        ops.LAUNCHES[name] = ops.LAUNCHES.get(name, 0) + 1
        if ops.PROFILE is not None:
            ops.PROFILE.append((name, _Event(), _Event(), meta))

    def frame(n):
        state["frames"].append(state["sparse"])
        rows = n * 192
        launch("nvsr_sample_gather", rows=rows, bytes=rows * 100)
        launch("nvsr_mlp_chain", rows=rows, count=None, flops_per_row=110848, bytes_per_row=100, flops=rows * 110848, bytes=rows * 100)
        cnt = torch.tensor([rows // 5], dtype=torch.int32) if state["sparse"] else None
        if state["sparse"]:
            launch("nvsr_sample_gather_rows", count=cnt, bytes_per_row=288)
        launch("nvsr_mlp_chain", rows=rows, count=cnt, flops_per_row=148224, bytes_per_row=300, flops=rows * 148224, bytes=rows * 300)
        launch("nvsr_composite", rows=rows, bytes=rows * 20)
        z3, z1 = torch.zeros(n, 3), torch.zeros(n)
        return (z3, z1, z1, z3, z1, z1, None, None, None)

    pkg.render_frame = lambda H, W, focal, pose, mc, mf, opt, sid, scfg, row_range=None: frame((row_range[1] - row_range[0]) * W)
    pkg.run_one_iter_of_nerf = lambda H, W, focal, mc, mf, batch, opt, sid, mode, scene_config=None: frame(batch.shape[1])
    pkg.get_ray_bundle = lambda H, W, focal, pose, row_range=None: (torch.zeros(row_range[1] - row_range[0], W, 3),) * 2
    return pkg


@pytest.mark.parametrize("flags,headline_sparse", [([], False), (["--sparse"], True), (["--precision", "fp32"], False)])
def test_bench_line_contract(monkeypatch, capfd, flags, headline_sparse):
    state = {"sparse": None, "frames": [], "precision": None}
    pkg = _fake_package(state)
    monkeypatch.setitem(sys.modules, "nvsr_b200", pkg)
    monkeypatch.setitem(sys.modules, "nvsr_b200.ops", pkg.ops)
    import importlib
    real_sharding = importlib.import_module("neural-volume-super-resolution_b200.sharding")
    monkeypatch.setitem(sys.modules, "nvsr_b200.sharding", real_sharding)
    pkg.sharding = real_sharding
    monkeypatch.setattr(bench, "ClockSampler", _Clock)
    monkeypatch.setattr(bench, "build_scene", lambda dev: (None, None, "sid", torch.eye(4), 1111.0, None, None))
    monkeypatch.setattr(bench, "time_cpu_oracle", lambda **k: {"rays_per_s": 2000.0, "cores": 8, "sample": "fake"})
    monkeypatch.setattr(bench, "RES", 16)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch, "device", lambda *a, **k: "cpu")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"})
                        if a and isinstance(a[0], int) and a[0] > 1 << 20 else real_empty(*a, **k))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3"] + flags)
    monkeypatch.setenv("NVSR_BENCH_WATCHDOG_S", "600")
    bench.main()
    out = [l for l in capfd.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "kernels", "cpu_baseline"):
        assert k in d, k
    assert d["config"]["sparse_rgb"] is headline_sparse and d["steps"] == 4 and d["n_gpus"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    fp32 = "fp32" in flags
    comp = "dense" if headline_sparse else "sparse"
    assert comp in d and (d[comp] is None) == fp32
    if not fp32:
        assert d[comp]["value"] > 0 and d[comp]["e2e"]["value"] > 0
        assert d["roofline"]["unit"] == "TFLOP/s" and 0 < d["roofline"]["frac"]
        if headline_sparse:
            assert abs(d["config"]["rgb_rows_evaluated"] - 0.2) < 1e-4
        else:
            assert d["config"]["rgb_rows_evaluated"] == 1.0 and abs(d["sparse"]["rgb_rows_evaluated"] - 0.2) < 1e-4
    # the headline region ran in the headline mode, the library is left in it, and both modes were exercised
    assert state["sparse"] is (headline_sparse and not fp32)
    assert state["frames"][:7] == [headline_sparse and not fp32] * 7
    assert fp32 or (True in state["frames"] and False in state["frames"])
