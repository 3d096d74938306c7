"""bench.py's control flow and JSON contract, exercised on the CPU with the GPU pieces replaced by stand-ins
(host logic only: the numbers are fake, the keys, the headline/companion modes and the launch accounting are real).
"""
import json
import os
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
    pkg.render = types.SimpleNamespace(_state={"ray_chunk": 327680, "sparse_rgb": True}, clear_caches=lambda: None)
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

    pkg.render_frame = lambda H, W, focal, pose, mc, mf, opt, sid, scfg, row_range=None, **kw: frame((row_range[1] - row_range[0]) * W)
    pkg.run_one_iter_of_nerf = lambda H, W, focal, mc, mf, batch, opt, sid, mode, scene_config=None, **kw: frame(batch.shape[1])
    pkg.get_ray_bundle = lambda H, W, focal, pose, row_range=None: (torch.zeros(row_range[1] - row_range[0], W, 3),) * 2
    return pkg


def _patch_bench(setattr_, setitem, state, argv, elapsed_ms=10.0):
    """Replace the GPU pieces bench.py touches (package, events, device, pinned memory, clock sampler, CPU baseline)."""
    pkg = _fake_package(state)
    setitem(sys.modules, "nvsr_b200", pkg)
    setitem(sys.modules, "nvsr_b200.ops", pkg.ops)
    import importlib
    real_sharding = importlib.import_module("neural-volume-super-resolution_b200.sharding")
    setitem(sys.modules, "nvsr_b200.sharding", real_sharding)
    pkg.sharding = real_sharding

    class Ev(_Event):
        def elapsed_time(self, other):
            return elapsed_ms

    class Proxy:
        """the real module with a few names replaced — only bench.py sees it, torch itself stays untouched"""

        def __init__(self, real, **over):
            self.__dict__.update(_real=real, _over=over)

        def __getattr__(self, k):
            return self._over[k] if k in self._over else getattr(self._real, k)

    cuda = Proxy(torch.cuda, Event=Ev, synchronize=lambda *a, **k: None, set_device=lambda *a, **k: None)
    real_empty = torch.empty
    small = lambda *a, **k: (real_empty(1024, dtype=k.get("dtype")) if a and isinstance(a[0], int) and a[0] > 1 << 20
                             else real_empty(*a, **k))
    setattr_(bench, "torch", Proxy(torch, cuda=cuda, device=lambda *a, **k: torch.device("cpu"), empty=small))
    setattr_(bench, "ClockSampler", _Clock)
    def fake_workload(name, dev):
        spec = dict(bench.WORKLOADS[name], H=16, W=16)
        w = bench.Workload(name, **spec)
        w.mc = w.mf = w.opt = w.scfg = None
        w.sid, w.pose, w.focal = "sid", torch.eye(4), 1111.0
        if name == "cfg5":
            w.sids = ["s%d" % i for i in range(8)]
        return w

    setattr_(bench, "build_workload", fake_workload)
    setattr_(bench, "time_cpu_oracle", lambda **k: {"rays_per_s": 2000.0, "evals_per_s": 512000.0, "ms_per_step": 5.0, "cores": 8,
                                                    "cpu_model": "fake", "sample": "fake", "rays": 100})
    setattr_(bench, "time_torch_gpu_frame", lambda w, dev, **k: {"ms_per_step": 2000.0, "value": 128.0, "unit": "rays/s"})
    setattr_(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    setattr_(sys, "argv", ["bench.py"] + argv)
    return pkg


@pytest.mark.parametrize("flags,headline_sparse", [([], False), (["--sparse"], True), (["--precision", "fp32"], False)])
def test_bench_line_contract(monkeypatch, capfd, flags, headline_sparse):
    state = {"sparse": None, "frames": [], "precision": None}
    _patch_bench(monkeypatch.setattr, monkeypatch.setitem, state, ["--steps", "4", "--warmup", "3"] + flags)
    monkeypatch.setenv("NVSR_BENCH_WATCHDOG_S", "600")
    bench.main()
    out = [l for l in capfd.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "kernels", "cpu_baseline",
              "configs", "precision_modes", "torch_gpu_baseline"):
        assert k in d, k
    assert sorted(d["configs"]) == ["cfg1", "cfg2", "cfg3a", "cfg3b", "cfg4", "cfg5"]
    assert all(c["value"] > 0 and c["ms_per_step"] > 0 for c in d["configs"].values())
    assert d["configs"]["cfg5"]["frames_timed"] == 8 and d["configs"]["cfg2"]["note"] == "the headline of this line"
    assert d["cpu_baseline"]["cfg1"]["value"] > 0
    assert d["config"]["sparse_rgb"] is headline_sparse and d["steps"] == 4 and d["n_gpus"] == 1
    assert d["per_step"] == {"median_ms": 10.0, "min_ms": 10.0, "max_ms": 10.0, "rank": 0}
    assert d["ms_per_step"] == pytest.approx(10.0 / 4)
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    fp32 = "fp32" in flags
    assert (d["precision_modes"] is None) == fp32 and (d["torch_gpu_baseline"] is None) == fp32
    if not fp32:
        assert d["precision_modes"]["fp32"]["value"] > 0 and state["precision"] == "fp16"
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


def _rank_main(rank, world, port, out_dir):
    """One rank of a 2-rank bench run on the CPU: gloo instead of nccl, per-rank event times that DIFFER (so any
    decision taken from a rank's own clock would make the ranks issue different numbers of collectives and hang)."""
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), NVSR_BENCH_WATCHDOG_S="120")
    torch.set_num_threads(1)
    state = {"sparse": None, "frames": [], "precision": None}
    setitem = lambda d, k, v: d.__setitem__(k, v)
    _patch_bench(setattr, setitem, state, ["--gpus", str(world), "--steps", "3", "--warmup", "3"], elapsed_ms=4.0 + 3.0 * rank)
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, device_id=None, **k: real_init("gloo", rank=rank, world_size=world)
    fd = os.open(os.path.join(out_dir, "rank%d.out" % rank), os.O_WRONLY | os.O_CREAT)
    os.dup2(fd, 1)
    bench.main()
    with open(os.path.join(out_dir, "rank%d.frames" % rank), "w") as f:
        f.write(str(len(state["frames"])))


def test_two_rank_bench_flow_issues_the_same_collectives_on_every_rank(tmp_path):
    """bench.py under torchrun, world size 2, on the CPU (gloo): the short timed region forces the clock-sampling
    fallback, whose frame count must come from the all-reduced step time — with a rank-local count the ranks would issue
    different numbers of all_gathers and this test would die on the watchdog instead of finishing."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    lines = [l for l in open(tmp_path / "rank0.out").read().splitlines() if l.startswith("{")]
    assert len(lines) == 1 and not [l for l in open(tmp_path / "rank1.out").read().splitlines() if l.startswith("{")]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["config"]["sharding"] == "2 row bands" and d["cpu_baseline"] is None
    assert d["ms_per_step"] == pytest.approx(7.0 / 3)          # MAX over ranks of the per-rank event time / steps
    assert "timed region shorter" in d["clocks"].get("note", "")   # the fallback path ran
    assert open(tmp_path / "rank0.frames").read() == open(tmp_path / "rank1.frames").read()


def test_reference_arm_line_contract(monkeypatch, capfd):
    """`bench.py --impl reference`: rank 0 prints one line with impl / cpu_baseline / zero-byte e2e on the b200 arm's
    metric and unit; other ranks print nothing and return."""
    monkeypatch.setattr(bench, "time_cpu_oracle", lambda **k: {"rays_per_s": 2345.0, "evals_per_s": 2345.0 * 256, "ms_per_step": 982.0,
                                                               "rays": 2304, "cores": 24, "sample": "fake lattice",
                                                               "workload": bench.WORKLOADS[k.get("config", "cfg2")]["workload"],
                                                               "metric": "rays/s (800x800 render, 64 coarse + 128 fine samples/ray)"})
    for rank, expect in ((0, 1), (1, 0)):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "2")
        monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"])
        bench.main()
        out = [l for l in capfd.readouterr().out.splitlines() if l.startswith("{")]
        assert len(out) == expect
        if expect:
            d = json.loads(out[0])
            assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] == 2345.0 and d["higher_is_better"] is True
            assert d["metric"].startswith("rays/s (800x800 render") and d["config"]["workload"] == "cfg2_800x800_64+128_planes200"
            assert d["cpu_baseline"] == {"value": 2345.0, "unit": "rays/s", "cores": 24, "kind": "port", "sample": "fake lattice"}
            assert d["e2e"] == {"value": 2345.0, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
            assert d["n_gpus"] == 2 and d["steps"] == 2 and d["gpu_launches"] == 0
