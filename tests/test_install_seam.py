"""The drop-in seam: `install()` rebinds `train_utils.run_one_iter_of_nerf` — which the reference's
`eval_nerf` resolves through its module globals at call time (train_utils.py:311) — and optionally
`nerf_helpers.get_ray_bundle`; the reference's own functions keep running under autograd (train(), train_nerf.py:860)
and `uninstall()` restores them.  Host logic only: no kernel is launched here."""
import types

import torch

import nvsr_b200
from nvsr_b200 import render


def _fake_reference_modules():
    tu = types.ModuleType("train_utils")
    nh = types.ModuleType("nerf_helpers")
    calls = []

    def run_one_iter_of_nerf(*a, **k):
        calls.append(("ref_run", a, k))
        return ("ref",) * 9

    def get_ray_bundle(h, w, f, pose, padding_size=0, downsampling_offset=0):
        calls.append(("ref_grb", h, w))
        return "ref_ro", "ref_rd"

    # like train_utils.eval_nerf: the callee is looked up in the module's globals when eval_nerf RUNS
    exec("def eval_nerf(*a, **k):\n    return run_one_iter_of_nerf(*a, **k)\n", tu.__dict__)
    tu.run_one_iter_of_nerf = run_one_iter_of_nerf
    nh.get_ray_bundle = get_ray_bundle
    return tu, nh, calls


def test_install_dispatch_and_uninstall(monkeypatch):
    tu, nh, calls = _fake_reference_modules()
    ref_run, ref_grb = tu.run_one_iter_of_nerf, nh.get_ray_bundle
    seen = []
    monkeypatch.setattr(render, "run_one_iter_of_nerf", lambda *a, **k: seen.append((a, k)) or ("b200",) * 9)
    nvsr_b200.install(tu, nh)
    assert tu.run_one_iter_of_nerf is not ref_run and nh.get_ray_bundle is not ref_grb
    wrapped = tu.run_one_iter_of_nerf
    nvsr_b200.install(tu, nh)                       # idempotent: no double wrapping
    assert tu.run_one_iter_of_nerf is wrapped

    with torch.enable_grad():                       # train(): the reference's autograd path is untouched
        assert tu.eval_nerf(1, 2, mode="train")[0] == "ref"
    assert calls[-1][0] == "ref_run" and calls[-1][1] == (1, 2) and calls[-1][2] == {"mode": "train"}
    with torch.no_grad():                           # evaluate(): through eval_nerf's late-bound global
        assert tu.eval_nerf(3, 4, mode="validation")[0] == "b200"
    assert seen == [((3, 4), {"mode": "validation"})]

    # get_ray_bundle: host poses keep the reference's function, device poses take the kernel
    assert nh.get_ray_bundle(4, 4, 1.0, torch.eye(4)) == ("ref_ro", "ref_rd")
    assert calls[-1] == ("ref_grb", 4, 4)

    nvsr_b200.uninstall(tu)
    assert tu.run_one_iter_of_nerf is ref_run
    with torch.no_grad():
        assert tu.eval_nerf(5)[0] == "ref"


def test_no_cpu_fallback():
    """The product path has no CPU / PyTorch fallback: host tensors are refused, loudly."""
    import pytest
    with pytest.raises((nvsr_b200.NvsrError, RuntimeError, ValueError, AssertionError)):
        nvsr_b200.get_ray_bundle(4, 4, 1.0, torch.eye(4))


def test_install_training_seam(monkeypatch):
    """train() resolves `run_one_iter_of_nerf` in train_nerf's OWN globals (train_nerf.py:13,860): passing that module
    rebinds it too; grad-enabled calls stay on the reference unless `differentiable=True`, which routes them to
    nvsr_b200.autograd (no silent fallback either way)."""
    from nvsr_b200 import autograd
    for differentiable in (False, True):
        tu, nh, calls = _fake_reference_modules()
        tn = types.ModuleType("train_nerf")
        tn.run_one_iter_of_nerf = tu.run_one_iter_of_nerf            # `from train_utils import run_one_iter_of_nerf`
        exec("def train(*a, **k):\n    return run_one_iter_of_nerf(*a, **k)\n", tn.__dict__)
        ref_run = tu.run_one_iter_of_nerf
        monkeypatch.setattr(render, "run_one_iter_of_nerf", lambda *a, **k: ("b200",) * 9)
        monkeypatch.setattr(autograd, "run_one_iter_of_nerf", lambda *a, **k: ("b200_autograd",) * 9)
        nvsr_b200.install(tu, nh, train_nerf_module=tn, differentiable=differentiable)
        assert tn.run_one_iter_of_nerf is tu.run_one_iter_of_nerf is not ref_run
        with torch.enable_grad():
            assert tn.train(1, mode="train")[0] == ("b200_autograd" if differentiable else "ref")
        with torch.no_grad():
            assert tn.train(1, mode="validation")[0] == "b200"
        nvsr_b200.uninstall(tu)
        nvsr_b200.uninstall(tn)
        assert tu.run_one_iter_of_nerf is ref_run and tn.run_one_iter_of_nerf is ref_run


def test_reinstall_updates_options_and_uninstall_restores_everything(monkeypatch):
    """A second install() is not a silent no-op: it updates the wrapper's options and rebinds the modules it names;
    uninstall() restores every name any install() call touched."""
    from nvsr_b200 import autograd
    tu, nh, calls = _fake_reference_modules()
    tn = types.ModuleType("train_nerf")
    tn.run_one_iter_of_nerf = tu.run_one_iter_of_nerf
    ref_run, ref_grb = tu.run_one_iter_of_nerf, nh.get_ray_bundle
    monkeypatch.setattr(autograd, "run_one_iter_of_nerf", lambda *a, **k: ("b200_autograd",) * 9)
    nvsr_b200.install(tu)                                           # plain install first
    with torch.enable_grad():
        assert tu.run_one_iter_of_nerf(1)[0] == "ref"
    assert nh.get_ray_bundle is ref_grb and tn.run_one_iter_of_nerf is ref_run
    nvsr_b200.install(tu, nh, train_nerf_module=tn, differentiable=True)   # second call: takes effect
    assert tn.run_one_iter_of_nerf is tu.run_one_iter_of_nerf is not ref_run
    assert nh.get_ray_bundle is not ref_grb
    with torch.enable_grad():
        assert tn.run_one_iter_of_nerf(1)[0] == "b200_autograd"
    nvsr_b200.uninstall(tu)
    assert tu.run_one_iter_of_nerf is ref_run and tn.run_one_iter_of_nerf is ref_run and nh.get_ray_bundle is ref_grb


def test_install_sets_and_logs_precision(caplog):
    import logging
    tu, nh, _ = _fake_reference_modules()
    before = nvsr_b200.get_precision()
    with caplog.at_level(logging.INFO, logger="nvsr_b200"):
        nvsr_b200.install(tu, precision="fp32")
    assert nvsr_b200.get_precision() == "fp32"
    assert any("precision=fp32" in r.getMessage() for r in caplog.records)
    nvsr_b200.uninstall(tu)
    nvsr_b200.set_precision(before)


def test_clear_caches_reaches_the_pass_cache():
    from nvsr_b200 import scene
    render._pass_cache.store[123] = ("x",)
    scene._plane_cache.store[5] = ("y",)
    scene.clear_caches()
    assert not render._pass_cache.store and not scene._plane_cache.store and not scene._decoder_cache.store
