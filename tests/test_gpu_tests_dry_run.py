"""Dry run of the GPU tests of tests/test_gpu_next_rows.py on the CPU: the SAME test functions, with the device
set to "cpu", the C-ABI calls replaced by the host stand-ins of tests/host_ops.py (the kernels' own bodies built for
the host), entered right behind the entry points' CUDA-only guards.  What this checks is the tests themselves — their plumbing, shapes and
tolerances — so that their first run on a B200 measures the kernels and not a typo in a test."""
import shutil

import pytest
import torch

import host_ops as HO
import test_gpu_next_rows as T
from nvsr_b200 import ops


@pytest.fixture
def cpu_as_device(monkeypatch):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    for name, fn in HO.standins(HO.build_hostcheck()).items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(T, "DEV", "cpu")
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    # the two public entry points refuse CPU tensors (no CPU path in the product): the dry run enters right behind
    # that guard instead of faking `is_cuda` on every tensor of the process
    A = T.A
    monkeypatch.setattr(A, "volume_render_radiance_field",
                        lambda rf, z, rd, radiance_field_noise_std=0.0, white_background=False, mip_nerf=False, noise=None:
                        A._render(rf, z, rd, radiance_field_noise_std, white_background, noise, mip_nerf))
    monkeypatch.setattr(A, "run_one_iter_of_nerf",
                        lambda H, W, focal, mc, mf, batch, options, scene_id, mode="train", encode_position_fn=None,
                        encode_direction_fn=None, scene_config=None, randoms=None:
                        A._run_one_iter(H, W, focal, mc, mf, batch, options, scene_id, mode, scene_config, randoms, encode_position_fn))


@pytest.mark.parametrize("white,noise_std,mip", [(False, 0.0, False), (True, 0.6, False), (True, 0.3, True)])
def test_dry_composite_bwd(cpu_as_device, white, noise_std, mip):
    T.test_composite_bwd_matches_oracle_autograd(white, noise_std, mip)


def test_dry_gather_bwd(cpu_as_device):
    T.test_gather_bwd_matches_oracle_autograd()


def test_dry_train_step_golden(cpu_as_device, monkeypatch):
    T.test_train_step_gradients_match_reference_golden(monkeypatch)


def test_dry_mip_train_step(cpu_as_device):
    T.test_mip_train_step_gradients_match_oracle_autograd()


def test_dry_mass_conservation(cpu_as_device, monkeypatch):
    from nvsr_b200 import scene
    real = scene.make_synthetic_scene
    # the full-size test builds 200^2 planes; the dry run keeps the code path and shrinks the scene
    monkeypatch.setattr(scene, "make_synthetic_scene", lambda plane_res=200, view_res=32, **k: real(plane_res=24, view_res=8, **k))
    T.test_gather_bwd_full_batch_mass_conservation()
