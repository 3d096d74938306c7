import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; a no-op marker when the plugin is absent)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not silently skip: nothing to do here.
    # Without a GPU and without -m filtering, gpu tests are skipped.
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _training_decoder_mode(request):
    """The gradient tests pinned to the reference's fp32 goldens (and their CPU dry runs with host stand-ins) use the
    fp32 parity mode of the differentiable path; the tcgen05 training decoder (the default) has its own tests
    (tests/test_gpu_train_tc.py)."""
    if request.module.__name__ in ("test_backward_bodies", "test_gpu_next_rows", "test_gpu_tests_dry_run"):
        from nvsr_b200 import autograd
        autograd.set_decoder("fp32")
        yield
        autograd.set_decoder("tc")
    else:
        yield
