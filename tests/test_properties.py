"""Property tests (hypothesis) of the host logic and of the host-built kernel bodies: size-independent properties the
domain offers — row-order bijections, partition properties, linearity of the backward in the upstream gradient,
monotonicity and fixed points of the byte conversion, encode/decode round trips."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import nvsr_b200
from nvsr_b200 import _lib, frames, ops, sharding

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "neural-volume-super-resolution_b200", "csrc")
_HC = {}


def hostcheck():
    if "lib" not in _HC:
        if shutil.which("g++") is None:
            pytest.skip("g++ not available")
        import tempfile
        out = os.path.join(tempfile.mkdtemp(prefix="hostcheck"), "libhostcheck.so")
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", out,
                        os.path.join(HERE, "hostcheck", "hostcheck.cpp")], check=True)
        _HC["lib"] = C.CDLL(out)
    return _HC["lib"]


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5000), S=st.integers(1, 1024), order=st.sampled_from([ops.ROWS_RAY_MAJOR, ops.ROWS_BLOCKED]))
def test_rows_padded_matches_the_library(n, S, order):
    lib = _lib.load()          # host-side arithmetic of the C-ABI: no device needed
    assert lib.nvsr_rows_padded(n, S, order) == ops.rows_padded(n, S, order)
    assert ops.rows_padded(n, S, order) >= n * S
    if order == ops.ROWS_BLOCKED:
        assert ops.rows_padded(n, S, order) % ops.TILE_ROWS == 0


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 70), S=st.integers(1, 70))
def test_blocked_row_order_is_the_documented_bijection(n, S):
    """include/nvsr.h: tile = (ray/8)*ceil(S/16) + s/16, row = tile*128 + (s%16)*8 + ray%8; raw_to_nsc inverts it."""
    rows = ops.rows_padded(n, S, ops.ROWS_BLOCKED)
    raw = torch.full((4, rows), -1.0)
    ts = -(-S // 16)
    ray, s = torch.meshgrid(torch.arange(n), torch.arange(S), indexing="ij")
    row = ((ray // 8) * ts + s // 16) * 128 + (s % 16) * 8 + ray % 8
    assert row.max() < rows and torch.unique(row).numel() == n * S          # injective, in range
    raw[0, row.reshape(-1)] = (ray * 10000 + s).reshape(-1).float()
    back = ops.raw_to_nsc(raw, n, S, ops.ROWS_BLOCKED)
    assert torch.equal(back[..., 0], (ray * 10000 + s).float())


@settings(max_examples=200, deadline=None)
@given(h=st.integers(1, 2000), w=st.integers(1, 16))
def test_row_bands_partition_the_frame(h, w):
    bands = [sharding.row_band(h, r, w) for r in range(w)]
    assert bands[0][0] == 0 and bands[-1][1] == h
    assert all(b[1] == c[0] for b, c in zip(bands, bands[1:])) and all(b[1] >= b[0] for b in bands)
    assert max(b[1] - b[0] for b in bands) == sharding.rows_per_rank(h, w)


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 9), S=st.integers(1, 40), white=st.booleans(), mip=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_composite_bwd_is_linear_in_the_upstream_gradient(n, S, white, mip, seed):
    hc = hostcheck()
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(n, S, 4, generator=g) * 2
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S + int(mip), generator=g), -1).values
    rd = torch.randn(n, 3, generator=g)

    def run(g_rgb, g_acc, g_depth, g_w):
        out = torch.empty(n, S, 4)
        hc.hc_composite_bwd(_p(raw), _p(z), _p(rd), None, C.c_int64(n), S, int(white), int(mip), _p(g_rgb), _p(g_acc),
                            _p(g_depth), _p(g_w), _p(out))
        return out

    a = [torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g)]
    b = [torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, S, generator=g)]
    both = run(*[x + 2 * y for x, y in zip(a, b)])
    want = run(*a) + 2 * run(*b)
    scale = float(want.abs().max()) + 1e-20
    assert float((both - want).abs().max()) <= 1e-4 * scale
    zero = run(torch.zeros(n, 3), None, None, None)
    assert not zero.any()


def test_to_u8_fixed_points_and_monotonicity():
    hc = hostcheck()
    k = torch.arange(256, dtype=torch.float32)
    # requantising a decoded byte is the identity for the values 255*x hits exactly
    x = torch.cat([k / 255.0 + 1e-6, torch.linspace(-0.5, 1.5, 4001)]).contiguous()
    out = torch.empty(x.numel(), dtype=torch.uint8)
    hc.hc_frame_to_u8(_p(x), C.c_int64(x.numel()), _p(out))
    assert torch.equal(out[:256].long(), k.long())
    ramp = out[256:].long()
    assert bool((ramp[1:] >= ramp[:-1]).all()) and ramp[0] == 0 and ramp[-1] == 255


@settings(max_examples=40, deadline=None)
@given(h=st.integers(1, 40), w=st.integers(1, 40), c=st.sampled_from([1, 3, 4]), seed=st.integers(0, 10 ** 6))
def test_png_roundtrip_any_shape(h, w, c, seed):
    a = np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)
    assert np.array_equal(frames.decode_png(frames.encode_png(a)), a)


@settings(max_examples=60, deadline=None)
@given(n=st.integers(2, 12), m=st.integers(1, 80), seed=st.integers(0, 10 ** 6))
def test_pose_interpolation_keeps_the_original_rows(n, m, seed):
    rows = np.random.default_rng(seed).standard_normal((n, 17))
    out, rep = frames.interpolate_pose_rows(rows, m)
    assert out.shape[0] == rep * (n - 1) + 1 >= m and np.array_equal(out[::rep], rows)
    # every interpolated row lies between its neighbours
    lo, hi = np.minimum(rows[:-1], rows[1:]), np.maximum(rows[:-1], rows[1:])
    for j in range(out.shape[0] - 1):
        seg = min(j // rep, n - 2)
        assert np.all(out[j] >= lo[seg] - 1e-12) and np.all(out[j] <= hi[seg] + 1e-12)
