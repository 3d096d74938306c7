"""Frame sink and camera paths: host logic and the conversion kernel's body on the CPU.

Camera paths are checked against the live reference where its checkout exists (tests/golden/check_live_reference.py) and
against committed vectors the reference produced (tests/golden/stage_camera_paths.npz) everywhere else."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import helpers as H
import nvsr_b200
from nvsr_b200 import frames

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "neural-volume-super-resolution_b200", "csrc")


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", out,
                    os.path.join(HERE, "hostcheck", "hostcheck.cpp")], check=True)
    return C.CDLL(out)


def reference_u8(im):
    """write_image's expression, verbatim (train_nerf.py:270)"""
    with np.errstate(invalid="ignore"):
        return np.array(255 * torch.clamp(im, 0, 1).cpu()).astype(np.uint8)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 7, 4096 * 3 + 2])
def test_to_u8_body_matches_write_image(hc, n):
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, generator=g) * 1.6 - 0.3
    if n >= 5:
        x[0], x[1], x[2], x[3], x[4] = 1.0, 0.0, float("nan"), float("inf"), -float("inf")
    if n > 100:   # values right at the byte boundaries k/255
        k = torch.arange(256, dtype=torch.float32) / 255.0
        x[100:356] = k
        x[400:656] = torch.nextafter(k, torch.tensor(2.0))
        x[700:956] = torch.nextafter(k, torch.tensor(-1.0))
    out = torch.full((n + 4,), 77, dtype=torch.uint8)
    hc.hc_frame_to_u8(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_void_p(out.data_ptr()))
    assert np.array_equal(out[:n].numpy(), reference_u8(x))
    assert (out[n:] == 77).all()          # nothing written past the end


def test_png_roundtrip():
    rng = np.random.default_rng(0)
    for shape in ((5, 7, 3), (1, 1, 3), (16, 9), (4, 4, 4)):
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        data = frames.encode_png(a)
        back = frames.decode_png(data)
        assert np.array_equal(back.reshape(a.shape), a)
        try:                                   # an independent decoder, when the box has one
            import io
            from PIL import Image
        except ImportError:
            continue
        assert np.array_equal(np.array(Image.open(io.BytesIO(data))).reshape(a.shape), a)
    with pytest.raises(ValueError):
        frames.encode_png(np.zeros((2, 2, 3), np.float32))


def test_camera_paths_match_reference_vectors():
    g = H.golden("stage_camera_paths.npz")
    for i, (th, ph, r) in enumerate(g["spherical_args"]):
        assert np.array_equal(frames.pose_spherical(th, ph, r), g["spherical"][i])
    assert np.array_equal(frames.orbit_poses(40), g["orbit40"])
    poses = g["llff_poses"]
    # the LLFF generators are batched re-formulations (one matrix product for all frames): equal to a few ulp of fp64
    c2w = frames.poses_avg(poses)
    np.testing.assert_allclose(c2w, g["poses_avg"], rtol=0, atol=1e-12)
    spiral = frames.render_path_spiral(g["poses_avg"], g["up"], g["rads"], float(g["focal"]), float(g["zdelta"]), 0.5, 2, 30)
    assert spiral.shape == g["spiral"].shape
    np.testing.assert_allclose(spiral, g["spiral"], rtol=0, atol=1e-12)
    for f in spiral[:, :, :3]:                                   # orthonormal, right-handed frames
        np.testing.assert_allclose(f.T @ f, np.eye(3), atol=1e-12)
        assert np.linalg.det(f) > 0.999
    rows, rep = frames.interpolate_pose_rows(g["pose_rows"], int(g["min_eval_frames"]))
    assert rep == int(g["repeat"]) and rows.shape == g["pose_rows_interp"].shape
    np.testing.assert_allclose(rows, g["pose_rows_interp"], rtol=0, atol=1e-12)
    assert np.array_equal(rows[::rep], g["pose_rows"])


def test_to_uint8_refuses_cpu_tensors():
    with pytest.raises(nvsr_b200.NvsrError):
        frames.to_uint8(torch.zeros(2, 2, 3))


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_frame_sink_order_and_buffer_reuse(hc, monkeypatch):
    """FrameSink's host logic with the CUDA pieces replaced by stand-ins: frames reach the writer in submission order,
    exactly once, with the bytes write_image would produce; at most `depth` copies are in flight; pinned buffers are
    recycled only after their frame has been handed over."""
    import contextlib

    class Ev:
        def __init__(self, *a, **k):
            self.done = False

        def record(self):
            pending.append(self)

        def query(self):
            return self.done

        def synchronize(self):
            self.done = True

    class Stream:
        def __init__(self, *a, **k):
            pass

        def wait_stream(self, other):
            pass

    pending, allocated = [], []

    def to_u8(frame):
        x = frame.contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8)
        hc.hc_frame_to_u8(C.c_void_p(x.data_ptr()), C.c_int64(x.numel()), C.c_void_p(out.data_ptr()))
        return out

    real_pin = torch.Tensor.pin_memory
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: allocated.append(self) or self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "Stream", Stream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(frames, "to_uint8", to_u8)
    got = []
    sink = frames.FrameSink(writer=lambda i, arr: got.append((i, arr)), depth=2)
    g = torch.Generator().manual_seed(0)
    imgs = [torch.rand(6, 5, 3, generator=g) * 1.4 - 0.2 for _ in range(7)]
    for k, im in enumerate(imgs):
        assert sink.submit(im) == k
        assert len(sink._slots) <= 2                              # never more than `depth` copies in flight
        if k == 3:
            for e in pending:                                     # the copies issued so far complete "asynchronously"
                e.done = True
    sink.flush()
    assert [i for i, _ in got] == list(range(7)) and not sink._slots
    for (i, arr), im in zip(got, imgs):
        assert np.array_equal(arr, reference_u8(im))
    assert len(allocated) <= 3                                    # buffers are recycled, not allocated per frame
    # default writer keeps the frames
    keep = frames.FrameSink()
    keep.submit(imgs[0])
    assert len(keep.flush()) == 1 and np.array_equal(keep.frames[0], reference_u8(imgs[0]))


def test_avi_writer_round_trip(tmp_path):
    """Video sink (the reference: imageio.mimwrite, train_nerf.py:273): AVI written with the standard library alone;
    uncompressed frames come back byte-exact (odd widths: 4-byte row padding), MJPEG frames within JPEG quality."""
    from nvsr_b200 import frames as F
    rng = np.random.default_rng(0)
    for h, w in ((6, 7), (16, 24)):
        fr = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(4)]
        p = str(tmp_path / f"raw_{w}.avi")
        with F.AviWriter(p, fps=24, codec="raw") as wr:
            for i, a in enumerate(fr):
                wr(i, a)
        fps, got = F.read_avi_frames(p)
        assert fps == 24 and len(got) == 4 and all(np.array_equal(a, b) for a, b in zip(fr, got))
    # smooth frames through MJPEG
    yy, xx = np.mgrid[0:32, 0:48]
    fr = [np.stack([(4 * xx + 10 * k) % 256, 6 * yy % 256, (xx + yy) * 3 % 256], -1).astype(np.uint8) for k in range(3)]
    p = str(tmp_path / "m.avi")
    with F.AviWriter(p, fps=30, codec="mjpg", quality=95) as wr:
        for i, a in enumerate(fr):
            wr(i, a)
        with pytest.raises(ValueError):
            wr(7, fr[0])                      # out of order
    fps, got = F.read_avi_frames(p)
    assert fps == 30 and len(got) == 3
    assert all(float(np.abs(a.astype(int) - b.astype(int)).mean()) < 12.0 for a, b in zip(fr, got))
    # the RIFF size field covers the file
    import struct
    data = open(p, "rb").read()
    assert struct.unpack("<I", data[4:8])[0] + 8 == len(data)
