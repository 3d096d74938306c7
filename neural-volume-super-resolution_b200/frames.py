"""Either side of the render path: camera-path generators in front of it, the frame sink
behind it.

Camera paths (host, numpy — a few hundred 4x4 matrices per video):
  pose_spherical          load_blender.py:15-39      camera on a sphere looking at the origin (Blender scenes)
  orbit_poses             load_blender.py:308-311       the 360-degree evaluation orbit
  look_along / poses_avg / render_path_spiral                 load_llff.py:143-186   forward-facing spiral (batched)
  interpolate_pose_rows   load_llff.py:73-78         `min_eval_frames` pose interpolation
Frame sink:
  to_uint8                train_nerf.py:270,273      255*clamp(im,0,1) -> uint8 ON THE DEVICE (nvsr_frame_to_u8)
  encode_png              (imageio.imwrite in the reference) stdlib-only PNG writer: zlib + crc32
  FrameSink               converts on the device, copies one byte per channel into pinned host buffers on a side stream
                          (double-buffered, no host synchronisation per frame), hands finished frames to a writer
  AviWriter               (imageio.mimwrite, train_nerf.py:273) stdlib AVI container: uncompressed or MJPEG (Pillow) frames
There is no CPU path for the conversion: `to_uint8` refuses CPU tensors.
"""
import struct
import zlib

import numpy as np
import torch

from . import _lib, ops


# --------------------------------------------------------------------------------------------------
# camera paths
def pose_spherical(theta, phi, radius):
    """load_blender.py:15-39; angles in degrees.  Returns a [4,4] camera-to-world matrix (float64, as the reference's
    product of a float32 chain with an int64 matrix is)."""
    def trans_t(t):
        m = np.eye(4, dtype=np.float32)
        m[2, 3] = t
        return m

    def rot_phi(p):
        m = np.eye(4, dtype=np.float32)
        m[1, 1] = m[2, 2] = np.cos(p)
        m[1, 2] = -np.sin(p)
        m[2, 1] = -m[1, 2]
        return m

    def rot_theta(th):
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[2, 2] = np.cos(th)
        m[0, 2] = -np.sin(th)
        m[2, 0] = -m[0, 2]
        return m

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180.0 * np.pi) @ c2w
    c2w = rot_theta(theta / 180 * np.pi) @ c2w
    return np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w


def orbit_poses(n_frames=40, phi=-30.0, radius=4.0):
    """The evaluation orbit of the Blender loader: pose_spherical(angle, -30, 4) for angle in
    linspace(-180, 180, n+1)[:-1] (load_blender.py:308-311, n = 40 there; SURVEY §8d config 5 uses n = 200).  [n,4,4] float32."""
    return np.stack([pose_spherical(a, phi, radius) for a in np.linspace(-180, 180, n_frames + 1)[:-1]], 0).astype(np.float32)


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def look_along(z, up, pos):
    """Camera frames [...,3,4] = [x | y | z | pos] looking along `z` with `up` as the rough vertical — the frame
    construction of load_llff.viewmatrix (:147-153), batched over leading dimensions."""
    z = _unit(np.asarray(z, dtype=np.float64))
    x = _unit(np.cross(np.broadcast_to(up, z.shape), z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, np.broadcast_to(pos, z.shape)], -1)


def poses_avg(poses):
    """Average pose of [N,3,5] LLFF poses (load_llff.poses_avg :161-170): mean centre, summed z and y axes, the hwf
    column of the first pose.  Returns [3,5]."""
    poses = np.asarray(poses, dtype=np.float64)
    frame = look_along(poses[:, :, 2].sum(0), poses[:, :, 1].sum(0), poses[:, :, 3].mean(0))
    return np.concatenate([frame, poses[0, :, 4:5]], -1)


def render_path_spiral(c2w, up, rads, focal, zdelta, zrate, rots, N):
    """N poses [N,3,5] on the spiral of load_llff.render_path_spiral (:173-186) around the pose `c2w` [3,5]: centres
    c2w @ (cos t * rx, -sin t * ry, -sin(t * zrate) * rz, 1), every camera looking at the point `focal` in front of
    c2w.  All N frames are built at once (`zdelta` is accepted and unused, as in the reference)."""
    c2w = np.asarray(c2w, dtype=np.float64)
    rot, origin = c2w[:, :3], c2w[:, 3]
    t = np.linspace(0.0, 2.0 * np.pi * rots, N + 1)[:-1]
    rads = np.asarray(list(rads), dtype=np.float64)
    local = np.stack([np.cos(t), -np.sin(t), -np.sin(t * zrate)], -1) * rads          # [N,3] in the camera frame
    centres = local @ rot.T + origin
    target = origin - focal * rot[:, 2]
    frames_ = look_along(centres - target, up, centres)
    return np.concatenate([frames_, np.broadcast_to(c2w[:, 4:5], (N, 3, 1))], -1)


def interpolate_pose_rows(poses_arr, min_eval_frames):
    """load_llff.py:73-78: linear interpolation of the rows of poses_bounds.npy up to at least `min_eval_frames` rows
    (rounded up so that the original rows stay on the grid; they are written back verbatim).  Returns
    (rows [M,17], repeat) with M = repeat*(N-1)+1."""
    n = len(poses_arr)
    m = int(np.ceil(min_eval_frames / (n - 1)) * (n - 1) + 1)
    repeat = (m - 1) // (n - 1)
    x = np.linspace(start=0, stop=n - 1, num=m)
    # scipy.interpolate.interp1d(kind='linear', axis=0): y[lo] + slope * (x - lo), lo = searchsorted - 1 clipped
    hi = np.clip(np.searchsorted(np.arange(n), x), 1, n - 1)
    lo = hi - 1
    slope = (poses_arr[hi] - poses_arr[lo]) / (np.arange(n)[hi] - np.arange(n)[lo])[:, None]
    out = slope * (x - lo)[:, None] + poses_arr[lo]
    out[::repeat, :] = poses_arr
    return out, repeat


# --------------------------------------------------------------------------------------------------
# frame sink
def to_uint8(image, out=None):
    """255*clamp(image,0,1) -> uint8 on the device (train_nerf.py:270,273); same shape as `image`."""
    lib = _lib.load()
    if not image.is_cuda:
        raise _lib.NvsrError("image must be a CUDA tensor: nvsr_b200 has no CPU path")
    src = image.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        st = ops._call("nvsr_frame_to_u8", lib.nvsr_frame_to_u8, ops._ptr(src), src.numel(), ops._ptr(out), ops._stream(),
                       bytes=src.numel() * 5)
    _lib.check(st, "nvsr_frame_to_u8")
    return out


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(u8, level=3):
    """uint8 [H,W,3] (or [H,W] / [H,W,1] / [H,W,4]) -> PNG file bytes; filter type 0 on every scanline."""
    a = np.ascontiguousarray(u8)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise ValueError("encode_png expects a uint8 [H,W] or [H,W,C] array")
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    color = {1: 0, 3: 2, 4: 6}.get(c)
    if color is None:
        raise ValueError("encode_png: 1, 3 or 4 channels")
    raw = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * c)], 1).tobytes()
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0))
            + _chunk(b"IDAT", zlib.compress(raw, level)) + _chunk(b"IEND", b""))


def decode_png(data):
    """Inverse of encode_png for its own output (filter 0, 8 bit, non-interlaced) — used by the tests."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(data):
        n, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, color = hdr[:4]
    c = {0: 1, 2: 3, 6: 4}[color]
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * c)
    assert depth == 8 and not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, c)


class FrameSink:
    """Takes rendered frames off the GPU without stalling the render stream.

    submit(frame [H,W,3] fp32 CUDA): converts to uint8 on the device (render stream), then a side stream copies the
    bytes into one of `depth` pinned host buffers and records an event; the render stream never waits for the host.
    A buffer is reused only after its frame has been handed to `writer(index, uint8 ndarray [H,W,3])`, which happens
    on later submit() calls and in flush().  Default writer: keep the arrays in `self.frames`."""

    def __init__(self, writer=None, depth=2):
        self.writer = writer
        self.depth = int(depth)
        self.frames = []
        self._slots = []       # (index, host tensor, event, device uint8 tensor kept alive)
        self._free = []
        self._count = 0
        self._copy_stream = None

    def _emit(self, index, host):
        arr = host.numpy().copy()
        if self.writer is None:
            self.frames.append(arr)
        else:
            self.writer(index, arr)

    def _drain(self, wait_all):
        while self._slots and (wait_all or len(self._slots) >= self.depth or self._slots[0][2].query()):
            index, host, ev, _dev = self._slots.pop(0)
            ev.synchronize()
            self._emit(index, host)
            self._free.append(host)

    def submit(self, frame):
        u8 = to_uint8(frame)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=u8.device)
        self._drain(wait_all=False)
        host = next((h for h in self._free if h.shape == u8.shape), None)
        if host is not None:
            self._free.remove(host)
        else:
            host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
        done = torch.cuda.Event()
        self._copy_stream.wait_stream(torch.cuda.current_stream(u8.device))
        with torch.cuda.stream(self._copy_stream):
            host.copy_(u8, non_blocking=True)
            done.record()
        u8.record_stream(self._copy_stream)
        self._slots.append((self._count, host, done, u8))
        self._count += 1
        return self._count - 1

    def flush(self):
        self._drain(wait_all=True)
        return self.frames


def png_writer(directory, pattern="%d.png"):
    """writer for FrameSink: one PNG per frame, named like write_image does (train_nerf.py:268: '%d.png')."""
    import os
    os.makedirs(directory, exist_ok=True)

    def write(index, arr):
        with open(os.path.join(directory, pattern % index), "wb") as f:
            f.write(encode_png(arr))

    return write


class AviWriter:
    """Video sink for FrameSink: the reference assembles evaluation frames into a video with `imageio.mimwrite`
    (train_nerf.py:273; needs ffmpeg, absent here).  This writes an AVI (RIFF) file with the standard library alone:
    `codec='raw'` stores uncompressed bottom-up BGR frames ('DIB '), `codec='mjpg'` stores one JPEG per frame (needs
    Pillow for the JPEG encoder; plays in any player).  Frames must arrive in order (FrameSink delivers them so).

        with AviWriter("orbit.avi", fps=30, codec="mjpg") as w:
            sink = FrameSink(writer=w); ...; sink.flush()
    """

    def __init__(self, path, fps=30, codec="raw", quality=92):
        if codec not in ("raw", "mjpg"):
            raise ValueError("codec must be 'raw' or 'mjpg'")
        self.path, self.fps, self.codec, self.quality = path, int(fps), codec, int(quality)
        self._f = open(path, "wb")
        self._index = []          # (offset relative to 'movi', size)
        self._shape = None
        self._next = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _header(self, n_frames, movi_bytes, idx_bytes):
        h, w = self._shape
        fourcc = b"MJPG" if self.codec == "mjpg" else b"\x00\x00\x00\x00"
        handler = b"MJPG" if self.codec == "mjpg" else b"DIB "
        frame_bytes = max((s for _, s in self._index), default=w * h * 3)
        avih = struct.pack("<14I", 1000000 // self.fps, frame_bytes * self.fps, 0, 0x10, n_frames, 0, 1, frame_bytes, w, h,
                           0, 0, 0, 0)
        strh = struct.pack("<4s4sIHHIIIIIIIIhhhh", b"vids", handler, 0, 0, 0, 0, 1, self.fps, 0, n_frames, frame_bytes,
                           0xFFFFFFFF, 0, 0, 0, w, h)
        strf = struct.pack("<IiiHH4sIiiII", 40, w, h, 1, 24, fourcc, w * h * 3, 0, 0, 0, 0)
        strl = b"LIST" + struct.pack("<I", 4 + 8 + len(strh) + 8 + len(strf)) + b"strl" + \
            b"strh" + struct.pack("<I", len(strh)) + strh + b"strf" + struct.pack("<I", len(strf)) + strf
        hdrl = b"LIST" + struct.pack("<I", 4 + 8 + len(avih) + len(strl)) + b"hdrl" + b"avih" + struct.pack("<I", len(avih)) + avih + strl
        movi_head = b"LIST" + struct.pack("<I", 4 + movi_bytes) + b"movi"
        riff_size = 4 + len(hdrl) + len(movi_head) + movi_bytes + idx_bytes
        return b"RIFF" + struct.pack("<I", riff_size) + b"AVI " + hdrl + movi_head

    def __call__(self, index, arr):
        if index != self._next:
            raise ValueError(f"AviWriter: frame {index} out of order (expected {self._next})")
        arr = np.ascontiguousarray(arr)
        if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError("AviWriter takes uint8 [H,W,3] frames")
        if self._shape is None:
            self._shape = arr.shape[:2]
            self._hdr_len = len(self._header(0, 0, 0))
            self._f.write(b"\0" * self._hdr_len)       # patched in close()
            self._movi = 0
        elif arr.shape[:2] != self._shape:
            raise ValueError("AviWriter: frame size changed")
        if self.codec == "mjpg":
            import io
            from PIL import Image
            buf = io.BytesIO()
            Image.fromarray(arr).save(buf, format="JPEG", quality=self.quality)
            data, tag = buf.getvalue(), b"00dc"
        else:
            h, w = self._shape
            pad = (-(w * 3)) % 4
            rows = arr[::-1, :, ::-1]                    # bottom-up, BGR
            data = rows.tobytes() if pad == 0 else b"".join(r.tobytes() + b"\0" * pad for r in rows)
            tag = b"00db"
        chunk = tag + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")
        self._index.append((4 + self._movi, len(data)))
        self._f.write(chunk)
        self._movi += len(chunk)
        self._next += 1

    def close(self):
        if self._f is None:
            return
        if self._shape is not None:
            tag = b"00dc" if self.codec == "mjpg" else b"00db"
            idx = b"".join(tag + struct.pack("<III", 0x10, off, size) for off, size in self._index)
            idx1 = b"idx1" + struct.pack("<I", len(idx)) + idx
            self._f.write(idx1)
            self._f.seek(0)
            head = self._header(len(self._index), self._movi, len(idx1))
            assert len(head) == self._hdr_len
            self._f.write(head)
        self._f.close()
        self._f = None


def read_avi_frames(path):
    """Minimal reader of AviWriter's files (tests, round trips): -> (fps, list of uint8 [H,W,3] frames)."""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"RIFF" and data[8:12] == b"AVI "
    pos, frames, fps, wh, mjpg = 12, [], None, None, False
    end = 8 + struct.unpack("<I", data[4:8])[0]

    def walk(p, stop):
        nonlocal fps, wh, mjpg
        while p + 8 <= stop:
            tag, n = data[p:p + 4], struct.unpack("<I", data[p + 4:p + 8])[0]
            body = data[p + 8:p + 8 + n]
            if tag == b"LIST":
                walk(p + 12, p + 8 + n)
            elif tag == b"avih":
                fps = round(1e6 / struct.unpack("<I", body[:4])[0])
                wh = struct.unpack("<II", body[32:40])
            elif tag == b"strf":
                mjpg = body[16:20] == b"MJPG"
            elif tag in (b"00db", b"00dc"):
                w, h = wh
                if tag == b"00dc":
                    import io
                    from PIL import Image
                    frames.append(np.asarray(Image.open(io.BytesIO(body)).convert("RGB")))
                else:
                    stride = (w * 3 + 3) // 4 * 4
                    a = np.frombuffer(body, np.uint8).reshape(h, stride)[:, :w * 3].reshape(h, w, 3)
                    frames.append(a[::-1, :, ::-1].copy())
            p += 8 + n + (n & 1)

    walk(pos, end)
    return fps, frames
