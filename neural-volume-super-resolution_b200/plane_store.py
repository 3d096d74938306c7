"""Scene / plane store: the on-disk format in front of the gather.

The reference keeps one file per (model, scene): `<save_location>/<model_name>_<scene>.par`, a `torch.save` of
`{'params': nn.ParameterDict{plane name: [1,C,R,R]}, 'opt_states': [...], 'coords_normalization': box [2,5]}`
(models.py:612-613 param_path, :655-668 save_params) written through `safe_saving` (nerf_helpers.py:19-48:
write `.par_temp`, move the old file to `.par_bckp`, rename temp into place, drop the backup) and read through
`safe_loading` (:50-67: try the file, then `_temp`, then `_bckp`; `_best` variants) — `PlanesOptimizer.load_scene`
(models.py:589-610) then does `model.planes_ = params.cuda()` SYNCHRONOUSLY on the render stream and sets `box_coords`.

Here the same files are read and written with the same protocol (both directions are tested against the reference's
own functions), but a scene reaches the GPU differently:
  * `prefetch(scene)` reads and unpickles on a background thread into PINNED host tensors;
  * `to_device(scene)` issues the host->device copies on a side stream and returns tensors guarded by an event — the
    render stream waits on the event only when it first touches the scene (`attach`), so loading scene k+1 overlaps
    rendering scene k (multi-scene video, BASELINE config 5);
  * `broadcast(scene, src)` ships the planes of a scene from one rank to all others over torch.distributed (NCCL
    on GPUs, gloo in the CPU tests) instead of every rank reading the file system.
Unpickling a `.par` executes pickle code, exactly as in the reference: load only files you trust.
"""
import os
import threading

import torch
from torch import nn

SUFFIX = "par"


def param_path(save_location, model_name, scene, prefer_best=False, file_must_exist=True):
    """models.py:612-627: first location (a path or a list of paths) that holds the file (or, when
    `file_must_exist` is False, that is an existing directory); '' when none does.  The returned path never carries
    the `_best` suffix — `safe_loading(best=True)` appends it, as in the reference."""
    locs = save_location if isinstance(save_location, (list, tuple)) else [save_location]
    for loc in locs:
        p = os.path.join(loc, "%s_%s.%s" % (model_name, scene, SUFFIX))
        if file_must_exist:
            if os.path.isfile(p.replace(".par", ".par_best") if prefer_best else p):
                return p
        elif os.path.isdir(loc):
            return p
    return ""


def _variant(file_name, version):
    return file_name.replace(".%s" % SUFFIX, ".%s%s" % (SUFFIX, version))


def safe_loading(file_name, best=False, map_location="cpu"):
    """nerf_helpers.py:50-67 for `.par` files: the file itself, else its `_temp`, else its `_bckp` sibling."""
    if best:
        file_name = _variant(file_name, "_best")
    last = None
    for version in ("", "_temp", "_bckp"):
        try:
            return torch.load(_variant(file_name, version), map_location=map_location, weights_only=False)
        except Exception as e:  # corrupted or missing: fall through to the next sibling, as the reference does
            last = e
    raise last


def safe_saving(file_name, content, best=False):
    """nerf_helpers.py:35-48 (the run-signature guard of :20-33 belongs to the training loop and is not reproduced):
    the new content is complete on disk before the old file is touched."""
    if best:
        file_name = _variant(file_name, "_best")
    torch.save(content, _variant(file_name, "_temp"))
    had_old = os.path.isfile(file_name)
    if had_old:
        os.rename(file_name, _variant(file_name, "_bckp"))
    os.rename(_variant(file_name, "_temp"), file_name)
    if had_old:
        os.remove(_variant(file_name, "_bckp"))


def plane_names(scene, n_planes=4):
    """models.py:110-113 get_plane_name for d = 0..n_planes-1 (3 position planes + the view-direction plane)."""
    return ["sc%s_D%d" % (scene, d) for d in range(n_planes)]


class SceneRecord:
    """One scene's planes as read from disk: `planes` {name: fp32 [1,C,R,R] host tensor (pinned when possible)},
    `box` = coords_normalization [2,5], `opt_states` as stored (opaque to this package)."""

    def __init__(self, planes, box, opt_states=None):
        self.planes, self.box, self.opt_states = planes, box, opt_states

    def nbytes(self):
        return sum(p.numel() * p.element_size() for p in self.planes.values())


def _pin(t):
    t = t.detach().contiguous()
    if torch.cuda.is_available():
        try:
            return t.pin_memory()
        except RuntimeError:
            pass
    return t


class PlaneStore:
    """Reads/writes the reference's `.par` files and stages scenes for the GPU (see the module docstring)."""

    def __init__(self, save_location, model_name="coarse", device=None, prepack=None):
        """`prepack`: "fp16" | "bf16" | "fp32" — also build the gather's packed plane images (`nvsr_pack_plane`) on the
        COPY stream right behind the upload and seed the render path's plane cache with them, so the first frame of a
        scene finds its planes packed (the packing of scene k+1 overlaps the rendering of scene k as well)."""
        self.save_location, self.model_name = save_location, model_name
        self.device = device
        self.prepack = prepack
        self._host = {}        # scene -> SceneRecord (pinned)
        self._pending = {}     # scene -> (thread, result dict)
        self._device = {}      # scene -> (dict name -> device tensor, box, event or None)
        self._copy_stream = None
        self._lock = threading.Lock()

    # ---- disk ------------------------------------------------------------------------------------
    def path(self, scene, prefer_best=False, file_must_exist=True):
        return param_path(self.save_location, self.model_name, scene, prefer_best, file_must_exist)

    def read(self, scene, prefer_best=False):
        """load_scene_planes (models.py:670-678) -> SceneRecord with pinned planes."""
        f = self.path(scene, prefer_best)
        if not f:
            raise FileNotFoundError("Could not find the required feature planes file for scene %s" % scene)
        content = safe_loading(f, best=prefer_best)
        planes = {k: _pin(v.data.float()) for k, v in content["params"].items()}
        return SceneRecord(planes, content["coords_normalization"], content.get("opt_states"))

    def write(self, scene, planes, box, opt_states=None, as_best=False):
        """save_params' file content (models.py:667-668): ParameterDict + opt_states + coords_normalization."""
        f = self.path(scene, file_must_exist=False)
        if not f:
            raise FileNotFoundError("no existing directory among %r" % (self.save_location,))
        params = nn.ParameterDict([(k, nn.Parameter(v.detach().cpu().clone())) for k, v in planes.items()])
        safe_saving(f, {"params": params, "opt_states": opt_states if opt_states is not None else [None] * len(params),
                        "coords_normalization": box}, best=as_best)
        return f

    # ---- host staging ----------------------------------------------------------------------------
    def prefetch(self, scene, prefer_best=False):
        """Start reading `scene` on a background thread (file read + unpickle + pinning release the GIL mostly)."""
        with self._lock:
            if scene in self._host or scene in self._pending:
                return
            box = {}

            def work():
                try:
                    box["rec"] = self.read(scene, prefer_best)
                except BaseException as e:   # surfaced by host_record()
                    box["err"] = e

            t = threading.Thread(target=work, daemon=True)
            self._pending[scene] = (t, box)
            t.start()

    def host_record(self, scene, prefer_best=False):
        with self._lock:
            pend = self._pending.pop(scene, None)
        if pend is not None:
            pend[0].join()
            if "err" in pend[1]:
                raise pend[1]["err"]
            self._host[scene] = pend[1]["rec"]
        if scene not in self._host:
            self._host[scene] = self.read(scene, prefer_best)
        return self._host[scene]

    def evict(self, scene):
        self._host.pop(scene, None)
        self._device.pop(scene, None)

    # ---- device staging --------------------------------------------------------------------------
    def to_device(self, scene, device=None):
        """Issue the host->device copies of `scene` on the store's copy stream; returns immediately.
        The tensors become valid for a stream once it has waited on the scene's event (`attach` does)."""
        device = torch.device(device or self.device or "cuda")
        if scene in self._device:
            return self._device[scene]
        rec = self.host_record(scene)
        if device.type != "cuda":
            raise RuntimeError("PlaneStore.to_device needs a CUDA device: nvsr_b200 has no CPU render path")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(self._copy_stream):
            # Parameters are made HERE (not in attach) so that their identity can key the render path's plane cache
            planes = {k: nn.Parameter(v.to(device, non_blocking=True), requires_grad=False) for k, v in rec.planes.items()}
            if self.prepack is not None:
                from . import ops, render, scene as scene_mod
                dtype = render._PRECISION[self.prepack]
                for k, p in planes.items():
                    is_view = k.endswith("_D%d" % (len(planes) - 1))
                    dt = ops.NVSR_F32 if is_view else dtype            # the view plane is always gathered in fp32
                    per_dtype = scene_mod._plane_cache.get(p, scene_mod._Cache.key_of(p), dict)
                    per_dtype[dt] = ops.pack_plane(p, dt)
            ev = torch.cuda.Event()
            ev.record()
        self._device[scene] = (planes, rec.box, ev)
        return self._device[scene]

    def attach(self, models, scene, saved_scene=None):
        """PlanesOptimizer.load_scene (models.py:589-610) for the render path: every model in `models` reads the planes
        of `scene` (stored under `saved_scene`, scene_coupler.scene2saved) from now on; the CURRENT stream waits for
        the copies, the host does not."""
        saved = saved_scene or scene
        planes, box, ev = self.to_device(saved)
        if ev is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            from . import scene as scene_mod
            for p in planes.values():
                p.record_stream(cur)
                hit = scene_mod._plane_cache.store.get(id(p))
                if hit is not None and hit[0]() is p:                  # pre-packed images were allocated on the copy stream
                    for img in hit[2].values():
                        img.record_stream(cur)
        params = nn.ParameterDict([(k, v if isinstance(v, nn.Parameter) else nn.Parameter(v, requires_grad=False))
                                   for k, v in planes.items()])
        for m in models:
            m.planes_ = params
            m.box_coords = {saved: box, scene: box}
        return params

    # ---- multi-GPU -------------------------------------------------------------------------------
    def broadcast(self, scene, src=0, group=None, device=None):
        """Rank `src` holds (or reads) the scene; every other rank receives plane names, shapes, box and values over
        torch.distributed and ends up with the same host-or-device record without touching the file system."""
        import torch.distributed as dist
        rank = dist.get_rank(group)
        dev = torch.device(device) if device is not None else torch.device("cpu")
        meta = [None]
        if rank == src:
            rec = self.host_record(scene)
            meta = [([(k, tuple(v.shape)) for k, v in rec.planes.items()], rec.box.double().tolist())]
        dist.broadcast_object_list(meta, src=src, group=group)
        shapes, box = meta[0]
        out = {}
        for k, shape in shapes:
            t = rec.planes[k].to(dev) if rank == src else torch.empty(shape, dtype=torch.float32, device=dev)
            dist.broadcast(t, src=src, group=group)
            out[k] = t
        box_t = torch.tensor(box, dtype=torch.float64)
        if dev.type == "cuda":
            self._device[scene] = (out, box_t, None)
        else:
            self._host[scene] = SceneRecord(out, box_t)
        return out, box_t
