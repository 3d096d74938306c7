"""Scene / model state the render path reads, and how it is packed for the kernels.

Two things live here:
  1. duck-typed STAND-INS for the reference objects the hot path reads (`TwoDimPlanesModel`,
     `FlexibleNeRFModel`, `SceneCoupler`, `CfgNode`) — just enough state, with the reference's
     attribute names, to build synthetic Blender-shaped scenes on a box where the reference is
     not importable (tests, bench, smoke).  The render path itself works on the real reference
     objects as well: it only reads the attributes listed in SURVEY.md §8(b).
  2. the packers: planes NCHW fp32 -> nvsr_pack_plane images (fp32 channels-last / 16-bit x-pair records), decoder weights -> chain layers.
     Packing is cached per (tensor identity, version) so it happens once per scene / weight update,
     never per chunk (the reference re-uploads the SR plane for every network chunk, models.py:893).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import NVSR_BF16, NVSR_F16, NVSR_F32


# --------------------------------------------------------------------------------------------------
# config containers (cfgnode.py:36 CfgNode is an attribute-access dict; this is the minimal equal)
class Cfg(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def render_options(num_coarse=64, num_fine=128, perturb=False, lindisp=False, white_background=False,
                   noise_std=0.0, chunksize=131072, use_viewdirs=True, mip=False):
    """options.nerf.<mode>.* keys read by the path (train_utils.py:82-113,136-151,210-234)."""
    sub = dict(chunksize=chunksize, perturb=perturb, num_coarse=num_coarse, num_fine=num_fine,
               white_background=white_background, radiance_field_noise_std=noise_std, lindisp=lindisp)
    nerf = dict(use_viewdirs=use_viewdirs, train=dict(sub), validation=dict(sub))
    if mip:
        nerf["encode_position_fn"] = "mip"
    return Cfg(nerf=nerf)


def scene_cfg(near=2.0, far=6.0, no_ndc=True):
    return Cfg(near=near, far=far, no_ndc=no_ndc)


# --------------------------------------------------------------------------------------------------
# stand-ins
def get_plane_name(scene_id, dimension):  # naming scheme of models.py:110-113
    return "_D%d" % dimension if scene_id is None else "sc%s_D%d" % (scene_id, dimension)


class SingleSceneCoupler:
    """SceneCoupler (models.py:928-1019) for scenes without LR/HR pairing, plus an optional explicit
    HR->LR map for the super-resolution configuration (planes are stored under the LR id)."""

    def __init__(self, sr_pairs=None):
        self.downsample_couples = dict(sr_pairs or {})  # hr_scene -> lr_scene
        self.upsample_couples = {}
        self.HR_planes = []
        self.ds_factor = 1

    def _scene_of(self, plane_name):
        return plane_name[2:plane_name.rindex("_D")]

    def scene_with_saved_plane(self, name, plane_not_scene=False):
        if plane_not_scene:
            sc = self._scene_of(name)
            return name.replace(sc, self.downsample_couples.get(sc, sc))
        return self.downsample_couples.get(name, name)

    def should_SR(self, name, plane_not_scene=False):
        sc = self._scene_of(name) if plane_not_scene else name
        return sc in self.downsample_couples

    def should_downsample(self, plane_name, for_LR_loading=False):
        return False


class _Projector(nn.Module):
    def __init__(self):
        super().__init__()
        eye = torch.eye(3)
        # CoordProjector (models.py:477-478): plane d keeps columns [1:] of these
        self.rot_mats_NON_LEARNED = nn.ParameterList(
            [nn.Parameter(m.clone(), requires_grad=False) for m in (eye, eye[:, [1, 0, 2]], eye[:, [2, 0, 1]])])


class TriPlaneModel(nn.Module):
    """State-compatible stand-in of TwoDimPlanesModel (models.py:118-421) for the shipped config
    (config/TrainModels.yml:66-93): 3 position planes + 1 view plane, proj 'avg', view 'concat_pos',
    rgb input 'projections', density 48->128x4->1, rgb 192->128x4->3, no skip firing."""

    def __init__(self, num_plane_channels=48, dec_channels=128, dec_density_layers=4, dec_rgb_layers=4,
                 scene_coupler=None):
        super().__init__()
        self.use_viewdirs = True
        self.num_density_planes = 3
        self.num_plane_channels = num_plane_channels
        self.num_viewdir_plane_channels = num_plane_channels
        self.rgb_dec_input = "projections"
        self.proj_combination = "avg"
        self.viewdir_proj_combination = "concat_pos"
        self.plane_interp = "bilinear"
        self.align_corners = True
        self.skip_connect_every = 3
        self.skip_SR_ = False
        self.coord_projector = _Projector()
        self.scene_coupler = scene_coupler or SingleSceneCoupler()
        c = num_plane_channels
        self.density_dec = nn.ModuleDict({"0": nn.ModuleList(
            [nn.Linear(c, dec_channels)] + [nn.Linear(dec_channels, dec_channels) for _ in range(dec_density_layers - 1)])})
        self.fc_alpha = nn.ModuleDict({"0": nn.Linear(dec_channels, 1)})
        self.rgb_dec = nn.ModuleDict({"0": nn.ModuleList(
            [nn.Linear(4 * c, dec_channels)] + [nn.Linear(dec_channels, dec_channels) for _ in range(dec_rgb_layers - 1)])})
        self.fc_rgb = nn.ModuleDict({"0": nn.Linear(dec_channels, 3)})
        self.planes_ = nn.ParameterDict()
        self.box_coords = {}
        self.cur_id = None

    def set_cur_scene_id(self, scene_id):
        self.cur_id = scene_id

    def raw_plane(self, plane_name, downsample=False, detach=False):
        return self.planes_[plane_name]

    def assign_SR_model(self, sr_model):
        self.SR_model = sr_model
        self.skip_SR_ = False

    def planes(self, dim_num, super_resolve, grid=None):
        name = self.scene_coupler.scene_with_saved_plane(get_plane_name(self.cur_id, dim_num), plane_not_scene=True)
        return self.SR_model(name) if super_resolve else self.planes_[name]


class PlaneUpsampler(nn.Module):
    """Stand-in for PlanesSR (models.py:824-926): stock-PyTorch plane super-resolution, run once per
    plane and cached.  Only its OUTPUT tensor matters to the hot path (SR inference is out of scope,
    SURVEY.md §2 row 7); the architecture here is a bilinear x`scale` upsample plus a small residual conv."""

    def __init__(self, channels=48, scale=4, hidden=32):
        super().__init__()
        self.scale_factor = scale
        self.body = nn.Sequential(nn.Conv2d(channels, hidden, 3, padding=1), nn.ReLU(), nn.Conv2d(hidden, channels, 3, padding=1))
        self.LR_planes, self.SR_planes = {}, {}

    def set_LR_plane(self, plane, id):
        self.LR_planes[id] = plane

    @torch.no_grad()
    def forward(self, plane_name):
        if plane_name not in self.SR_planes:
            up = nn.functional.interpolate(self.LR_planes[plane_name], scale_factor=self.scale_factor, mode="bilinear",
                                           align_corners=True)
            self.SR_planes[plane_name] = up + 0.1 * self.body(up)
        return self.SR_planes[plane_name]


class _ResidualBlock(nn.Module):
    """state-compatible with models._Residual_Block (models.py:773-789): conv-relu-conv, x0.1, + (cropped) identity"""

    def __init__(self, hidden_size):
        super().__init__()
        self.conv1 = nn.Conv2d(hidden_size, hidden_size, 3, bias=False)
        self.conv2 = nn.Conv2d(hidden_size, hidden_size, 3, bias=False)


class EDSRNet(nn.Module):
    """State-compatible stand-in of models.EDSR (models.py:792-822) as PlanesSR builds it (padding=0: 'valid' 3x3
    convolutions on a replicate-padded input, no biases, no receptive-field bound): conv_input, `n_blocks` residual
    blocks, conv_mid, log2(scale) x [conv hidden -> 4 hidden, PixelShuffle(2)], conv_output.  Holds parameters only —
    the forward pass of the render path is nvsr_b200.sr.PlaneSuperResolver."""

    def __init__(self, in_channels, out_channels, hidden_size, n_blocks, scale_factor):
        super().__init__()
        self.conv_input = nn.Conv2d(in_channels, hidden_size, 3, bias=False)
        self.residual = nn.Sequential(*[_ResidualBlock(hidden_size) for _ in range(n_blocks)])
        self.conv_mid = nn.Conv2d(hidden_size, hidden_size, 3, bias=False)
        ups = []
        for _ in range(int(round(math.log2(scale_factor)))):
            ups += [nn.Conv2d(hidden_size, hidden_size * 4, 3, bias=False), nn.PixelShuffle(2)]
        self.upscale = nn.Sequential(*ups)
        self.conv_output = nn.Conv2d(hidden_size, out_channels, 3, bias=False)
        # receptive-field bookkeeping of models.py:796-803 (every conv is 3x3 here): padding in LR pixels
        rp, rf = 0.0, 1.0
        rp += rf            # conv_input
        rp += rf * 2 * n_blocks
        rp += rf            # conv_mid
        for _ in range(int(round(math.log2(scale_factor)))):
            rp += rf
            rf /= 2
        rp += rf            # conv_output
        self.required_padding = rp


class PlanesSRModel(nn.Module):
    """State-compatible stand-in of models.PlanesSR (models.py:824-926): `inner_model` (EDSR), scale factor, LR /
    SR plane dictionaries, the padding bookkeeping of :840-842 and the weight initialisation of :843-850.  Calling it
    runs the device-resident super-resolution of nvsr_b200.sr (no CPU cache, no per-call re-upload)."""

    def __init__(self, scale_factor=4, in_channels=48, out_channels=48, hidden_size=256, n_blocks=32, plane_interp="bilinear",
                 weight_gain=1.0):
        super().__init__()
        self.scale_factor = scale_factor
        self.plane_interp = plane_interp
        self.align_corners = True
        self.inner_model = EDSRNet(in_channels, out_channels, hidden_size, n_blocks, scale_factor)
        rp = self.inner_model.required_padding
        self.HR_overpadding = int(rp * scale_factor)
        self.inner_model.required_padding = int(math.ceil(rp))
        self.HR_overpadding = self.inner_model.required_padding * scale_factor - self.HR_overpadding
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n) / 10 * weight_gain)
        self.LR_planes, self.SR_planes, self.residual_planes = {}, {}, {}

    def set_LR_plane(self, plane, id, save_interpolated=False):
        self.LR_planes[id] = plane

    def clear_SR_planes(self, all_planes=False):
        self.SR_planes = {}
        if all_planes:
            self.LR_planes, self.residual_planes = {}, {}

    @torch.no_grad()
    def forward(self, plane_name):
        from . import sr
        if plane_name not in self.SR_planes:
            self.SR_planes[plane_name] = sr.resolver_of(self).super_resolve(plane_name, want_nchw=True)[1]
        return self.SR_planes[plane_name]


class MipMLP(nn.Module):
    """Stand-in of FlexibleNeRFModel as train_nerf.py:342-348 builds it for the mip baseline
    (models.py:14-108 with defaults num_layers=4, hidden=128, skip=4, use_viewdirs)."""

    def __init__(self, num_encoding_fn_xyz=6, num_encoding_fn_dir=4, hidden_size=128):
        super().__init__()
        self.dim_xyz = 2 * 3 * num_encoding_fn_xyz          # include_input_xyz=False
        self.dim_dir = 3 + 2 * 3 * num_encoding_fn_dir      # include_input_dir=True
        self.skip_connect_every = 4
        self.use_viewdirs = True
        self.xyz_input_2_dir = False
        self.layer1 = nn.Linear(self.dim_xyz, hidden_size)
        self.layers_xyz = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(3)])
        self.layers_dir = nn.ModuleList([nn.Linear(self.dim_dir + hidden_size, hidden_size // 2)])
        self.fc_alpha = nn.Linear(hidden_size, 1)
        self.fc_rgb = nn.Linear(hidden_size // 2, 3)
        self.fc_feat = nn.Linear(hidden_size, hidden_size)


DEFAULT_BOX = [[-1.5, -1.5, -1.5, -math.pi, -math.pi / 2], [1.5, 1.5, 1.5, math.pi, math.pi / 2]]


def shape_density(model, target_std=15.0, shift=-12.0, feat_std=0.5, n_draws=4096, seed=1234):
    """Random-init decoders give degenerate volumes (coarse acc == 1, fine sigma == 0).
    Standardise the density head on synthetic decoder inputs so that raw sigma ~ N(shift, target_std^2):
    positive on a minority of samples, with O(10) values.  Only the head's weight/bias are rescaled
    (weight initialisation, not part of the render path)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        if isinstance(model.fc_alpha, nn.ModuleDict):       # tri-plane decoder: input = mean of 3 bilinear samples
            head = model.fc_alpha["0"]
            c = model.num_plane_channels
            h = torch.randn(n_draws, c, generator=g) * (feat_std * (4.0 / 9.0 / 3.0) ** 0.5)
            for lin in model.density_dec["0"]:
                h = torch.relu(lin(h))
        else:                                               # mip MLP: input = IPE features in [-1,1]
            head = model.fc_alpha
            h = model.layer1(torch.sin(torch.rand(n_draws, model.dim_xyz, generator=g) * 6.2831853))
            for lin in model.layers_xyz:
                h = torch.relu(lin(h))
        y = head(h).squeeze(-1)
        mu, sd = float(y.mean()), float(y.std())
        a = target_std / max(sd, 1e-8)
        head.weight.mul_(a)
        head.bias.copy_(a * (head.bias - mu) + shift)


def make_synthetic_scene(plane_res=200, view_res=32, channels=48, seed=0, device="cpu", scene_id=None,
                         plane_std=0.5, density_std=10.0, density_shift=-10.0, sr_scale=None, sr_hidden=256, sr_blocks=32,
                         sr_weight_gain=6.0):
    """Synthetic Blender-shaped scene: seeded coarse+fine decoders, then planes.

    With `sr_scale`, planes are stored at plane_res under an LR id, and the FINE model reads the output of a
    `PlanesSRModel` — the reference's PlanesSR + EDSR architecture (config/TrainModels.yml:174-184: hidden 256,
    32 residual blocks), random-init with the reference's own initialiser scaled by `sr_weight_gain` so that the SR
    residual is a visible fraction of the plane — at plane_res*sr_scale, while the coarse model reads the LR planes
    (apply_2_coarse: False, config/TrainModels.yml:166) — BASELINE config 3a."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    sid = scene_id or "synth_DS2_PlRes%d_%d" % (plane_res * (sr_scale or 1), view_res)
    pairs = None
    stored_sid = sid
    if sr_scale:
        stored_sid = "synth_DS%d_PlRes%d_%d" % (2 * sr_scale, plane_res, view_res)
        pairs = {sid: stored_sid}
    coarse = TriPlaneModel(channels, scene_coupler=SingleSceneCoupler())
    fine = TriPlaneModel(channels, scene_coupler=SingleSceneCoupler(pairs))
    if sr_scale:
        coarse.scene_coupler = SingleSceneCoupler(pairs)  # same stored planes, but no SR model attached
    fine.coord_projector = coarse.coord_projector
    planes = nn.ParameterDict()
    for d in range(4):
        r = plane_res if d < 3 else view_res
        # smooth blobs + texel noise: spatially coherent density (some rays empty, some opaque)
        low = nn.functional.interpolate(torch.randn(1, channels, max(r // 8, 2), max(r // 8, 2)), size=(r, r),
                                        mode="bicubic", align_corners=True)
        planes[get_plane_name(stored_sid, d)] = nn.Parameter(plane_std * (0.9 * low + 0.45 * torch.randn(1, channels, r, r)))
    box = torch.tensor(DEFAULT_BOX, dtype=torch.float64)
    for m in (coarse, fine):
        m.planes_ = planes
        m.box_coords = {sid: box, stored_sid: box}
        shape_density(m, density_std, density_shift, plane_std)
        m.eval()
    coarse.to(device)
    fine.to(device)
    if sr_scale:
        sr = PlanesSRModel(sr_scale, channels, channels, sr_hidden, sr_blocks, weight_gain=sr_weight_gain).to(device).eval()
        fine.assign_SR_model(sr)
        for d in range(3):
            n = get_plane_name(stored_sid, d)
            sr.set_LR_plane(fine.planes_[n].detach(), n)
    return coarse, fine, sid


def add_synthetic_scene(coarse, fine, scene_id, plane_res=200, view_res=32, seed=1, plane_std=0.5):
    """Another scene's planes for an existing decoder pair (BASELINE config 5: several scenes share one decoder —
    in the reference `planes_` is one ParameterDict keyed by scene id, models.py:545-550, and `set_cur_scene_id`
    selects the scene per frame).  Same plane statistics as make_synthetic_scene, different seed."""
    g = torch.Generator().manual_seed(seed)
    channels = coarse.num_plane_channels
    dev = next(coarse.parameters()).device
    for d in range(4):
        r = plane_res if d < 3 else view_res
        low = nn.functional.interpolate(torch.randn(1, channels, max(r // 8, 2), max(r // 8, 2), generator=g), size=(r, r),
                                        mode="bicubic", align_corners=True)
        p = nn.Parameter((plane_std * (0.9 * low + 0.45 * torch.randn(1, channels, r, r, generator=g))).to(dev))
        coarse.planes_[get_plane_name(scene_id, d)] = p      # fine shares the same ParameterDict object
    box = torch.tensor(DEFAULT_BOX, dtype=torch.float64)
    for m in (coarse, fine):
        m.box_coords[scene_id] = box
    return scene_id


def make_mip_models(seed=0, device="cpu", density_std=10.0, density_shift=-10.0):
    torch.manual_seed(seed)
    coarse, fine = MipMLP(), MipMLP()
    for m in (coarse, fine):
        shape_density(m, density_std, density_shift)
        m.eval().to(device)
    return coarse, fine


def blender_camera(width, theta=30.0, phi=-30.0, radius=4.0, camera_angle_x=0.6911112070083618):
    """pose_spherical (load_blender.py:34-39) + focal from camera_angle_x (load_blender.py:257-258,288)."""
    def rx(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)

    def ry(t):
        return np.array([[np.cos(t), 0, -np.sin(t), 0], [0, 1, 0, 0], [np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1]], np.float32)

    tz = np.eye(4, dtype=np.float32)
    tz[2, 3] = radius
    c2w = ry(theta / 180.0 * np.pi) @ (rx(phi / 180.0 * np.pi) @ tz)
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    focal = 0.5 * width / np.tan(0.5 * camera_angle_x)
    return torch.from_numpy(c2w).float(), float(focal)


# --------------------------------------------------------------------------------------------------
# packers
# Bumped whenever parameters may have changed WITHOUT their `_version` moving — a CUDA-graph replay that contains the
# optimizer update (autograd.GraphedStep) runs no Python, so no version counter ticks.  Part of every cache signature.
_GENERATION = [0]


def bump_generation():
    _GENERATION[0] += 1


class _Cache:
    """Identity + version cache.  Entries hold a WEAK reference to the object they were made from:
    `id()`/`data_ptr()` alone can be recycled by a new object after the old one died, which would
    silently serve another scene's packed planes.  An entry is valid only while the very same object
    is alive and its (data_ptr, _version) signature is unchanged."""

    def __init__(self):
        self.store = {}

    @staticmethod
    def key_of(t):
        return (t.data_ptr(), t._version, tuple(t.shape), str(t.device), _GENERATION[0])

    def get(self, obj, sig, make):
        import weakref
        hit = self.store.get(id(obj))
        if hit is not None and hit[0]() is obj and hit[1] == sig:
            return hit[2]
        for k in [k for k, v in self.store.items() if v[0]() is None]:  # entries whose owners died
            del self.store[k]
        val = make()
        self.store[id(obj)] = (weakref.ref(obj), sig, val)
        return val


_plane_cache = _Cache()
_decoder_cache = _Cache()


def clear_caches():
    """Drop every packed plane / decoder image AND the per-(model, scene) pass objects of the render module that own
    them (`render.clear_caches`).  Needed after in-place writes through `.data` — they do not bump `_version`, which
    the cache keys rely on — and to release the device memory of scenes that are no longer rendered."""
    from . import render
    render.clear_caches()


def _should_sr(model, d):
    """models.py:296-300"""
    name = get_plane_name(model.cur_id, d)
    sr = hasattr(model, "SR_model") and (not hasattr(model, "scene_coupler") or
                                         model.scene_coupler.should_SR(name, plane_not_scene=True))
    return bool(sr and not getattr(model, "skip_SR_", False))


def _source_plane(model, d):
    """The NCHW plane tensor `model` reads for dimension d of the current scene, with a STABLE identity.

    PlanesSR caches its output on the CPU and re-uploads it on every call (models.py:893,925), so
    `model.planes()` returns a fresh tensor each time; here the cached tensor itself is used as the
    cache key, making the SR plane device-resident after the first frame."""
    sr = d < 3 and _should_sr(model, d)
    if sr and _native_sr(model):
        # device-resident SR (nvsr_b200.sr): the SR plane is a function of the LR plane and the SR weights — the LR
        # plane is the identity the caches key on, the SR weights are part of the model's parameter signature
        return _lr_plane_of(model, d)[1]
    if sr and hasattr(model.SR_model, "SR_planes"):
        name = model.scene_coupler.scene_with_saved_plane(get_plane_name(model.cur_id, d), plane_not_scene=True)
        if name not in model.SR_model.SR_planes:
            model.planes(d, super_resolve=True)  # stock-PyTorch SR inference, once per plane
        if name in model.SR_model.SR_planes:
            return model.SR_model.SR_planes[name]
    return model.planes(d, super_resolve=sr)


def _native_sr(model):
    from . import sr
    return hasattr(model, "SR_model") and sr.is_sr_model(model.SR_model)


def _lr_plane_of(model, d):
    """(stored plane name, LR plane tensor) the SR model super-resolves for dimension d of the current scene; registers
    it with the SR model like TwoDimPlanesModel.assign_LR_planes (models.py:426-434) if that has not happened"""
    name = model.scene_coupler.scene_with_saved_plane(get_plane_name(model.cur_id, d), plane_not_scene=True)
    lr_planes = model.SR_model.LR_planes
    if name not in lr_planes:
        model.SR_model.set_LR_plane(model.raw_plane(name, detach=True) if hasattr(model, "raw_plane") else model.planes_[name],
                                    id=name, save_interpolated=False)
    return name, lr_planes[name]


def check_supported_planes_model(model):
    """Engagement rule: unsupported configurations raise — there is no fallback."""
    bad = []
    if getattr(model, "num_density_planes", 3) != 3:
        bad.append("num_density_planes != 3")
    if getattr(model, "plane_interp", "bilinear") != "bilinear" or not getattr(model, "align_corners", True):
        bad.append("plane_interp/align_corners")
    if getattr(model, "proj_combination", "avg") not in ("avg", "sum"):
        bad.append("proj_combination not in ('avg', 'sum')")
    if getattr(model, "viewdir_proj_combination", "concat_pos") != "concat_pos":
        bad.append("viewdir_proj_combination != 'concat_pos'")
    if getattr(model, "rgb_dec_input", "projections") != "projections" or not getattr(model, "use_viewdirs", True):
        bad.append("rgb_dec_input/use_viewdirs")
    s = getattr(model, "skip_connect_every", None)
    n_layers = max(len(model.density_dec["0"]), len(model.rgb_dec["0"]))
    if s is not None and any(((i - 1) % s == 0 and (i - 1) > 0) for i in range(n_layers)):
        bad.append("a skip connection fires")
    if len(model.density_dec) != 1:
        bad.append("ensemble_size != 1")
    if getattr(model, "point_coords_noise", 0) and model.training:
        bad.append("point_coords_noise in training mode")
    if bad:
        raise NotImplementedError("nvsr_b200: unsupported TwoDimPlanesModel configuration: " + ", ".join(bad))


def _packed_plane(src, dtype):
    """packed device image of one NCHW plane tensor, cached per (tensor object, version) and dtype"""
    per_dtype = _plane_cache.get(src, _Cache.key_of(src), dict)
    if dtype not in per_dtype:
        per_dtype[dtype] = ops.pack_plane(src.cuda(), dtype)
    return per_dtype[dtype]


def pack_scene_planes(model, scene_id, dtype):
    """Channels-last device planes for `scene_id` as `model` would read them (models.py:270-310)."""
    model.set_cur_scene_id(scene_id)
    packed = []
    for d in range(3):
        if _should_sr(model, d) and _native_sr(model):
            # SR inference on the device, written by nvsr_sr_finalize directly as the gather's image (nvsr_b200.sr)
            from . import sr
            name, _ = _lr_plane_of(model, d)
            packed.append(sr.resolver_of(model.SR_model).super_resolve(name, dtype)[0])
            continue
        src = _source_plane(model, d)
        packed.append(_packed_plane(src, dtype))
    vsrc = _source_plane(model, 3)
    vplane = _packed_plane(vsrc, NVSR_F32)
    box = model.box_coords[scene_id].detach().double().cpu()
    lo = box[0].float()                  # .type(coords.type()) of the fp64 box (models.py:264)
    rng = (box[1] - box[0]).float()      # difference in fp64, then cast (models.py:265)
    rots = model.coord_projector.rot_mats_NON_LEARNED
    proj = [rots[d].detach().float().cpu()[:, 1:].tolist() for d in range(3)]
    return ops.PackedPlanes(packed, dtype, lo[:3].tolist(), rng[:3].tolist(), proj, vplane,
                            (float(lo[3]), float(rng[3]), float(lo[4]), float(rng[4])),
                            combine=getattr(model, "proj_combination", "avg"))


def pack_scene_planes_hilo(model, scene_id):
    """The fp16-split mode's planes: (PackedPlanes of the fp16 x-pair images of fp16(p), list of the three images of
    p - fp16(p)), both cut from the SAME fp32 planes (an SR plane is evaluated in fp32 once), so that hi + lo carries ~22
    bits of every texel — what nvsr_sample_gather_hilo interpolates the combined features from."""
    base = pack_scene_planes(model, scene_id, NVSR_F32)          # fp32 channels-last [Rh,Rw,C]
    hi, lo = [], []
    for img in base.planes:
        src = img.permute(2, 0, 1)[None].contiguous()              # NCHW, as nvsr_pack_plane reads it
        hi.append(ops.pack_plane(src, ops.NVSR_F16))
        lo.append(ops.pack_plane((src - src.half().float()).contiguous(), ops.NVSR_F16))
    packed = ops.PackedPlanes(hi, ops.NVSR_F16, base.box_lo, base.box_rng, base.proj, base.vplane, base.view_lo_rng,
                              combine=base.combine)
    return packed, lo


class PackedPlanesDecoder:
    """Decoder chains of one TwoDimPlanesModel instance, in fp32 (SIMT) or bf16 (tcgen05) form."""

    def __init__(self, model, precision):
        dd, rd = list(model.density_dec["0"]), list(model.rgb_dec["0"])
        fa, fr = model.fc_alpha["0"], model.fc_rgb["0"]
        c = model.num_plane_channels
        self.precision = precision
        self.view_w = rd[0].weight.detach()[:, 3 * c:]      # per-ray columns of rgb_dec[0] (models.py:186)
        self.view_b = rd[0].bias.detach().float().contiguous()

        def wpack(w):
            w = w.detach().float()
            return ops.pack_weight16(w, dtype=precision) if precision != NVSR_F32 else w.contiguous()

        def f(t):
            return t.detach().float().contiguous()

        self.density = []
        for i, lin in enumerate(dd):
            last = i == len(dd) - 1
            self.density.append(ops.ChainLayer(wpack(lin.weight), f(lin.bias), lin.in_features, lin.out_features, True,
                                               head_w=f(fa.weight) if last else None,
                                               head_b=f(fa.bias) if last else None, head_ch=3))
        self.rgb = []
        for i, lin in enumerate(rd):
            last = i == len(rd) - 1
            w = lin.weight.detach()[:, :3 * c] if i == 0 else lin.weight
            self.rgb.append(ops.ChainLayer(wpack(w), None if i == 0 else f(lin.bias), w.shape[1], lin.out_features, True,
                                           head_w=f(fr.weight) if last else None,
                                           head_b=f(fr.bias) if last else None, head_ch=0))

    def density_split(self, model):
        """(w_hi, w_lo, biases, head_w, head_b) of the density chain for nvsr_mlp_chain_split: W = fp16(W) + fp16(W - fp16(W))"""
        if not hasattr(self, "_split"):
            dd, fa = list(model.density_dec["0"]), model.fc_alpha["0"]
            if len(dd) != 4 or any(l.out_features != 128 for l in dd):
                raise NotImplementedError("nvsr_b200: the 'fp16-split' mode serves the 4 x 128 density chain")
            w_hi, w_lo = [], []
            for lin in dd:
                w = lin.weight.detach().float()
                hi = w.half().float()
                w_hi.append(ops.pack_weight16(hi, dtype=NVSR_F16))
                w_lo.append(ops.pack_weight16(w - hi, dtype=NVSR_F16))
            f = lambda t: t.detach().float().contiguous()
            self._split = (w_hi, w_lo, [f(l.bias) for l in dd], f(fa.weight), f(fa.bias))
        return self._split

    def rgb_chain(self, row_bias):
        first = self.rgb[0]
        l0 = ops.ChainLayer(first.w, None, first.k, first.n_out, True, row_bias=row_bias,
                            head_w=first.head_w, head_b=first.head_b, head_ch=first.head_ch)
        return [l0] + self.rgb[1:]


def pack_planes_decoder(model, precision):
    params = list(model.density_dec["0"].parameters()) + list(model.rgb_dec["0"].parameters()) + \
        list(model.fc_alpha["0"].parameters()) + list(model.fc_rgb["0"].parameters())
    sig = tuple((p.data_ptr(), p._version) for p in params) + (_GENERATION[0],)
    per_prec = _decoder_cache.get(model, sig, dict)
    if precision not in per_prec:
        per_prec[precision] = PackedPlanesDecoder(model, precision)
    return per_prec[precision]


class PackedMipDecoder:
    """FlexibleNeRFModel chain (models.py:85-108): layer1 (no ReLU) -> 3x(128,ReLU) -> [fc_alpha tap]
    -> fc_feat (ReLU) -> layers_dir[0] (per-ray view bias, ReLU) -> fc_rgb head."""

    def __init__(self, model, precision):
        if not model.use_viewdirs or getattr(model, "xyz_input_2_dir", False) or len(model.layers_dir) != 1:
            raise NotImplementedError("nvsr_b200: unsupported FlexibleNeRFModel configuration")
        n_xyz = len(model.layers_xyz)
        if any(i % model.skip_connect_every == 0 and i > 0 and i != n_xyz for i in range(n_xyz)):
            raise NotImplementedError("nvsr_b200: FlexibleNeRFModel with a firing skip connection")
        self.precision = precision
        self.dim_xyz = model.dim_xyz
        self.k0 = (model.dim_xyz + 15) // 16 * 16 if precision != NVSR_F32 else model.dim_xyz

        def wpack(w, k_pad=None):
            w = w.detach().float()
            if precision != NVSR_F32:
                return ops.pack_weight16(w, k_pad, precision)
            if k_pad is not None and k_pad != w.shape[1]:
                w = torch.cat([w, w.new_zeros(w.shape[0], k_pad - w.shape[1])], 1)
            return w.contiguous()

        def f(t):
            return t.detach().float().contiguous()

        hid = model.layer1.out_features
        L = [ops.ChainLayer(wpack(model.layer1.weight, self.k0), f(model.layer1.bias), self.k0, hid, False)]
        for i, lin in enumerate(model.layers_xyz):
            last = i == n_xyz - 1
            L.append(ops.ChainLayer(wpack(lin.weight), f(lin.bias), lin.in_features, lin.out_features, True,
                                    head_w=f(model.fc_alpha.weight) if last else None,
                                    head_b=f(model.fc_alpha.bias) if last else None, head_ch=3))
        L.append(ops.ChainLayer(wpack(model.fc_feat.weight), f(model.fc_feat.bias), hid, hid, True))
        dl = model.layers_dir[0]
        self.dir_w = dl.weight.detach()[:, hid:]
        self.dir_b = f(dl.bias)
        self.dir_layer = ops.ChainLayer(wpack(dl.weight.detach()[:, :hid]), None, hid, dl.out_features, True,
                                        head_w=f(model.fc_rgb.weight), head_b=f(model.fc_rgb.bias), head_ch=0)
        self.front = L

    def chain(self, row_bias):
        d = self.dir_layer
        return self.front + [ops.ChainLayer(d.w, None, d.k, d.n_out, True, row_bias=row_bias, head_w=d.head_w,
                                            head_b=d.head_b, head_ch=0)]


def pack_mip_decoder(model, precision):
    params = list(model.parameters())
    sig = tuple((p.data_ptr(), p._version) for p in params) + (_GENERATION[0],)
    per_prec = _decoder_cache.get(model, sig, dict)
    if precision not in per_prec:
        per_prec[precision] = PackedMipDecoder(model, precision)
    return per_prec[precision]
