"""Time the gather and composite kernels alone on a synthetic scene (coarse- and fine-pass sizes of one
32768-ray chunk).  NVSR_B200_LIB selects a prebuilt library for same-box A/B timing of kernel variants."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvsr_b200
from nvsr_b200 import ops, scene
from nvsr_b200._lib import NVSR_F16

dev = torch.device("cuda", 0)
torch.manual_seed(0)
n, Sc, nf = 32768, 64, 128
Sf = Sc + nf
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tag = os.path.basename(os.environ.get("NVSR_B200_LIB", "default"))
mc, mf, sid = scene.make_synthetic_scene(plane_res=200, view_res=32, seed=0, device=dev)
pose, focal = scene.blender_camera(800)
with torch.no_grad():
    ro, rd = nvsr_b200.get_ray_bundle(800, 800, focal, pose.to(dev))
ro, rd = ro.reshape(-1, 3)[300 * 800:300 * 800 + n].contiguous(), rd.reshape(-1, 3)[300 * 800:300 * 800 + n].contiguous()
u = torch.linspace(0, 1, nf, device=dev)
t_vals = torch.linspace(0, 1, Sc, device=dev)
packed = scene.pack_scene_planes(mc, sid, NVSR_F16)
# realistic coarse weights: render the coarse pass once through the real path to get z_merged
fp, fm, zc = ops.sample_gather(ro, rd, 2.0, 6.0, packed, ops.FEAT_TILE_F16, t_vals=t_vals)
rows_c = ops.rows_padded(n, Sc, ops.ROWS_BLOCKED)
raw_c = (torch.randn(4, rows_c, device=dev) * 2)
# density concentrated around a surface so that the resampled depths cluster like a real scene
raw_c[3] = raw_c[3] * 3 - 4
oc = ops.composite(raw_c, zc, rd, Sc, n_fine=nf, u=u, row_order=ops.ROWS_BLOCKED)
zf = oc["z_merged"]
rows_f = ops.rows_padded(n, Sf, ops.ROWS_BLOCKED)
raw_f = (torch.randn(4, rows_f, device=dev) * 2)
if which in ("all", "composite"):
    tc = timed(lambda: ops.composite(raw_c, zc, rd, Sc, n_fine=nf, u=u, row_order=ops.ROWS_BLOCKED))
    tf = timed(lambda: ops.composite(raw_f, zf, rd, Sf, row_order=ops.ROWS_BLOCKED))
    print(f"{tag:24s} composite coarse {tc:7.1f} us   fine {tf:7.1f} us")
if which in ("all", "gather"):
    bc = ops.feature_buffers(n, Sc, 48, ops.FEAT_TILE_F16, dev)
    bf = ops.feature_buffers(n, Sf, 48, ops.FEAT_TILE_F16, dev)
    tc = timed(lambda: ops.sample_gather(ro, rd, 2.0, 6.0, packed, ops.FEAT_TILE_F16, t_vals=t_vals, out=bc))
    tf = timed(lambda: ops.sample_gather(ro, rd, 2.0, 6.0, packed, ops.FEAT_TILE_F16, z_in=zf, out=bf))
    print(f"{tag:24s} gather    coarse {tc:7.1f} us   fine {tf:7.1f} us")
    dc = ops.feature_buffers(n, Sc, 48, ops.FEAT_TILE_F16, dev, density_only=True)
    df = ops.feature_buffers(n, Sf, 48, ops.FEAT_TILE_F16, dev, density_only=True)
    tc = timed(lambda: ops.sample_gather(ro, rd, 2.0, 6.0, packed, ops.FEAT_TILE_F16, t_vals=t_vals, out=dc, density_only=True))
    tf = timed(lambda: ops.sample_gather(ro, rd, 2.0, 6.0, packed, ops.FEAT_TILE_F16, z_in=zf, out=df, density_only=True))
    print(f"{tag:24s} gather-M  coarse {tc:7.1f} us   fine {tf:7.1f} us   (density features only)")
