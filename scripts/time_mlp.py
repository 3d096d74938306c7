"""Time the decoder kernel alone on synthetic features (density and rgb chains, fine-pass size)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvsr_b200
from nvsr_b200 import ops, scene
from nvsr_b200._lib import NVSR_F16
dev = torch.device("cuda", 0)
mc, mf, sid = scene.make_synthetic_scene(plane_res=64, view_res=16, seed=0, device=dev)
dec = scene.pack_planes_decoder(mf, NVSR_F16)
n, S = 32768, 192
rows = ops.rows_padded(n, S, ops.ROWS_BLOCKED)
tiles = rows // 128
fp = (torch.randn(tiles, 18, 128, 8, device=dev) * 0.3).half()
fm = (torch.randn(tiles, 6, 128, 8, device=dev) * 0.3).half()
rb = torch.randn(n, 128, device=dev)
raw = ops.raw_buffer(n, S, ops.ROWS_BLOCKED, dev)
mmc, mmf = scene.make_mip_models(seed=0, device=dev)
mdec = scene.pack_mip_decoder(mmf, NVSR_F16)
menc = (torch.randn(rows // 128, mdec.k0 // 8, 128, 8, device=dev) * 0.5).half()
mrb = torch.randn(n, mdec.dir_layer.n_out, device=dev)
raw_m = ops.raw_buffer(n, S, ops.ROWS_RAY_MAJOR, dev)
def run(which):
    if which == "mip":
        ops.mlp_chain(menc, mdec.chain(mrb), n * S, raw_m, NVSR_F16, S, n)
    elif which == "density":
        ops.mlp_chain(fm, dec.density, rows, raw, NVSR_F16, S, n, ops.ROWS_BLOCKED)
    else:
        ops.mlp_chain(fp, dec.rgb_chain(rb), rows, raw, NVSR_F16, S, n, ops.ROWS_BLOCKED)
for which in ("density", "rgb", "mip"):
    for _ in range(3): run(which)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(int(os.environ.get("REPS", "10"))): run(which)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / int(os.environ.get("REPS", "10"))
    fl = 2 * rows * (55424 if which == "density" else (74112 - 48 * 128 if which == "rgb" else 80384))
    print(f"NVSR_DBG={os.environ.get('NVSR_DBG','0'):>3s} {which:8s} {ms:7.3f} ms  {fl/ms/1e9:7.1f} TFLOP/s  cycles/tile {ms*1e-3*1.965e9/(tiles/148):7.0f}")
    # debug build with -DNVSR_TC_TIMING: per-warp, per-layer phase cycle sums of CTA 0
    import ctypes as C
    lib = nvsr_b200._lib.load()
    if hasattr(lib, "nvsr_debug_tc_timing"):
        buf = (C.c_ulonglong * 1024)()
        lib.nvsr_debug_tc_timing(buf)
        print(f"  {which}: average cycles per layer step (CTA 0, mean over the 8 warps of slot 0 half 0 / half 1): wait_acc / epilogue / arrive(+issue) / rest | cycles per issue")
        for l in range(6 if which == "mip" else 4):
            for half in (0, 1):
                ws = [w for w in range(8) if ((w & 7) >> 2) == half]
                t = [sum(buf[(w * 8 + l) * 8 + i] for w in ws) for i in range(8)]
                n_ = max(t[4], 1)
                print(f"   layer {l} half {half}: {t[0] / n_:7.0f} {t[1] / n_:7.0f} {t[2] / n_:7.0f} {t[3] / n_:7.0f} | sum {sum(t[:4]) / n_:7.0f} | issue {t[6] / max(t[5], 1):7.0f} (x{t[5]})")
