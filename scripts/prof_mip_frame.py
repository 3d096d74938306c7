"""One 800x800 mip/IPE frame (BASELINE config 3b) for an ncu launch list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nvsr_b200
from nvsr_b200 import scene
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
nvsr_b200.set_precision("fp16")
mc, mf = scene.make_mip_models(seed=0, device=dev)
pose, focal = scene.blender_camera(800)
pose = pose.to(dev)
opt, scfg = scene.render_options(64, 128, mip=True), scene.scene_cfg()
enc = nvsr_b200.IntegratedPositionalEncoding(3, 7)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        out = nvsr_b200.render_frame(800, 800, focal, pose, mc, mf, opt, "synth_DS2", scfg, encode_position_fn=enc,
                                     encode_direction_fn=object())
torch.cuda.synchronize()
print("acc", float(out[5].mean()))
