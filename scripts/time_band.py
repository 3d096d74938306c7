"""Frame time of a row band (what one rank of an N-GPU run renders), GPU time vs host wall time: is the host
(Python + launches + allocator) keeping up when the band is an eighth of the frame?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, nvsr_b200
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_w = bench.build_workload('cfg2', dev)
mc, mf, sid, pose, focal, opt, scfg = _w.mc, _w.mf, _w.sid, _w.pose, _w.focal, _w.opt, _w.scfg
pose = pose.to(dev)
for rows in (800, 400, 200, 100):
    with torch.no_grad():
        for _ in range(3):
            nvsr_b200.render_frame(bench.RES, bench.RES, focal, pose, mc, mf, opt, sid, scfg, row_range=(0, rows))
        torch.cuda.synchronize()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            nvsr_b200.render_frame(bench.RES, bench.RES, focal, pose, mc, mf, opt, sid, scfg, row_range=(0, rows))
        e1.record()
        t_host = (time.perf_counter() - t0) / n * 1e3     # host time to ENQUEUE a frame
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    print(f"rows {rows:4d}: {ms:7.2f} ms/frame on the device, host enqueue {t_host:6.2f} ms/frame, ideal {62.6 * rows / 800:6.2f}")
