#!/usr/bin/env python
"""bench.py — rays/s & samples/s of the 800x800 render (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full frame of BASELINE config 2: synthetic Blender-shaped scene (random-init tri-planes
200^2 x 48 ch + 32^2 view plane, shared 4+4-layer x128 decoder pair), 800x800 rays, 64 coarse + 128
fine samples (hierarchical sample_pdf) -> rgb/disp/acc, coarse and fine.
  value : whole-job rays/s, rays resident in HBM, timed on the device (CUDA events), max over ranks
  e2e   : same metric through run_one_iter_of_nerf with HOST ray buffers: H2D of the rays and D2H of
          the six result maps inside the timed region
  N > 1 : the frame's rows are split into N contiguous bands (ray order preserved), one rank per GPU,
          one NCCL all_gather of the result tiles per frame (inside the timed region)
  --impl reference : the reference algorithm's CPU path (oracle port, all host threads) on a bounded
          ray sample of the same workload.
  sparse colour path (the library's default): the rgb decoder is evaluated only for samples whose
          density (+ noise) is > 0 — every other sample has alpha = 0 and weight exactly 0, so every output map is
          bit-identical to evaluating all samples (tests/test_gpu_e2e.py::test_sparse_rgb_equals_dense).  The
          HEADLINE (value, e2e, roofline, kernels) is nevertheless measured with EVERY sample through both
          decoders — the work the reference does; the same frame with the sparse path is reported beside it under
          `sparse` (value, e2e, rgb_rows_evaluated = fraction of the samples that reached the rgb decoder).
          `--sparse` swaps the two (the companion is then `dense`).  samples_per_s is the reference's nominal
          count, rays/s x (Nc + (Nc + Nf)).
  --config cfgN : make another BASELINE.json config the headline workload (default cfg2, the one the metric is quoted on).
          The default line also carries a `configs` block — every BASELINE config at FULL size, 1 warm-up + 3 timed frames
          each, same event timing and sharding — plus, at N = 1: `precision_modes` (the frame in the fp32 1e-3-parity
          mode), `torch_gpu_baseline` (informative: the reference's op sequence in stock PyTorch on this GPU), `train_step`
          (informative, SURVEY §8f-1: the 4 096-ray training iteration of the same path — ours eager / replayed from a CUDA
          graph / stock PyTorch) and the CPU baselines BASELINE.md §4 asks for (cfg1 exactly: 100x100, 64+0, 1 + 3 frames,
          median).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = 800
NC, NF = 64, 128
PLANE_RES = 200
EVALS_PER_RAY = NC + (NC + NF)          # decoder evaluations per ray (coarse net + fine net on merged set)
FLOP_PER_EVAL = 259072                  # true MACs x 2, planes decoder


def ncu_traffic(kernel, evals_per_launch):
    """dram read+write bytes per launch of `kernel`: the committed `ncu --set full` capture (profiles/) gives
    bytes per decoder evaluation (launch sizes follow the ray chunk), scaled to this run's launches"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            per_eval = json.load(f).get(kernel + "_bytes_per_eval")
        return None if per_eval is None else per_eval * evals_per_launch
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Workload:
    """One BASELINE.json config at full size: scene, camera, sampling, and how a frame of it is rendered."""

    def __init__(self, name, H, W, nc, nf, workload, desc, flop_per_eval=FLOP_PER_EVAL):
        self.name, self.H, self.W, self.nc, self.nf = name, H, W, nc, nf
        self.workload, self.desc, self.flop_per_eval = workload, desc, flop_per_eval
        self.evals_per_ray = nc + ((nc + nf) if nf > 0 else 0)
        self.rays = H * W
        self.enc = self.encd = None
        self.offset = 0.0
        self.sids = None          # cfg5: several scenes behind one decoder pair, rendered round-robin
        self._frame = 0

    def metric(self):
        return "rays/s (%dx%d render, %d coarse + %d fine samples/ray)" % (self.W, self.H, self.nc, self.nf)

    def scene_id(self, advance=True):
        if not self.sids:
            return self.sid
        sid = self.sids[self._frame % len(self.sids)]
        self._frame += int(advance)
        return sid

    def render(self, nv, a, b):
        """device frame of image rows [a, b): rays generated on the device"""
        kw = {}
        if self.offset:
            kw["downsampling_offset"] = self.offset
        if self.enc is not None:
            kw.update(encode_position_fn=self.enc, encode_direction_fn=self.encd)
        return nv.render_frame(self.H, self.W, self.focal, self.pose, self.mc, self.mf, self.opt, self.scene_id(), self.scfg,
                               row_range=(a, b), **kw)

    def run_rays(self, nv, batch):
        """the reference-facing call on a batch of rays [2, n, 3]"""
        kw = {}
        if self.enc is not None:
            kw.update(encode_position_fn=self.enc, encode_direction_fn=self.encd)
        return nv.run_one_iter_of_nerf(self.H, self.W, self.focal, self.mc, self.mf, batch, self.opt, self.scene_id(),
                                       "validation", scene_config=self.scfg, **kw)


WORKLOADS = {
    "cfg1": dict(H=100, W=100, nc=64, nf=0, workload="cfg1_100x100_64+0_planes200",
                 desc="BASELINE configs[0]: 100x100 view, 64 coarse samples, no fine pass"),
    "cfg2": dict(H=RES, W=RES, nc=NC, nf=NF, workload="cfg2_800x800_64+128_planes200",
                 desc="BASELINE configs[1]: 800x800, 64 coarse + 128 fine samples (the config the metric is quoted on)"),
    "cfg3a": dict(H=800, W=800, nc=64, nf=128, workload="cfg3a_800x800_64+128_SRplanes_200->800",
                  desc="BASELINE configs[2], SR half: the fine model reads 4x super-resolved 800^2 planes (EDSR, once per scene)"),
    "cfg3b": dict(H=800, W=800, nc=64, nf=128, workload="cfg3b_800x800_65+129_mip_IPE", flop_per_eval=160768,
                  desc="BASELINE configs[2], mip half: IPE + FlexibleNeRFModel (the reference cannot combine IPE with planes)"),
    "cfg4": dict(H=756, W=1008, nc=128, nf=256, workload="cfg4_1008x756_128+256_ndc",
                 desc="BASELINE configs[3]: LLFF-shaped forward-facing scene, NDC rays, 128 + 256 samples"),
    "cfg5": dict(H=800, W=800, nc=64, nf=128, workload="cfg5_8scenes_one_decoder_800x800_64+128",
                 desc="BASELINE configs[4]: 8 scenes' planes behind one decoder pair, frames rendered scene after scene"),
}


def build_workload(name, device):
    """Synthetic scene of a BASELINE config on `device` ('cpu' for the oracle arm).  Seeds as SURVEY.md §8d."""
    from nvsr_b200 import scene
    spec = WORKLOADS[name]
    w = Workload(name, **spec)
    dev = torch.device(device)
    if name == "cfg3b":
        w.mc, w.mf = scene.make_mip_models(seed=0, device=dev)
        w.sid = "synth_DS2"
        w.opt = scene.render_options(w.nc, w.nf, mip=True)
        if dev.type == "cpu":
            from oracle import nvsr_oracle as O
            w.enc = lambda mc_: O.integrated_pos_enc(mc_[0], mc_[1], 7)
            w.encd = lambda x: O.positional_encoding(x, 4, True)
        else:
            import nvsr_b200
            w.enc, w.encd = nvsr_b200.IntegratedPositionalEncoding(3, 7), object()
    else:
        w.mc, w.mf, w.sid = scene.make_synthetic_scene(plane_res=PLANE_RES, view_res=32, seed=0, device=dev,
                                                       sr_scale=4 if name == "cfg3a" else None)
        w.opt = scene.render_options(w.nc, w.nf)
        if name == "cfg3a":
            w.offset = (2 - 1) / (2 * 2)     # downsampling_offset of a DS2 scene (train_nerf.py:610)
        if name == "cfg5":
            w.sids = [w.sid] + [scene.add_synthetic_scene(w.mc, w.mf, "s%d_DS2_PlRes%d_32" % (i, PLANE_RES), plane_res=PLANE_RES,
                                                          view_res=32, seed=10 + i) for i in range(1, 8)]
    if name == "cfg4":
        w.pose, w.focal = torch.eye(4), 0.8 * w.W
        w.scfg = scene.scene_cfg(near=0.0, far=1.0, no_ndc=False)
    else:
        w.pose, w.focal = scene.blender_camera(w.W)
        w.scfg = scene.scene_cfg(2.0, 6.0, True)
    w.pose = w.pose.to(dev)
    return w


def cpu_sample_rays(w, n_side):
    """a bounded sample of the SAME workload: an n_side x n_side lattice of the frame's rays (the whole frame when
    n_side is None)"""
    from oracle import nvsr_oracle as O
    ro, rd = O.get_ray_bundle(w.H, w.W, w.focal, w.pose.cpu(), 0, w.offset)
    if n_side is not None:
        iy = torch.linspace(0, w.H - 1, n_side).round().long()
        ix = torch.linspace(0, w.W - 1, n_side).round().long()
        ro, rd = ro[iy][:, ix], rd[iy][:, ix]
    return torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)


def time_cpu_oracle(n_side, steps, warmup, config="cfg2"):
    """the reference algorithm's CPU path (oracle port, torch CPU ops on all host threads): `steps` timed passes over a
    bounded ray sample of `config`'s frame after `warmup`; value = rays / MEDIAN pass time"""
    from oracle import nvsr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    w = build_workload(config, "cpu")
    batch = cpu_sample_rays(w, n_side)
    n = batch.shape[1]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.run_one_iter_of_nerf(w.H, w.W, w.focal, w.mc, w.mf, batch, w.opt, w.scene_id(False), "validation",
                                   encode_position_fn=w.enc, encode_direction_fn=w.encd, scene_config=w.scfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    med = statistics.median(times)
    what = "the whole %dx%d frame" % (w.W, w.H) if n_side is None else \
        "%dx%d lattice of the %dx%d frame's rays" % (n_side, n_side, w.W, w.H)
    return dict(rays_per_s=n / med, evals_per_s=n * w.evals_per_ray / med, ms_per_step=1e3 * med, rays=n,
                cores=torch.get_num_threads(), cpu_model=_cpu_model(), workload=w.workload, metric=w.metric(),
                evals_per_ray=w.evals_per_ray,
                sample=f"{what} ({n} rays, {w.nc}+{w.nf} samples), median of {len(times)} timed passes after {warmup} warm-up")


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def run_reference(args, rank, guard):
    if rank != 0:
        return
    # each step = a bounded ray sample of the workload's frame (cfg1: the whole 100x100 frame, it is the CPU-sized config)
    # (64 x 64 = 4 096 rays: ~1.7 s per pass on the box's 16 cores, so the driver's 20 + 5 steps end within a minute)
    n_side = None if args.config == "cfg1" else 64
    r = time_cpu_oracle(n_side=n_side, steps=max(1, args.steps), warmup=min(args.warmup, 1), config=args.config)
    line = {
        "impl": "reference", "metric": r["metric"], "value": r["rays_per_s"],
        "unit": "rays/s", "samples_per_s": r["evals_per_s"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": r["workload"], "rays_per_step": r["rays"],
                   "sample": "each step renders a bounded ray sample of the workload's frame (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    guard.emit(json.dumps(line))


def time_torch_gpu_frame(w, dev, frames=1, chunk=16384):
    """INFORMATIVE second baseline: cfg2's frame with stock PyTorch ops on this GPU — F.grid_sample,
    nn.Linear (fp32), cumprod, searchsorted, sort — i.e. what the reference's own code executes on a CUDA device, in
    ray chunks like its `chunksize`.  Not the metric's reference arm (that is the CPU path); it says what the kernels
    displace on this very GPU."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from bench_train_step import torch_step
    import nvsr_b200
    mc, mf, sid = w.mc, w.mf, w.sid
    saved = [dict(m.box_coords) for m in (mc, mf)]
    try:
        for m in (mc, mf):
            m.box_coords = {k: v.to(dev) for k, v in m.box_coords.items()}
        ro, rd = nvsr_b200.get_ray_bundle(w.H, w.W, w.focal, w.pose)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        t = torch.linspace(0.0, 1.0, w.nc).to(dev)
        z_row = 2.0 * (1.0 - t) + 6.0 * t
        u_row = torch.linspace(0.0, 1.0, w.nf).to(dev)

        def frame():
            for i in range(0, ro.shape[0], chunk):
                n = min(chunk, ro.shape[0] - i)
                torch_step(mc, mf, sid, ro[i:i + n], rd[i:i + n], vd[i:i + n], z_row.expand(n, w.nc).contiguous(),
                           u_row.expand(n, w.nf).contiguous(), False)

        frame()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            frame()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / frames
    finally:
        for m, b in zip((mc, mf), saved):
            m.box_coords = b
    return {"ms_per_step": ms, "value": w.rays / ms * 1e3, "unit": "rays/s", "chunk_rays": chunk, "frames": frames,
            "what": "stock PyTorch ops on this GPU (F.grid_sample, fp32 nn.Linear, cumprod, searchsorted, sort): the "
                    "reference's own op sequence; informative, not the reference arm"}


class StdoutGuard:
    """Everything libraries print to fd 1 during the run (e.g. the 'NCCL version ...' banner) is sent to
    stderr; only `emit()` writes to the real stdout, so rank 0 prints exactly ONE JSON line there."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config that is the headline workload (default: cfg2, the one the metric is quoted on)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32", "fp16-split"])
    ap.add_argument("--ray-chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the configs / precision_modes / torch_gpu_baseline blocks (profiling runs)")
    ap.add_argument("--sparse", action="store_true",
                    help="headline with the exact sparse colour path (default headline: every sample through both "
                         "decoders; the other mode is always reported beside it)")
    ap.add_argument("--dense", action="store_true", help="(default) headline with every sample through both decoders")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, guard)
        return
    args.warmup = max(args.warmup, 3)

    # Watchdog: a collective mismatch or a stuck rank must end as a failed run, never as a hung box.
    limit_s = float(os.environ.get("NVSR_BENCH_WATCHDOG_S", "420" if world > 1 else "900"))

    def _watchdog():
        time.sleep(limit_s)
        sys.stderr.write(f"bench.py: rank {rank} still running after {limit_s:.0f} s - aborting\n")
        sys.stderr.flush()
        os._exit(3)

    threading.Thread(target=_watchdog, daemon=True).start()

    import nvsr_b200
    from nvsr_b200 import ops
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    nvsr_b200.set_precision(args.precision)
    sparse = bool(args.sparse) and not args.dense and args.precision != "fp32"   # mode of the headline
    nvsr_b200.set_sparse_rgb(sparse)
    if args.ray_chunk:
        nvsr_b200.set_ray_chunk(args.ray_chunk)
    wl = build_workload(args.config, dev)
    H, W = wl.H, wl.W

    # row-band sharding: rank r renders rows [r0, r1); one all_gather of the result tiles per frame
    from nvsr_b200 import sharding
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    per_step = {}

    def timed(fn, steps, tag=None):
        """K steps bracketed by barrier + synchronize, device time from CUDA events, MAX over ranks.  With `tag`, an
        event is also recorded after every step (no synchronisation added) and this rank's per-step median / min are
        kept in per_step[tag] (SURVEY.md §8d asks for median + min next to the mean)."""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1 if tag else 2)]
        ev[0].record()
        for i in range(steps):
            fn()
            if tag:
                ev[i + 1].record()
        if not tag:
            ev[1].record()
        barrier()
        ms = torch.tensor([ev[0].elapsed_time(ev[-1])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if tag:
            each = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
            per_step[tag] = {"median_ms": statistics.median(each), "min_ms": each[0], "max_ms": each[-1], "rank": rank}
        return float(ms) / steps

    class Runner:
        """frames of one workload on this rank's row band: device-resident and through host buffers"""

        def __init__(self, w):
            self.w = w
            self.sh = sharding.FrameSharder(w.H, w.W, rank, world, dev)
            self.host_rays = self.host_out = None

        def frame_device(self):
            """rays generated on the device (get_ray_bundle kernel), everything resident"""
            flush.zero_()
            return self.sh.render(lambda a, b: self.w.render(nvsr_b200, a, b))

        def prepare_e2e(self):
            # host buffers for the e2e leg (the call a user of the reference makes: rays in, maps out)
            sh, w = self.sh, self.w
            with torch.no_grad():
                ro, rd = nvsr_b200.get_ray_bundle(w.H, w.W, w.focal, w.pose, row_range=(sh.r0, sh.r1))
            self.host_rays = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0).cpu().pin_memory()
            self.host_out = torch.empty((sh.per * w.W, 10), dtype=torch.float32).pin_memory()

        def frame_e2e(self):
            sh, w = self.sh, self.w
            flush.zero_()
            batch = self.host_rays.to(dev, non_blocking=True)
            res = sh.render(lambda a, b: w.run_rays(nvsr_b200, batch))
            self.host_out.copy_(res[rank * sh.per * w.W:(rank + 1) * sh.per * w.W] if world > 1 else res, non_blocking=True)

    run = Runner(wl)
    run.prepare_e2e()
    frame_device, frame_e2e = run.frame_device, run.frame_e2e
    n_local = run.sh.n_local

    with torch.no_grad():
        for _ in range(args.warmup):
            frame_device()
        clocks = ClockSampler(local)
        clocks.start()
        ops.LAUNCHES.clear()
        ms_dev = timed(frame_device, args.steps, tag="device")
        launches = sum(ops.LAUNCHES.values())
        clk = clocks.stop()
        # A short timed region (many GPUs, few steps) gives nvidia-smi too few samples: then the clocks are sampled
        # over a longer, untimed stretch of the same step.  Every rank must run the SAME number of frames (each
        # frame holds a collective), so both the decision and the count derive from ms_dev — the all-reduced
        # maximum, identical on every rank — never from a rank's own sampler or wall clock.
        if ms_dev * args.steps < 250.0:
            n_extra = max(1, int(math.ceil(600.0 / max(ms_dev, 1e-3))))
            clocks = ClockSampler(local)
            clocks.start()
            for _ in range(n_extra):
                frame_device()
            barrier()
            clk = clocks.stop()
            clk["note"] = f"timed region shorter than 0.25 s; sampled over {n_extra} more frames of the same step right after it"
        for _ in range(2):
            frame_e2e()
        ms_e2e = timed(frame_e2e, args.steps)

        # per-kernel live timing (CUDA events around every launch of ours, same stream) for the roofline
        ops.PROFILE = []
        barrier()
        for _ in range(min(3, args.steps)):
            frame_device()
        barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        # companion: the same frame in the OTHER mode (sparse colour path <-> every sample through both decoders)
        other = None
        if args.precision != "fp32" and args.config != "cfg3b":
            nvsr_b200.set_sparse_rgb(not sparse)
            for _ in range(2):
                frame_device()
            ms_o = timed(frame_device, min(3, args.steps))
            frame_e2e()
            ms_o_e2e = timed(frame_e2e, min(3, args.steps))
            other = {"ms_per_step": ms_o, "value": wl.rays / (ms_o * 1e-3), "unit": "rays/s",
                     "e2e": {"value": wl.rays / (ms_o_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_o_e2e},
                     "samples_per_s": wl.rays * wl.evals_per_ray / (ms_o * 1e-3)}
            if not sparse:
                # the companion is the sparse path: how many samples reached the rgb decoder (counted on the device)
                ops.PROFILE = []
                barrier()
                frame_device()
                barrier()
                p2, ops.PROFILE = ops.PROFILE, None
                lit = sum(int(m["count"].item()) for n_, _, _, m in p2 if n_ == "nvsr_mlp_chain" and m.get("count") is not None)
                cap = sum(m["rows"] for n_, _, _, m in p2 if n_ == "nvsr_mlp_chain" and m.get("count") is not None)
                other["rgb_rows_evaluated"] = (lit / cap) if cap else None
                other["note"] = ("same frame with the exact sparse colour path (rgb decoder only where sigma + noise > 0; "
                                 "maps bit-identical, DESIGN.md 4.5)")
            else:
                other["note"] = "same frame with every sample through both decoders; maps bit-identical"
            nvsr_b200.set_sparse_rgb(sparse)

        # ---- every BASELINE config at full size, same timing rules (1 warm-up + 3 timed frames each; the frame
        # count is fixed, so every rank issues the same collectives)
        configs = None
        if not args.no_extras:
            configs = {}
            for name in sorted(WORKLOADS):
                if name == args.config:
                    configs[name] = {"workload": wl.workload, "ms_per_step": ms_dev, "value": wl.rays / (ms_dev * 1e-3),
                                     "unit": "rays/s", "samples_per_s": wl.rays * wl.evals_per_ray / (ms_dev * 1e-3),
                                     "note": "the headline of this line"}
                    continue
                w2 = build_workload(name, dev)
                r2 = Runner(w2)
                n_frames = len(w2.sids) if w2.sids else 3
                for _ in range(len(w2.sids) if w2.sids else 1):      # warm-up: packs planes / weights (every scene)
                    r2.frame_device()
                ms2 = timed(r2.frame_device, n_frames)
                configs[name] = {"workload": w2.workload, "what": w2.desc, "ms_per_step": ms2, "frames_timed": n_frames,
                                 "value": w2.rays / (ms2 * 1e-3), "unit": "rays/s", "rays_per_step": w2.rays,
                                 "samples_per_s": w2.rays * w2.evals_per_ray / (ms2 * 1e-3),
                                 "tflops": w2.rays * w2.evals_per_ray * w2.flop_per_eval / (ms2 * 1e-3) / 1e12,
                                 "sparse_rgb": bool(sparse) and name != "cfg3b"}
                if name == "cfg3a":
                    # "SR inference ... timed separately, once per scene": the three position planes through
                    # the EDSR conv chain (cuDNN, channels-last, the frame's 16-bit type) + nvsr_sr_finalize writing the
                    # gather's plane images — device-resident, never re-uploaded (the reference caches the SR plane on
                    # the CPU and re-uploads it for every network chunk, models.py:893,925)
                    try:
                        from nvsr_b200 import sr as SR
                    except ImportError:          # (the control-flow tests run this file against a stand-in package)
                        SR = None
                    if SR is not None and getattr(w2.mf, "SR_model", None) is not None:
                        res = SR.resolver_of(w2.mf.SR_model)
                        lr_names = list(w2.mf.SR_model.LR_planes)
                        code = {"fp16": nvsr_b200.NVSR_F16, "bf16": nvsr_b200.NVSR_BF16}.get(args.precision, nvsr_b200.NVSR_F16)

                        def sr_scene():
                            res.cache.clear()
                            for nm in lr_names:
                                res.super_resolve(nm, code)
                        sr_scene()
                        configs[name]["sr_inference_ms_per_scene"] = timed(sr_scene, 2)
                        configs[name]["sr_model"] = ("EDSR 32 blocks x 256 channels, x4 (config/TrainModels.yml:174-184), "
                                                     "3 planes 200^2 -> 800^2")
                del w2, r2
                nvsr_b200.render.clear_caches()
                torch.cuda.empty_cache()
        # ---- N = 1 only: the fp32 1e-3-parity mode's frame time and the stock-PyTorch frame on this GPU
        precision_modes = torch_gpu = None
        if world == 1 and not args.no_extras and args.precision != "fp32" and args.config != "cfg3b":
            nvsr_b200.set_precision("fp32")
            frame_device()
            ms32 = timed(frame_device, 2)
            ms_split = None
            if args.config != "cfg3b":
                nvsr_b200.set_precision("fp16-split")
                frame_device()
                ms_split = timed(frame_device, 3)
            nvsr_b200.set_precision(args.precision)
            precision_modes = {
                args.precision: {"ms_per_step": ms_dev, "value": wl.rays / (ms_dev * 1e-3), "unit": "rays/s",
                                 "contract": "tcgen05 decoder, 16-bit operands, fp32 accumulate: stated per-mode bounds "
                                             "(DESIGN.md section 2)"},
                "fp32": {"ms_per_step": ms32, "value": wl.rays / (ms32 * 1e-3), "unit": "rays/s", "frames_timed": 2,
                         "contract": "SIMT fp32 decoder, fp32 planes and features: the 1e-3 parity mode"}}
            if ms_split is not None:
                precision_modes["fp16-split"] = {
                    "ms_per_step": ms_split, "value": wl.rays / (ms_split * 1e-3), "unit": "rays/s", "frames_timed": 3,
                    "contract": "tcgen05 decoder; the density chain with split fp16 operands (3 MMA passes per layer) on "
                                "features interpolated in fp32 from the planes as fp16 hi + lo halves, the colour chain in fp16: meets the 1e-3 map contract "
                                "(tests/parity_attribution.py, mode 'fp16-split')"}
            if args.config == "cfg2":
                torch_gpu = time_torch_gpu_frame(wl, dev)
    agg = {}
    for name, a, b, meta in prof:
        key = name
        rows_done = meta.get("rows", 0)
        if meta.get("count") is not None:           # sparse launch: the rows it evaluated are counted on the device
            rows_done = int(meta["count"].item())
        if name == "nvsr_mlp_chain":
            key = "mlp_density" if meta["flops_per_row"] < 120000 else "mlp_rgb"
            if args.config == "cfg3b":
                key = "mlp_mip"
        d = agg.setdefault(key, dict(ms=0.0, n=0, bytes=0, flops=0, rows=0, rows_cap=0))
        d["ms"] += a.elapsed_time(b)
        d["n"] += 1
        d["rows"] += rows_done
        d["rows_cap"] += meta.get("rows", rows_done)
        d["bytes"] += rows_done * meta["bytes_per_row"] if "bytes_per_row" in meta else meta.get("bytes", 0)
        d["flops"] += rows_done * meta["flops_per_row"] if "flops_per_row" in meta else meta.get("flops", 0)
    pk = peaks()
    total_ms = sum(d["ms"] for d in agg.values())
    kernels = {}
    for k, d in agg.items():
        e = {"launches": d["n"], "avg_ms": d["ms"] / d["n"], "share": d["ms"] / total_ms}
        if d["flops"]:
            e["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
            e["frac_tensor_peak"] = e["tflops"] / pk["tf_sust"]
        if d["bytes"]:
            e["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            e["frac_hbm_peak"] = e["gbs"] / pk["hbm"]
        kernels[k] = e
    rgb_frac = (agg["mlp_rgb"]["rows"] / max(agg["mlp_rgb"]["rows_cap"], 1)) if "mlp_rgb" in agg else None
    mlp = [agg[k] for k in ("mlp_rgb", "mlp_density", "mlp_mip") if k in agg]
    mlp_ms = sum(d["ms"] for d in mlp)
    mlp_fl = sum(d["flops"] for d in mlp)
    mlp_n = sum(d["n"] for d in mlp)
    roofline = None
    if mlp_ms > 0 and args.precision != "fp32":
        ach = mlp_fl / (mlp_ms * 1e-3) / 1e12
        roofline = {"kernel": "mlp_chain_tc_kernel (decoder, tcgen05)", "bound": "tensor", "achieved": ach,
                    "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                    "traffic": ncu_traffic("mlp_chain_tc_kernel", mlp_fl / mlp_n / (wl.flop_per_eval / 2.0)),
                    "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "avg_launch_ms": mlp_ms / mlp_n, "flop_per_launch": mlp_fl / mlp_n, "share_of_step": mlp_ms / total_ms}
    elif mlp_ms > 0:
        ach = mlp_fl / (mlp_ms * 1e-3) / 1e12
        roofline = {"kernel": "mlp_chain_f32_kernel (decoder, SIMT fp32 parity mode)", "bound": "tensor", "achieved": ach,
                    "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "traffic": None,
                    "peak_source": pk["src"], "share_of_step": mlp_ms / total_ms}

    train_step = None
    if rank == 0 and world == 1 and not args.no_extras and args.config == "cfg2" and args.precision == "fp16":
        # SURVEY section 8f rank 1 (informative, not the metric): the training iteration of the same path, ours vs stock PyTorch
        try:
            import argparse as _ap
            import importlib.util as _iu
            _spec = _iu.spec_from_file_location("bench_train_step", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                 "scripts", "bench_train_step.py"))
            _mod = _iu.module_from_spec(_spec)
            _spec.loader.exec_module(_mod)
            train_step = _mod.measure(_ap.Namespace(rays=4096, steps=10, warmup=3, plane_res=200, no_torch=False, quick=True))
        except Exception as e:   # the block is informative: never lose the bench line to it
            train_step = {"error": repr(e)[:300]}
        finally:
            torch.cuda.empty_cache()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            # a bounded sample of the SAME workload: a 100 x 100 lattice of the frame's rays (cfg1: its whole frame)
            r = time_cpu_oracle(n_side=None if args.config == "cfg1" else 100, steps=3, warmup=1, config=args.config)
            cpu = {"value": r["rays_per_s"], "unit": "rays/s", "evals_per_s": r["evals_per_s"], "cores": r["cores"],
                   "cpu_model": r["cpu_model"], "kind": "port", "sample": r["sample"]}
            if args.config != "cfg1" and not args.no_extras:
                # BASELINE.md section 4, exactly: config 1 (100x100, 64 coarse, no fine), 1 warm-up + 3 timed frames, median
                r1 = time_cpu_oracle(n_side=None, steps=3, warmup=1, config="cfg1")
                cpu["cfg1"] = {"value": r1["rays_per_s"], "unit": "rays/s", "evals_per_s": r1["evals_per_s"],
                               "ms_per_frame": r1["ms_per_step"], "sample": r1["sample"]}
        rays = wl.rays
        line = {
            "metric": wl.metric(),
            "value": rays / (ms_dev * 1e-3), "unit": "rays/s",
            # nominal decoder evaluations of the reference per second: rays/s x (Nc + (Nc + Nf)); with sparse_rgb the
            # rgb chain is evaluated only on `rgb_rows_evaluated` of them (the others have weight exactly 0)
            "samples_per_s": rays * wl.evals_per_ray / (ms_dev * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "per_step": per_step.get("device"),   # rank 0's own per-frame median / min / max over the timed steps
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16 operands, f32 accumulate", "bf16": "bf16 operands, f32 accumulate", "fp32": "f32",
                      "fp16-split": "f16 operands (density chain: split hi+lo, 3 passes), f32 accumulate"}[args.precision],
            "data": "synthetic",
            "config": {"workload": wl.workload, "what": wl.desc, "rays_per_step": rays,
                       "planes": "3x48x200^2 + 48x32^2" if args.config != "cfg3b" else None,
                       "decoder": "48->128x4->1 + 192->128x4->3 (coarse+fine)" if args.config != "cfg3b"
                       else "FlexibleNeRFModel 36->128x4 (+27 dir) (coarse+fine)", "sharding": f"{world} row bands",
                       "ray_chunk": nvsr_b200.render._state["ray_chunk"],
                       "sparse_rgb": bool(sparse), "rgb_rows_evaluated": rgb_frac,
                       "l2": "256 MiB buffer rewritten before every step; per-step intermediates (>60 GB) exceed L2"},
            "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(run.host_rays.numel() * 4), "d2h_bytes_per_step": int(n_local * 10 * 4)},
            ("dense" if sparse else "sparse"): other,
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "configs": configs,
            "precision_modes": precision_modes,
            "torch_gpu_baseline": torch_gpu,
            "train_step": train_step,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
