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


def build_scene(device):
    import nvsr_b200
    from nvsr_b200 import scene
    mc, mf, sid = scene.make_synthetic_scene(plane_res=PLANE_RES, view_res=32, seed=0, device=device)
    pose, focal = scene.blender_camera(RES)
    return mc, mf, sid, pose, focal, scene.render_options(NC, NF), scene.scene_cfg(2.0, 6.0, True)


def cpu_sample_rays(pose, focal, n_side):
    """a bounded sample of the SAME workload: an n_side x n_side lattice of the 800x800 frame's rays"""
    from oracle import nvsr_oracle as O
    ro, rd = O.get_ray_bundle(RES, RES, focal, pose)
    idx = torch.linspace(0, RES - 1, n_side).round().long()
    ro, rd = ro[idx][:, idx], rd[idx][:, idx]
    return torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)


def time_cpu_oracle(n_side, steps, warmup):
    """the reference algorithm's CPU path (oracle port, torch CPU ops on all host threads)"""
    from oracle import nvsr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    mc, mf, sid, pose, focal, opt, scfg = build_scene("cpu")
    batch = cpu_sample_rays(pose, focal, n_side)
    n = batch.shape[1]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.run_one_iter_of_nerf(RES, RES, focal, mc, mf, batch, opt, sid, "validation", scene_config=scfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return dict(rays_per_s=n * len(times) / total, ms_per_step=1e3 * total / len(times), rays=n,
                cores=torch.get_num_threads(),
                sample=f"{n_side}x{n_side} lattice of the 800x800 frame's rays ({n} rays, 64+128 samples, planes 200^2), "
                       f"{len(times)} timed passes after {warmup} warm-up")


def run_reference(args, rank, guard):
    if rank != 0:
        return
    r = time_cpu_oracle(n_side=48, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "rays/s (800x800 render, 64 coarse + 128 fine samples/ray)", "value": r["rays_per_s"],
        "unit": "rays/s", "samples_per_s": r["rays_per_s"] * EVALS_PER_RAY, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2_800x800_64+128_planes200", "rays_per_step": r["rays"],
                   "sample": "each step renders a bounded ray sample of the workload's frame (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    guard.emit(json.dumps(line))


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
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--ray-chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    limit_s = float(os.environ.get("NVSR_BENCH_WATCHDOG_S", "300" if world > 1 else "900"))

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
    mc, mf, sid, pose, focal, opt, scfg = build_scene(dev)
    pose = pose.to(dev)

    # row-band sharding: rank r renders rows [r0, r1); one all_gather of the result tiles per frame
    from nvsr_b200 import sharding
    sh = sharding.FrameSharder(RES, RES, rank, world, dev)
    r0, r1, rows_per, n_local = sh.r0, sh.r1, sh.per, sh.n_local
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def frame_device():
        """rays generated on the device (get_ray_bundle kernel), everything resident"""
        flush.zero_()
        return sh.render(lambda a, b: nvsr_b200.render_frame(RES, RES, focal, pose, mc, mf, opt, sid, scfg, row_range=(a, b)))

    # host buffers for the e2e leg (the call a user of the reference makes: rays in, maps out)
    with torch.no_grad():
        ro, rd = nvsr_b200.get_ray_bundle(RES, RES, focal, pose, row_range=(r0, r1))
    host_rays = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0).cpu().pin_memory()
    host_out = torch.empty((rows_per * RES, 10), dtype=torch.float32).pin_memory()

    def frame_e2e():
        flush.zero_()
        batch = host_rays.to(dev, non_blocking=True)
        res = sh.render(lambda a, b: nvsr_b200.run_one_iter_of_nerf(RES, RES, focal, mc, mf, batch, opt, sid, "validation",
                                                                    scene_config=scfg))
        host_out.copy_(res[rank * rows_per * RES:(rank + 1) * rows_per * RES] if world > 1 else res, non_blocking=True)

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
        if args.precision != "fp32":
            nvsr_b200.set_sparse_rgb(not sparse)
            for _ in range(2):
                frame_device()
            ms_o = timed(frame_device, min(3, args.steps))
            frame_e2e()
            ms_o_e2e = timed(frame_e2e, min(3, args.steps))
            other = {"ms_per_step": ms_o, "value": RES * RES / (ms_o * 1e-3), "unit": "rays/s",
                     "e2e": {"value": RES * RES / (ms_o_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_o_e2e},
                     "samples_per_s": RES * RES * EVALS_PER_RAY / (ms_o * 1e-3)}
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
    agg = {}
    for name, a, b, meta in prof:
        key = name
        rows_done = meta.get("rows", 0)
        if meta.get("count") is not None:           # sparse launch: the rows it evaluated are counted on the device
            rows_done = int(meta["count"].item())
        if name == "nvsr_mlp_chain":
            key = "mlp_density" if meta["flops_per_row"] < 120000 else "mlp_rgb"
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
    mlp = [agg[k] for k in ("mlp_rgb", "mlp_density") if k in agg]
    mlp_ms = sum(d["ms"] for d in mlp)
    mlp_fl = sum(d["flops"] for d in mlp)
    mlp_n = sum(d["n"] for d in mlp)
    roofline = None
    if mlp_ms > 0 and args.precision != "fp32":
        ach = mlp_fl / (mlp_ms * 1e-3) / 1e12
        roofline = {"kernel": "mlp_chain_tc_kernel (decoder, tcgen05)", "bound": "tensor", "achieved": ach,
                    "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                    "traffic": ncu_traffic("mlp_chain_tc_kernel", mlp_fl / mlp_n / (FLOP_PER_EVAL / 2.0)),
                    "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "avg_launch_ms": mlp_ms / mlp_n, "flop_per_launch": mlp_fl / mlp_n, "share_of_step": mlp_ms / total_ms}
    elif mlp_ms > 0:
        ach = mlp_fl / (mlp_ms * 1e-3) / 1e12
        roofline = {"kernel": "mlp_chain_f32_kernel (decoder, SIMT fp32 parity mode)", "bound": "tensor", "achieved": ach,
                    "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "traffic": None,
                    "peak_source": pk["src"], "share_of_step": mlp_ms / total_ms}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            r = time_cpu_oracle(n_side=64, steps=2, warmup=1)
            cpu = {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        rays = RES * RES
        line = {
            "metric": "rays/s (800x800 render, 64 coarse + 128 fine samples/ray)",
            "value": rays / (ms_dev * 1e-3), "unit": "rays/s",
            # nominal decoder evaluations of the reference per second: rays/s x (Nc + (Nc + Nf)); with sparse_rgb the
            # rgb chain is evaluated only on `rgb_rows_evaluated` of them (the others have weight exactly 0)
            "samples_per_s": rays * EVALS_PER_RAY / (ms_dev * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "per_step": per_step.get("device"),   # rank 0's own per-frame median / min / max over the timed steps
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16 operands, f32 accumulate", "bf16": "bf16 operands, f32 accumulate", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "cfg2_800x800_64+128_planes200", "rays_per_step": rays, "planes": "3x48x200^2 + 48x32^2",
                       "decoder": "48->128x4->1 + 192->128x4->3 (coarse+fine)", "sharding": f"{world} row bands",
                       "ray_chunk": nvsr_b200.render._state["ray_chunk"],
                       "sparse_rgb": bool(sparse), "rgb_rows_evaluated": rgb_frac,
                       "l2": "256 MiB buffer rewritten before every step; per-step intermediates (>60 GB) exceed L2"},
            "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(host_rays.numel() * 4), "d2h_bytes_per_step": int(n_local * 10 * 4)},
            ("dense" if sparse else "sparse"): other,
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
