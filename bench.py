#!/usr/bin/env python
"""bench.py — frames/sec of the ElasticFusion hot path (track + fuse + predict) on B200, with the ICP-reduction roofline.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one ElasticFusion::processFrame call on one 640x480 frame of the synthetic ICL-NUIM-shaped room sequence
(BASELINE.json configs[1]); every rank (one per GPU) tracks and fuses its own independent sequence, so `value` is the
whole-job frames/sec (weak scaling, no data-path collective; NCCL is used for the two barriers and the max-over-ranks).

  value    frames/sec with the K frames already resident in HBM (ef_process_frame_device), CUDA-event timed per frame
           on the context's stream, L2 flushed between frames outside the timed spans.
  e2e      the same frames through ef_process_frame: HOST buffers in, host->device copies and the device->host read
           of the pose inside the timed region (what a caller of libefusion.so sees).
  roofline the ICP residual+Jacobian+6x6 reduction at level 0 (north_star's kernel): algorithmic 48 B/pixel + 116 B,
           CUDA-event duration with L2 flushed before every launch, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (a port of the reference algorithm, all host cores) on a bounded sample of the workload.

--impl reference times the reference arm: the reference's own CUDA tracking kernels compiled unmodified into oracle/_ref
(driven launch-for-launch like Core/Utils/RGBDOdometry.cpp) plus the CPU oracle for the GLSL mapping half, which cannot
run without OpenGL.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BIG = 2147483647 // 2
CLOCK_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, device_index: int):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={CLOCK_Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_frames(K, n, seed):
    """First n frames of the synthetic sequence (frame i depends only on (seed, i), so a longer cached run serves any
    shorter request: the reference arm and repeated runs on one box do not render again)."""
    from elasticfusion_b200 import synth

    cache = os.path.join(os.environ.get("EF_BENCH_CACHE", "/tmp"), f"ef_bench_{K.width}x{K.height}_{seed}.npz")
    have_rgb = have_depth = None
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            have_rgb, have_depth = z["rgb"], z["depth"]
        except Exception:
            have_rgb = have_depth = None
    m = 0 if have_rgb is None else len(have_rgb)
    if m >= n:
        return have_rgb[:n], have_depth[:n]
    rgb = np.empty((n, K.height, K.width, 3), np.uint8)
    depth = np.empty((n, K.height, K.width), np.uint16)
    if m:
        rgb[:m], depth[:m] = have_rgb, have_depth
    traj = synth.trajectory(n, seed=seed)
    for i in range(m, n):
        rgb[i], depth[i] = synth.render(traj[i], K, noise_seed=seed * 100003 + i)[:2]
    try:
        tmp = cache + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, rgb=rgb, depth=depth)
        os.replace(tmp, cache)
    except Exception:
        pass
    return rgb, depth


def workload(args):
    from elasticfusion_b200 import synth

    if args.workload == "640x480":
        return synth.K_DEFAULT, 5_000_000, "640x480 synthetic planar-room sequence (ICL-NUIM-shaped), full track+fuse, 5M surfel cap"
    if args.workload == "1280x960":
        return synth.K_DEFAULT.scaled(2), 20_000_000, "1280x960 high-res synthetic sequence, full track+fuse, 20M surfel cap"
    raise SystemExit("unknown workload")


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, dist):
    import torch

    from elasticfusion_b200 import capi

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, cap, wl_name = workload(args)
    n_total = args.warmup + args.steps
    rgb, depth = make_frames(K, n_total + 1, 42 + rank)  # +1: the last timed frame still prefetches its successor
    stream = torch.cuda.Stream(device=dev)
    la = not args.no_lookahead
    cfg = capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=cap, time_delta=BIG, device=local)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize(dev)

    # ---------------- value: inputs resident in HBM ----------------
    ctx = capi.Context(cfg, stream=stream.cuda_stream)
    rgb_d = torch.from_numpy(rgb).to(dev)
    depth_d = torch.from_numpy(depth.view(np.int16)).to(dev)
    torch.cuda.synchronize(dev)
    def step_device(i):
        # look-ahead: frame i was staged by the previous step; its successor is staged while frame i is in flight
        if la:
            ctx.process_frame_device(None, None, i)
            ctx.prefetch_frame_device(rgb_d[i + 1].data_ptr(), depth_d[i + 1].data_ptr())
        else:
            ctx.process_frame_device(rgb_d[i].data_ptr(), depth_d[i].data_ptr(), i)

    with torch.cuda.stream(stream):
        if la:
            ctx.prefetch_frame_device(rgb_d[0].data_ptr(), depth_d[0].data_ptr())
        for i in range(args.warmup):
            step_device(i)
    ctx.sync()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    l0 = ctx.launch_count()
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            i = args.warmup + k
            if not args.no_flush:
                flush.fill_(k & 0xff)
            starts[k].record(stream)
            step_device(i)
            stops[k].record(stream)
    ctx.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - l0
    frame_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    dev_ms = float(sum(frame_ms))
    n_surfels = ctx.map_count()
    pose = ctx.get_pose()

    # ---------------- roofline: ICP reduction at level 0, cold L2 ----------------
    roof = icp_roofline(ctx, stream, flush, K, rgb, depth, local)
    ctx.close()
    del rgb_d, depth_d
    roof_hi = None
    if rank == 0 and args.workload == "640x480" and not args.no_hires_roofline:
        # the same kernel on BASELINE configs[2]'s image size (59 MB per launch): the size at which the pass is bandwidth-
        # rather than launch-latency-dominated
        from elasticfusion_b200 import synth

        Kh = synth.K_DEFAULT.scaled(2)
        rgb_h, depth_h = make_frames(Kh, 3, 42)
        ctx_h = capi.Context(capi.default_config(Kh.width, Kh.height, Kh.fx, Kh.fy, Kh.cx, Kh.cy, capacity=3_000_000, time_delta=BIG, device=local),
                             stream=stream.cuda_stream)
        for i in range(3):
            ctx_h.process_frame(rgb_h[i], depth_h[i], i)
        roof_hi = icp_roofline(ctx_h, stream, flush, Kh, rgb_h, depth_h, local)
        ctx_h.close()

    # ---------------- e2e: host buffers through the public call ----------------
    ctx2 = capi.Context(cfg, stream=stream.cuda_stream)

    def step_host(i):
        # the public per-frame call sequence with host buffers: consume the staged frame, stage the next one (pinned copy +
        # H2D + preprocess on the side stream) while the GPU works, then wait for the pose
        if la:
            ctx2.process_frame_device(None, None, i)
            ctx2.prefetch_frame(rgb[i + 1], depth[i + 1])
            ctx2.finish_frame()
        else:
            ctx2.process_frame(rgb[i], depth[i], i)

    if la:
        ctx2.prefetch_frame(rgb[0], depth[0])
    for i in range(args.warmup):
        step_host(i)
    barrier()
    e2e_s = 0.0
    for k in range(args.steps):
        i = args.warmup + k
        if not args.no_flush:
            flush.fill_(k & 0xff)
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_host(i)
        e2e_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()  # sampled across the value, roofline and e2e loops
    pose2 = ctx2.get_pose()
    ctx2.close()

    # max over ranks
    tt = torch.tensor([dev_ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = tt.tolist()
    if rank != 0:
        return
    hbm, peak_src = load_peaks()
    total_frames = args.steps * world
    out = {
        "metric": "frames/sec, full track+fuse+predict (processFrame)", "value": total_frames / (dev_ms_max / 1000.0), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "frames_per_gpu": args.steps, "surfels_at_end": int(n_surfels),
                   "l2": "flushed between frames (256 MiB write, outside the timed spans)" if not args.no_flush else "not flushed",
                   "parallelism": f"{world} independent sequences, one per GPU", "open_loop": True,
                   "lookahead": ("next frame's upload + depth preprocess + pyramids staged on a side stream inside the timed span of the "
                                 "frame in flight (ef_prefetch_frame)") if la else "off"},
        "e2e": {"value": total_frames / (e2e_ms_max / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": int(K.width * K.height * 5),
                "d2h_bytes_per_step": 132, "ms_per_step": e2e_ms_max / args.steps},
        "gpu_launches": int(launches), "launches_per_frame": launches / args.steps,
        "clocks": clocks, "roofline": dict(roof, peak=hbm, frac=roof["achieved"] / hbm, peak_source=peak_src),
        **({"roofline_1280x960": dict(roof_hi, peak=hbm, frac=roof_hi["achieved"] / hbm, peak_source=peak_src)} if roof_hi else {}),
        "frame_ms": {"median": statistics.median(frame_ms), "p10": float(np.percentile(frame_ms, 10)), "p90": float(np.percentile(frame_ms, 90))},
        "wall_s_value_loop": t_wall, "pose_check": float(np.abs(pose - pose2).max()),
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(K, rgb, depth, cap)
    print(json.dumps(out))


def icp_roofline(ctx, stream, flush, K, rgb, depth, local):
    """Duration of the dominant kernel (k_iter1: ICP residual + Jacobian + per-CTA 29-term reduction, one Gauss-Newton
    iteration of level 0).

    cold (the HBM-roofline number): the launch is repeated round-robin over R independent contexts whose level-0 maps
    together exceed the 126 MB L2 (R x 14.7 MB at 640x480, R x 59 MB at 1280x960), so every launch finds its inputs evicted;
    two CUDA events bracket the whole batch on the launching stream and the average per launch is reported -- inputs larger
    than L2, no flush inside the timed region, launch gaps included as in the frame loop. warm = the same batch on one
    context (L2 resident). single_launch_event_us = one launch between two events after a 256 MiB flush (adds the latency of
    an isolated launch + two event records, ~4 us)."""
    import torch

    from elasticfusion_b200 import capi

    T = ctx.get_pose()
    R = T[:3, :3].astype(np.float32)
    t = T[:3, 3].astype(np.float32)
    ctx.icp_step_async(0, R, t, np.linalg.inv(R).astype(np.float32), t)
    ctx.sync()
    nbytes = 48 * K.width * K.height + 116
    n_ctx = max(3, int(np.ceil(260e6 / nbytes)))  # level-0 maps in rotation: twice the 126 MB L2
    cfg = capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=400_000, time_delta=BIG, device=local)
    ring = []
    for j in range(n_ctx):
        c = capi.Context(cfg, stream=stream.cuda_stream)
        for i in range(2):
            c.process_frame(rgb[(i + j) % len(rgb)], depth[(i + j) % len(depth)], i)
        Tj = c.get_pose()
        Rj, tj = Tj[:3, :3].astype(np.float32), Tj[:3, 3].astype(np.float32)
        c.icp_step_async(0, Rj, tj, np.linalg.inv(Rj).astype(np.float32), tj)
        c.sync()
        ring.append(c)

    def batch(ctxs, rounds):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            for c in ctxs:  # untimed pass: instruction cache, TLBs
                c.icp_dense_pass_async(0)
            s.record(stream)
            for _ in range(rounds):
                for c in ctxs:
                    c.icp_dense_pass_async(0)
            e.record(stream)
        e.synchronize()
        return s.elapsed_time(e) * 1000.0 / (rounds * len(ctxs))

    cold = statistics.median([batch(ring, 4) for _ in range(5)])
    warm = statistics.median([batch([ctx], 4 * n_ctx) for _ in range(5)])
    for c in ring:
        c.close()

    def single(fn):
        ev = []
        with torch.cuda.stream(stream):
            for k in range(20):
                flush.fill_(k)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                fn()
                e.record(stream)
                ev.append((s, e))
        ctx.sync()
        return statistics.median([s.elapsed_time(e) * 1000.0 for s, e in ev])

    single_cold = single(lambda: ctx.icp_dense_pass_async(0))
    full_cold = single(lambda: ctx.icp_step_async(0))
    traffic = ncu_us = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(f"k_iter1_{K.width}x{K.height}")
            ncu_us = tj.get(f"k_iter1_{K.width}x{K.height}_ncu_duration_us")
        except Exception:
            traffic = None
    return {"kernel": "k_iter1 (ICP residual + Jacobian + per-CTA 29-term reduction, level 0; ef_reduce.cu)", "bound": "hbm", "unit": "GB/s",
            "achieved": nbytes / (cold * 1e-6) / 1e9, "algorithmic_bytes": nbytes, "duration_us": cold,
            "achieved_warm_l2": nbytes / (warm * 1e-6) / 1e9, "duration_warm_us": warm,
            "single_launch_event_us": single_cold, "complete_reduction_single_launch_event_us": full_cold,
            "traffic": traffic, "ncu_duration_us": ncu_us,
            "units_per_launch": f"{K.width * K.height} pixels (one Gauss-Newton iteration of pyramid level 0), 48 B each",
            "timing": (f"two CUDA events on the launching stream around a batch of 4 x {n_ctx} launches rotating over {n_ctx} contexts "
                       f"({n_ctx * nbytes / 1e6:.0f} MB of level-0 maps, twice the 126 MB L2, so every launch is L2-cold), average per launch, median of 5 batches")}


def cpu_baseline(K, rgb, depth, cap, seconds=20.0):
    """CPU oracle (port of the reference algorithm, OpenMP over all host cores) on the first frames of the same workload."""
    from oracle import ef_oracle as eo

    f = eo.Fusion(K, capacity=min(cap, 3_000_000))
    n, t0 = 0, time.perf_counter()
    while n < len(rgb) and (time.perf_counter() - t0 < seconds or n < 5):
        f.process_frame(rgb[n], depth[n], n)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": eo.get_threads(), "kind": "port",
            "sample": f"first {n} frames of the same sequence through the CPU oracle pipeline ({dt:.1f} s)", "stages_s": f.timers()}


# ------------------------------------------------------------------------------------------------------------------
def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota (oversubscribing OpenMP
    beyond that makes the spin-waiting teams collapse)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    return max(1, n)


def run_reference(args, rank, world):
    """Reference arm: reference CUDA tracking (oracle/_ref, unmodified kernels driven like RGBDOdometry.cpp) + CPU oracle
    for the GLSL mapping half. Falls back to the pure CPU oracle when oracle/_ref is absent."""
    if rank != 0:
        return
    pinned = "OMP_NUM_THREADS" in os.environ  # torchrun pins it to 1; the reference arm may use the usable host cores
    K, cap, wl_name = workload(args)
    n_total = args.warmup + args.steps
    budget_frames = min(n_total, 150)
    rgb, depth = make_frames(K, n_total, 42)
    try:
        from oracle import ef_ref

        have_ref = ef_ref.available()
    except Exception:
        have_ref = False
    from oracle import ef_oracle as eo

    if pinned:
        eo.set_threads(min(usable_cores(), 32))  # the oracle's row-parallel loops stop scaling (and can collapse) beyond this
    ncores = eo.get_threads()
    if have_ref:
        from oracle import ef_ref

        runner = ef_ref.HybridFusion(K, capacity=min(cap, 3_000_000))
        kind, sample = "reference", "reference CUDA tracking kernels (oracle/_ref) + CPU-oracle mapping (GL half cannot run here)"
    else:
        runner = eo.Fusion(K, capacity=min(cap, 3_000_000))
        kind, sample = "port", "CPU oracle pipeline (oracle/_ref absent)"
    w = min(args.warmup, budget_frames // 4)
    k = min(args.steps, budget_frames - w)
    for i in range(w):
        runner.process_frame(rgb[i], depth[i], i)
    t0 = time.perf_counter()
    for i in range(w, w + k):
        runner.process_frame(rgb[i], depth[i], i)
    dt = time.perf_counter() - t0
    v = k / dt  # one host, one set of cores: the CPU arm's throughput does not grow with the number of GPUs
    out = {"impl": "reference", "metric": "frames/sec, full track+fuse+predict (processFrame)", "value": v, "unit": "frames/s", "n_gpus": world,
           "steps": k, "warmup": w, "ms_per_step": dt / k * 1000.0, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": {"workload": wl_name, "frames_per_gpu": k, "open_loop": True},
           "cpu_baseline": {"value": v, "unit": "frames/s", "cores": ncores, "kind": kind, "sample": sample + f"; {k} frames"},
           "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "stages_s": runner.timers() if hasattr(runner, "timers") else None}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="640x480", choices=["640x480", "1280x960"])
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true", help="process each frame without staging its successor on the side stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hires-roofline", action="store_true", help="skip the 1280x960 measurement of the roofline kernel")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist_mod.init_process_group(backend="nccl")
        dist = dist_mod
    run_ours(args, rank, world, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
