#!/usr/bin/env python
"""bench.py -- scenes/s of the lifting + superpoint-pool path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic ScanNet-shaped scene:
    stable sort of points by superpoint id -> run table -> fused projection/visibility/bilinear gather/
    view mean + per-run superpoint partials -> ordered combine          (6 kernels of libsd3d.so).
Outputs per step: points_2dfeats [N,C] f32, count [N] i32, sp_feats [S,C] f32.

N > 1: independent scene replicas (one scene stream per rank, no data-path collective)
-> "scaling": "weak". `--mode viewshard` runs the one-large-scene variant (views sharded across ranks,
NCCL exchange of per-point (sum,count)) -> "scaling": "strong".

`--impl reference` times the CPU restatement of the reference path (oracle/: the reference itself is not
importable here and has no lifting code) on the host cores with all threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] (== configs[0] shape): the configuration the metric is quoted on
    "cfg2": dict(n_points=100_000, n_views=40, hd=480, wd=640, stride=8, channels=256, sp_voxel=0.6, sp_target=500),
    # configs[2]: the same scene shape with fp16 DINO-X maps and ScanNet-native u16 depth (the 8-scene batch is the
    # replica mode at --gpus 8); the SpConvUNet pooling width of that prototype (C=32) is timed by tools/bench_ops.py
    "cfg3": dict(n_points=100_000, n_views=40, hd=480, wd=640, stride=8, channels=256, sp_voxel=0.6, sp_target=500,
                 fmap_dtype="float16", depth_u16=True),
    # configs[3]: large scene
    "cfg4": dict(n_points=1_000_000, n_views=300, hd=480, wd=640, stride=8, channels=256, sp_voxel=0.2,
                 sp_target=5000),
    # a scaled-down large scene for functional multi-GPU runs of --mode viewshard
    "cfg4s": dict(n_points=250_000, n_views=120, hd=480, wd=640, stride=8, channels=256, sp_voxel=0.5,
                  sp_target=1500),
    # tiny, for CPU-side plumbing checks of this script
    "tiny": dict(n_points=4000, n_views=6, hd=120, wd=160, stride=8, channels=64, sp_voxel=0.6, sp_target=40),
}


def build_scene(wl: dict, seed: int, fmap_device=None):
    """make_scene for a WORKLOADS entry (handles the dtype keys)."""
    import torch
    from segdino3d_b200.synth import make_scene
    kw = {k: v for k, v in wl.items() if k not in ("fmap_dtype", "depth_u16")}
    if "fmap_dtype" in wl:
        kw["fmap_dtype"] = getattr(torch, wl["fmap_dtype"])
    sc = make_scene(seed=seed, fmap_device=fmap_device, **kw)
    if wl.get("depth_u16"):
        sc.depth = sc.depth_u16()
    return sc


def elem_sizes(wl: dict):
    return (2 if wl.get("fmap_dtype") in ("float16", "bfloat16") else 4), (2 if wl.get("depth_u16") else 4)


def algorithmic_bytes(n, v, hd, wd, hf, wf, c, s, sf=4, sd_=4):
    """compulsory traffic -- every distinct input byte once, every output byte once."""
    b_lift = v * (hf * wf * c * sf + hd * wd * sd_) + n * 12 + v * 64 + n * c * 4 + n * 4
    b_path = b_lift + n * 8 + s * c * 4
    return b_lift, b_path


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region (pynvml in a thread)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        srt = sorted(self.samples)
        return {"sm_mhz": srt[len(srt) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(srt)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY places bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------------------
def cpu_reference_step(sc, use_c: bool = True):
    import torch
    from oracle import c_ref
    from oracle import lift_oracle as lo
    from oracle import scatter_oracle as so
    if use_c:
        acc, cnt, _, _ = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
        feat = c_ref.finalize_ref(acc, cnt)
        sp = c_ref.scatter_mean_ref(feat, sc.sp_ids, sc.n_superpoints)
    else:
        acc, cnt, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
        feat = lo.lift_finalize_oracle(acc, cnt)
        sp = so.scatter_mean_oracle(feat, sc.sp_ids, dim=0)
    return feat, cnt, sp


def cpu_baseline(sc, budget_s: float = 12.0, max_scenes: int = 200):
    """Times the C/OpenMP port of the oracle (all host threads) on whole scenes of the same workload."""
    import torch
    from oracle import c_ref
    c_ref.use_all_host_threads()
    cpu_reference_step(sc)  # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    n = 0
    while n < max_scenes and (time.perf_counter() - t0 < budget_s or n < 2):
        cpu_reference_step(sc)
        n += 1
    dt = time.perf_counter() - t0
    torch.set_num_threads(os.cpu_count() or 1)
    t1 = time.perf_counter()
    cpu_reference_step(sc, use_c=False)
    torch_dt = time.perf_counter() - t1
    return {"value": n / dt, "unit": "scenes/s", "cores": c_ref.threads(), "kind": "port",
            "sample": f"{n} full scenes of the same workload in {dt:.1f} s (oracle/lift_ref.c, OpenMP, all host threads)",
            "torch_oracle_scenes_per_s": 1.0 / torch_dt, "host_cpus": os.cpu_count()}


def _subsample_scene(sc, n_keep):
    """First n_keep points of the scene (same views, maps, superpoint id space): a bounded sample."""
    import dataclasses
    return dataclasses.replace(sc, xyz=sc.xyz[:n_keep].contiguous(), sp_ids=sc.sp_ids[:n_keep].contiguous())


def run_reference(args):
    """The reference's CPU path for this metric = the oracle port (oracle/lift_ref.c via ctypes, OpenMP over
    all host threads). Each step is one scene, or -- when K full scenes would not finish in ~2 minutes --
    a bounded prefix of its points, with scenes/s scaled by the fraction (cost is linear in points)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_ref
    from segdino3d_b200.synth import make_scene
    cores = c_ref.use_all_host_threads()  # torchrun exports OMP_NUM_THREADS=1: the arm must still use every host core
    assert cores > 1 or (os.cpu_count() or 1) == 1, f"reference arm runs on {cores} thread(s) of {os.cpu_count()} CPUs"
    wl = WORKLOADS[args.workload]
    sc = build_scene(wl, 1235)
    cpu_reference_step(sc)
    t0 = time.perf_counter()
    cpu_reference_step(sc)
    t_full = time.perf_counter() - t0
    budget = 120.0
    frac = min(1.0, budget / max(t_full * (args.steps + args.warmup), 1e-9))
    n_keep = wl["n_points"] if frac >= 1.0 else max(1000, int(wl["n_points"] * frac))
    frac = n_keep / wl["n_points"]
    sample = sc if n_keep == wl["n_points"] else _subsample_scene(sc, n_keep)
    for _ in range(args.warmup):
        cpu_reference_step(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sample)
    dt = time.perf_counter() - t0
    val = args.steps * frac / dt
    desc = (f"{args.steps} steps x {n_keep} of {wl['n_points']} points x {wl['n_views']} views "
            f"(oracle/lift_ref.c, OpenMP, all host threads); scenes/s scaled by the point fraction {frac:.3f}; "
            "the reference repo has no lifting code and is not importable (SURVEY F1/F6)")
    line = {
        "impl": "reference", "metric": "scenes/s lifting+SP-pool", "value": val, "unit": "scenes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **{k: wl[k] for k in ("n_points", "n_views", "channels", "stride")},
                   "n_superpoints": sc.n_superpoints},
        "points_per_s": val * wl["n_points"],
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": c_ref.threads(), "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import segdino3d_b200 as sd
    from segdino3d_b200 import _lib
    from segdino3d_b200.synth import make_scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL writes its version banner to stdout when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()  # fail loudly before any timing if libsd3d.so is missing

    if args.mode == "viewshard":
        from segdino3d_b200 import dist as sdist
        return sdist.bench_viewshard(args, rank, world, dev)

    wl = WORKLOADS[args.workload]
    n_rot = args.rotate
    scenes = []
    for i in range(n_rot):
        sc = build_scene(wl, 1235 + i + 1000 * rank, fmap_device=dev)
        scenes.append(sc.to(dev))
    n, v = wl["n_points"], wl["n_views"]
    hf, wf, c = wl["hd"] // wl["stride"], wl["wd"] // wl["stride"], wl["channels"]
    s_max = max(sc.n_superpoints for sc in scenes)
    sf, sdep = elem_sizes(wl)
    scene_mb = int((v * (hf * wf * c * sf + wl["hd"] * wl["wd"] * sdep) + n * 20) / 1e6)
    b_lift, b_path = algorithmic_bytes(n, v, wl["hd"], wl["wd"], hf, wf, c, scenes[0].n_superpoints, sf, sdep)
    # the dominant kernel (gather) alone: maps once, per-point view masks, xyz, cameras in; features + count out
    # (the visible-pair term is filled in after the warm-up, when the count of visible (point, view) pairs is known)
    b_gather_fixed = v * hf * wf * c * sf + n * 4 + n * 4 + n * c * 4 + n * 4  # maps, order, nvis in; features, count out

    bufs = {}  # (scene, stream slot) -> persistent outputs + workspace: the timed step allocates nothing

    def step(sc, events=None, slot=0):
        if args.no_refine:
            plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, run=args.run)
            return sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True,
                           variant=args.variant, events=events)
        b = None
        if events is None and not args.no_overlap:
            key = (id(sc), slot)
            b = bufs.get(key)
            if b is None:
                b = bufs[key] = sd.LiftPoolBuffers(n, v, c, sc.n_superpoints, args.run, dev)
        return sd.lift_and_pool(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.sp_ids, sc.n_superpoints, stride=sc.stride,
                                run=args.run, variant=args.variant, overlap=not args.no_overlap, events=events, buffers=b)

    # scenes are independent: `--streams K` keeps K scenes in flight on K CUDA streams, so the latency-bound
    # plan / projection kernels of one scene fill the gaps of another scene's gather (throughput mode)
    streams = [torch.cuda.Stream() for _ in range(max(args.streams, 1))]
    main_stream = torch.cuda.current_stream()

    graphs = {}  # --graph: one captured CUDA graph per (scene, stream) slot, replayed instead of re-launching

    def run_steps(k, events_list=None):
        for st in streams:
            st.wait_stream(main_stream)
        for i in range(k):
            with torch.cuda.stream(streams[i % len(streams)]):
                g = graphs.get((i % n_rot, i % len(streams)))
                if g is not None and events_list is not None:
                    events_list[i][0].record()
                    g.replay()
                    events_list[i][1].record()  # brackets the whole step in graph mode, not the gather alone
                elif g is not None:
                    g.replay()
                else:
                    step(scenes[i % n_rot], None if events_list is None else events_list[i], slot=i % len(streams))
        for st in streams:
            main_stream.wait_stream(st)

    # visible (point, view) pairs per scene: one 16-byte sample record each (written by the projection kernel, read
    # by the gather), and 4 tap rows of C channels each through L1. Counted once, before the warm-up.
    def _count(r):
        return r["count"] if isinstance(r, dict) else r[1]
    pairs = sum(int(_count(step(sc)).sum()) for sc in scenes) / len(scenes)
    if not args.no_refine and not args.no_overlap:  # every (scene, stream slot) pair the loop will use: allocate now
        import math
        for i in range(n_rot * len(streams) // math.gcd(n_rot, len(streams))):
            step(scenes[i % n_rot], slot=i % len(streams))
    torch.cuda.synchronize()
    b_gather = int(b_gather_fixed)  # compulsory bytes only: the K1 -> K2 sample records are an implementation intermediate
    l1_bytes = pairs * 4 * c * sf
    run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()
    if args.graph:
        import math
        for slot in range(n_rot * len(streams) // math.gcd(n_rot, len(streams))):
            key = (slot % n_rot, slot % len(streams))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[key[1]]):
                step(scenes[key[0]], slot=key[1])
            graphs[key] = g
        torch.cuda.synchronize()
        run_steps(max(args.warmup, 3))
        torch.cuda.synchronize()

    lift_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    t_host = time.perf_counter()
    run_steps(args.steps)  # one C call per step, persistent buffers
    host_us = (time.perf_counter() - t_host) / args.steps * 1e6  # launch-side cost per step (no sync inside)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    # the dominant kernel's duration under the SAME load (scenes in flight on the same streams), bracketed by CUDA
    # events on its launch stream: a second pass right after the timed region through the step-by-step entries (same
    # kernels; the bracketing events need the gather as a call of its own)
    run_steps(args.steps, lift_events)
    torch.cuda.synchronize()
    lift_ms = sum(a.elapsed_time(b) for a, b in lift_events) / args.steps
    if world > 1:
        t = torch.tensor([total_ms, lift_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, lift_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = world * 1e3 / ms_per_step

    # the gather kernel timed ALONE (one scene at a time, nothing else on the device): with several scenes in flight
    # the in-region duration above also contains the time the kernel shares the SMs with other scenes' kernels
    alone_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for i, evs in enumerate(alone_events):
        step(scenes[i % n_rot], evs)
        torch.cuda.synchronize()
    lift_alone_ms = sum(a.elapsed_time(b) for a, b in alone_events) / len(alone_events)

    # ---- end to end through the public API with HOST buffers (H2D of every input, D2H of every output) ----
    e2e = None
    if not args.no_e2e:
        from segdino3d_b200.pipeline import ScenePipeline
        host = []
        for sc in scenes[: min(2, n_rot)]:
            h = {k: getattr(sc, k).cpu().pin_memory() for k in ("xyz", "K", "w2c", "depth", "fmap", "sp_ids")}
            h["n_superpoints"], h["stride"] = sc.n_superpoints, sc.stride
            host.append(h)
        pipe = ScenePipeline(dev, depth=3, run=args.run, variant=args.variant)
        k_e2e = max(6, min(args.steps, 40))

        def feed(k):
            for i in range(k):
                yield host[i % len(host)]

        checksum = 0.0
        for out in pipe.run(feed(4)):  # warm-up: allocates the slot buffers, pins the host outputs
            checksum += float(out[2][0, 0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for feat_h, cnt_h, sp_h in pipe.run(feed(k_e2e)):
            checksum += float(sp_h[0, 0])  # the host really reads every step's result
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {"value": world * k_e2e / dt, "unit": "scenes/s", "h2d_bytes_per_step": int(pipe.h2d_bytes),
               "d2h_bytes_per_step": int(pipe.d2h_bytes), "steps": k_e2e,
               "note": "segdino3d_b200.pipeline.ScenePipeline: pinned host scene -> H2D (xyz,K,w2c,depth,fmap,sp_ids) "
                       "-> plan+lift -> D2H (points_2dfeats,count,sp_feats); copies and kernels overlap on 3 streams",
               "gb_per_s_h2d": pipe.h2d_bytes * world * k_e2e / dt / 1e9}

    # ---- the north-star multi-GPU split, in front of the driver: ONE large scene (cfg4), views sharded over all ranks ----
    viewshard = None
    n_sp0 = scenes[0].n_superpoints
    if not args.no_viewshard:
        from segdino3d_b200 import dist as sdist
        for b in list(bufs):
            del bufs[b]
        del scenes[:]
        torch.cuda.empty_cache()
        viewshard = sdist.viewshard_report(WORKLOADS[args.viewshard_workload], args.viewshard_workload, "p2p",
                                           args.viewshard_steps, 3, rank, world, dev, args.variant)

    mask_gemm = None
    if rank == 0 and not args.no_mask:
        if args.no_viewshard:
            for b in list(bufs):
                del bufs[b]
            del scenes[:]
            torch.cuda.empty_cache()
        mask_gemm = mask_gemm_report(dev)

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        traffic, traffic_src = ncu_traffic(args)
        ach = b_gather / (lift_ms * 1e-3) / 1e9
        line = {
            "metric": "scenes/s lifting+SP-pool", "value": value, "unit": "scenes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_points": n, "n_views": v, "depth": [wl["hd"], wl["wd"]],
                       "fmap_dtype": wl.get("fmap_dtype", "float32"), "depth_dtype": "uint16 mm" if wl.get("depth_u16") else "float32 m",
                       "fmap": [hf, wf, c], "stride": wl["stride"], "n_superpoints": n_sp0,
                       "parallelism": "scene replicas (no collective)" if world > 1 else "single GPU",
                       "l2": f"inputs rotate over {n_rot} distinct scenes ({n_rot * scene_mb} MB > 126 MB L2), no flush",
                       "run": args.run, "variant": args.variant, "streams": len(streams),
                       "blend": "fma (<= 1e-5 of the oracle)" if args.variant & 1 else "unfused, bit-exact (Appendix A order)",
                       "gather": "staged in shared memory (cp.async.bulk)" if args.variant & 32768 else "direct (LDG.128 tap rows)",
                       "projection_overlaps_plan": not args.no_overlap, "cuda_graph": bool(args.graph)},
            "points_per_s": value * n, "host_us_per_step": host_us,
            "roofline": {"bound": "hbm", "kernel": "gather_kernel (bilinear gather + view mean + run partials)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": b_gather, "kernel_ms": lift_ms, "peak_source": peak_src,
                         "kernel_ms_alone": lift_alone_ms, "frac_alone": b_gather / (lift_alone_ms * 1e-3) / 1e9 / peak,
                         "l1_path": {"note": "what actually bounds the kernel: every visible (point, view) pair pulls "
                                             "4 tap rows of C fp32 channels through L1 (LDG.128 = 4 wavefronts at "
                                             "~2 cycles each -> ~64 B/clk/SM, B300_MICROARCH.md load model)",
                                     "bytes_per_launch": int(l1_bytes),
                                     "achieved_gbs": l1_bytes / (lift_alone_ms * 1e-3) / 1e9,
                                     "peak_gbs": 148 * 64 * clocks.get("sm_mhz", 1965) * 1e6 / 1e9
                                     if isinstance(clocks, dict) and clocks.get("sm_mhz") else None},
                         "path_algorithmic_bytes": b_path,
                         "path_frac": b_path / (ms_per_step * 1e-3) / 1e9 / peak},
            "clocks": clocks,
            "gpu_launches": (5 if args.no_refine else 6) * args.steps,  # radix pass, run table, (refine,) projection, gather, combine
        }
        if e2e is not None:
            line["e2e"] = e2e
        if viewshard is not None:
            line["viewshard"] = viewshard
        if mask_gemm is not None:
            line["mask_gemm"] = mask_gemm
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(build_scene(wl, 1235))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()  # rank 0 measures the mask GEMM and prints while the others wait here
        dist.destroy_process_group()


def mask_gemm_report(dev):
    """The tensor-core leg of the path (a-5, instance_seg_3d_decoder.py:567-573), timed with CUDA events on rank 0:
    the eval-scale contraction (queries = superpoints, 5000 x 5000 x 256) through the TMA-fed tcgen05 kernel with
    bf16 and bf16x3 (fp32-tolerance) operands, and the ScanNet200 decoder shape of BASELINE configs[1] (200 queries x
    500 superpoints, 8 scenes in one batched launch with the attention mask). Fractions are of the measured dense
    bf16 peak (MEASURED_PEAKS.json: cuBLAS burst); algorithmic flops = 2 n S d (bf16x3 issues 3x that)."""
    import torch
    import segdino3d_b200 as sd
    from segdino3d_b200.synth import make_decoder_operands
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json, burst)"
    except Exception:
        peak, peak_src = 2250.0, "nominal dense bf16 (B200_PROFILING.md)"

    def timeit(fn, iters):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e-3

    n = s = 5000
    d = 256
    q, mf = make_decoder_operands(n, s, d)
    q, mf = q.to(dev), mf.to(dev)
    _, q16 = sd.layernorm_cast(q, normalize=False, want_f32=False)
    _, mf16 = sd.layernorm_cast(mf, normalize=False, want_f32=False)
    q2, mf2 = sd.split_bf16(q), sd.split_bf16(mf)
    flop = 2.0 * n * s * d
    t16 = timeit(lambda: sd.mask_logits_bf16(q16, mf16), 50)
    t16m = timeit(lambda: sd.mask_logits_bf16(q16, mf16, threshold=0.5), 50)
    tx3 = timeit(lambda: sd.mask_logits_bf16(q2, mf2, split=True), 50)
    tref = timeit(lambda: torch.einsum("nd,md->nm", q, mf), 20)
    ops_ = [make_decoder_operands(200, 500, d, seed=i) for i in range(8)]
    qs, mfs = [o[0].to(dev) for o in ops_], [o[1].to(dev) for o in ops_]
    tb = timeit(lambda: sd.mask_logits_batched(qs, mfs, precision="bf16", threshold=0.5), 100)

    def torch_head():
        for a_, b_ in zip(qs, mfs):
            pm = torch.einsum("nd,md->nm", a_, b_)
            am = pm.sigmoid() < 0.5
            am[torch.where(am.sum(-1) == am.shape[-1])] = False
    tt = timeit(torch_head, 50)
    return {"bound": "tensor", "kernel": "mask_logits_tma_kernel (TMA loads, tcgen05.mma M128 N256 K16, TMEM, TMA stores)",
            "shape": [n, s, d], "dtype": "bf16 operands, fp32 accumulate / output", "us": t16 * 1e6,
            "achieved": flop / t16 / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flop / t16 / 1e12 / peak,
            "peak_source": peak_src, "us_with_attn_mask": t16m * 1e6,
            "bf16x3_fp32_tolerance": {"us": tx3 * 1e6, "algorithmic_tflops": flop / tx3 / 1e12,
                                      "issued_tflops": 3 * flop / tx3 / 1e12},
            "torch_einsum_fp32_us": tref * 1e6,
            "decoder_shape_8_scenes": {"shape": [200, 500, d], "batched_bf16_with_attn_mask_us": tb * 1e6,
                                       "torch_reference_sequence_us": tt * 1e6},
            "note": "output 100 MB fp32 per call (a fill_ of it alone: ~16.6 us); timings include the host side of each call"}


def ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one gather_kernel launch, from the committed `ncu --set full`
    capture of this same command (profiles/rNN_ncu_gather_summary.txt). Only valid for the configuration it was taken
    on (cfg2, default kernel), otherwise null."""
    import glob
    import re
    if args.workload != "cfg2" or (args.variant & ~1) != 0 or args.run != 32:
        return None, "no ncu capture for this configuration"
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                                          "r*_ncu_gather_summary.txt")))
    if not files:
        return None, "profiles/ has no gather capture"
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for ln in open(files[-1]):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", ln)
        if m:
            tot += float(m.group(2)) * unit[m.group(3)]
    return (int(tot), os.path.relpath(files[-1], os.path.dirname(os.path.abspath(__file__)))) if tot else (None, "parse")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="replicas", choices=["replicas", "viewshard"])
    ap.add_argument("--exchange", default="reduce_scatter", choices=["allreduce", "reduce_scatter", "p2p"])
    ap.add_argument("--rotate", type=int, default=4, help="distinct scenes cycled through (defeats L2 residency)")
    ap.add_argument("--run", type=int, default=32, help="points per warp run")
    ap.add_argument("--variant", type=int, default=1,
                    help="sd3d_lift variant bits (include/sd3d.h). Default 1 = bilinear blend contracted into FFMA: within "
                         "the north star's 1e-5 of the oracle (tests: test_lift_fma_variant_*), integers still bit-exact; "
                         "0 = the bit-exact unfused blend (Appendix A order); 32768 = shared-memory staged gather")
    ap.add_argument("--no-refine", action="store_true", help="skip the Morton refinement of the processing order")
    ap.add_argument("--streams", type=int, default=2, help="scenes kept in flight on separate CUDA streams")
    ap.add_argument("--graph", action="store_true",
                    help="capture each scene's step in a CUDA graph and replay it (launch-side cost ~0; in this mode "
                         "roofline.kernel_ms brackets the whole step, use kernel_ms_alone for the gather)")
    ap.add_argument("--no-overlap", action="store_true", help="projection kernel on the main stream (no side stream)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-viewshard", action="store_true", help="skip the view-sharded large-scene measurement")
    ap.add_argument("--viewshard-workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--viewshard-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mask", action="store_true", help="skip the mask-GEMM (tensor roofline) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
