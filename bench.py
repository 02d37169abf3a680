#!/usr/bin/env python
"""bench.py -- BEV raster throughput (Mpoints/s) on N B200s, one JSON line.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (numpy port)

A *step* is one pass of the hot path over one synthetic cloud.
  N = 1   BASELINE.json configs[1]: 100 M-point road segment -> 11520 x 1152 x 3 u8 at 0.05 m.
  N > 1   configs[2] geometry, weak scaling: every rank owns one 1440-row strip of an
          (1440 N) x 11520 scene with 1.25e8 points (N = 8 is exactly the 1 B-point scene);
          the step includes the NCCL halo exchange + merge; the finished mosaic stays sharded, one strip per GPU, and
          every rank delivers its own strip in the e2e leg (--gather root / all: also gather it on rank 0 / on every
          rank inside the step, and D2H it there); exchange, merge and gather of scene k run on side streams under the
          rasterisation of scene k+1 and all of them finish inside the timed region.
``value``  device-resident throughput (inputs in HBM when the clock starts), CUDA events,
           max over ranks.   ``e2e``: the same metric through the host-buffer API
           (pinned H2D of the points + D2H of the finished raster inside the timed region).
``roofline`` is for the dominant kernel, timed live with events between the pipeline stages of the same
           timed steps: ``sweep_kernel`` (algo sweep: the whole path in one kernel) or ``bin_points`` (algo binned).
``parity_checked`` / ``mismatches``: after the timed region the timed output is compared on the device with the
           global-atomic cross-check algorithm on the same points, and a 2 M-point sub-scene with the numpy oracle.
``configs``  (N = 1) the other BASELINE.json configs and the shuffled ordering, each device-resident, best of 3,
           with its own path fraction and parity flag.
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

import numpy as np

# strip boundaries on multiples of 32 rows: the 1440-row strips of the weak-scaling scene come out equal,
# so the mosaic gather is one all_gather_into_tensor straight from the strips (no padding, no concatenation)
STRIP_ALIGN = 32
METRIC = "bev_raster_mpoints_per_s"
UNIT = "Mpoints/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload(n_gpus: int, rank: int, points_override: int = 0):
    """-> (global spec, this rank's point count, description).  See module docstring."""
    from dataclasses import replace
    from lanemapping_b200.spec import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z
    from lanemapping_b200.synth import default_min_ele
    ch = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY)
    if n_gpus == 1:
        sp = BevSpec(11520, 1152, channels=ch)
        n = points_override or 100_000_000
        name = "configs[1]: 100M-point MLS road segment, 11520x1152 cells @0.05 m, 3ch u8 (max_i, mean_z, density)"
    else:
        sp = BevSpec(1440 * n_gpus, 11520, channels=ch)
        n = points_override or 125_000_000
        name = (f"configs[2] geometry, weak scaling: {n_gpus} strips of 1440x11520 cells @0.05 m, "
                f"{n / 1e6:.0f}M points per GPU, NCCL halo merge in the step")
    return replace(sp, local_min_ele=default_min_ele(BevSpec(1152, 1152))), n, name


# ---------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (a port: the reference has no rasteriser of its own) over a pool
# ---------------------------------------------------------------------------------------------
def cpu_rasterize_timed(cloud, spec, processes, reps):
    """Best-of-reps wall time of oracle.rasterize_pool on host cores.  Row strips + the point
    index range each strip needs (scan-ordered clouds: contiguous ranges with a 2 m margin)."""
    from oracle import bev_oracle as O
    n = len(cloud)
    P = max(1, processes)
    roads = max(1, int(round(spec.width * spec.img_reso[1] / 57.6)))     # synth.make_cloud's default
    ranges = O.scan_point_ranges(n, spec, P, roads=roads)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        O.rasterize_pool(cloud, spec, P, point_ranges=ranges)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_sample(spec, n_sample, n_total, seed_rank=0):
    """A bounded sample of the same workload at the workload's own density.  One-road scenes (config 2):
    the first rows.  Multi-road scenes (config 3 geometry): a rank's 1440 rows of the first roads, so that
    the pool's row strips stay long compared with the +-1 m scan jitter, as they are at full size.
    -> (window spec, cloud, description)"""
    from lanemapping_b200.synth import make_cloud
    density = n_total / spec.cells
    if spec.width <= 2 * 1152:
        rows = max(128, int(round(spec.height * n_sample / n_total / 128)) * 128)
        rows = min(rows, spec.height)
        sub = spec.window(0, rows)
    else:
        rows = min(spec.height, 1440)
        k = int(round(n_sample / (density * rows * 1152)))
        k = min(max(k, 1), spec.width // 1152)
        sub = spec.window(0, rows, 0, k * 1152)
        n_sample = int(density * rows * k * 1152)
    cloud = make_cloud(n_sample, sub, order="scan", seed=2021 + seed_rank)
    what = f"{n_sample} points = rows [0, {sub.height}) x columns [0, {sub.width}) of the workload at full density"
    return sub, cloud, what


def numpy_single_thread(cloud, spec, n=5_000_000):
    """Mpoints/s of the plain single-thread numpy oracle on the first n points."""
    from oracle import bev_oracle as O
    sub = cloud[:n]
    O.rasterize(sub[:100_000], spec)
    t0 = time.perf_counter()
    O.rasterize(sub, spec)
    return len(sub) / (time.perf_counter() - t0) / 1e6


def png_encode_ms(spec):
    """ms to PNG-encode one 1152^2 crop the way the offline converter does (cv2, zlib level 1 + RLE strategy,
    lanemapping_b200/convert_data.py::png_params), on one host thread."""
    import cv2
    from lanemapping_b200.convert_data import png_params
    from lanemapping_b200.synth import make_cloud
    from oracle import bev_oracle as O
    from dataclasses import replace
    sp = replace(spec, height=1152, width=1152, row0=0, col0=0)
    img = O.rasterize(make_cloud(2_000_000, sp, order="scan"), sp)["image"]
    cv2.imencode(".png", img, png_params())
    t0 = time.perf_counter()
    for _ in range(3):
        cv2.imencode(".png", img, png_params())
    return (time.perf_counter() - t0) / 3 * 1e3


def run_reference(args):
    """CPU arm: the numpy oracle (a port -- the reference has no rasteriser) over multiprocessing.Pool(all cores),
    the reference's own offline idiom (data/convert_data.py:429-436).  Honours --steps / --warmup; every step is
    one pooled rasterisation of the sample (default: the FULL config, sample_fraction 1.0)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec, n_full, name = workload(args.gpus, 0, args.points)
    n_total = n_full * args.gpus
    # default: the full config at N = 1 (10^8 points, sample_fraction 1.0); the multi-GPU scenes (up to 10^9 points) are
    # sampled at 10^8 points so that the arm still ends within a few minutes
    n_sample = min(n_total, args.cpu_points if args.cpu_points > 0 else 100_000_000)
    sub, cloud, what = cpu_sample(spec, n_sample, n_total)
    n_sample = len(cloud)
    P = os.cpu_count() or 1
    W, K = max(0, args.warmup), max(1, args.steps)
    for _ in range(W):
        cpu_rasterize_timed(cloud, sub, P, 1)
    t0 = time.perf_counter()
    for _ in range(K):
        cpu_rasterize_timed(cloud, sub, P, 1)
    dt = (time.perf_counter() - t0) / K
    v = n_sample / dt / 1e6
    sample = f"{what}; numpy oracle (floor keys, bincount, maximum.at) over multiprocessing.Pool({P}) row strips"
    cb = {"value": round(v, 3), "unit": UNIT, "cores": P, "kind": "port", "sample": sample,
          "sample_fraction": round(n_sample / n_total, 4)}
    try:
        cb["numpy_1thread_mpoints_per_s"] = round(numpy_single_thread(cloud, sub), 3)
        cb["png_encode_ms_per_1152_crop"] = round(png_encode_ms(spec), 2)
    except Exception as ex:      # the extras must never cost the line
        cb["extras_error"] = repr(ex)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": name, "sample": sample, "sample_fraction": cb["sample_fraction"]},
        "cpu_baseline": cb,
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML in a thread; the recipe's nvidia-smi line as a fallback)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "app_clocks", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x100: "display"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
        except Exception:
            self._nvml = None
            return self
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self._h).gpu
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((mhz, util))
                for b, name in self.REASONS.items():
                    if bits & b and name not in ("gpu_idle", "app_clocks", "sync_boost", "display"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        mhz = [s[0] for s in self.samples]
        out = {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(mhz)}
        return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def _timed_best(fn, stream, reps=3, warm=2):
    import torch
    for _ in range(warm):
        fn()
    stream.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def _same(a, b):
    """number of differing elements of two device tensors (0 = bit-identical)"""
    import torch
    return int((a != b).sum().item()) if a.shape == b.shape else -1


def other_configs(dev, peak, algo, pts_cfg2, spec2, stream):
    """BASELINE.json configs[0], [3], [4] and the shuffled ordering of configs[1]: device-resident, best of 3, each
    checked on the device against the global-atomic cross-check algorithm (bit-identical = parity true)."""
    import torch
    from lanemapping_b200.bev import BatchRasterizer, BevRasterizer, crop_tiles
    from lanemapping_b200.synth import config_spec, make_cloud
    out = []

    def entry(name, n, spec, ms, mism, extra=None):
        b_alg = float(spec.algorithmic_bytes(n)) if extra is None or "b_alg" not in extra else extra.pop("b_alg")
        e = {"config": name, "points": n, "ms": round(ms, 4), "mpoints_per_s": round(n / ms / 1e3, 1),
             "path_algorithmic_bytes": b_alg, "path_frac": round(b_alg / (ms * 1e-3) / 1e9 / peak, 4),
             "parity": mism == 0, "mismatches": mism}
        if extra:
            e.update(extra)
        out.append(e)

    # ---- configs[1] through the experimental single-pass sweep (no record pool in HBM), same cloud
    if algo != "sweep":
        rs = BevRasterizer(spec2, len(pts_cfg2), device=dev, algo="sweep", outputs=("image",))
        os_ = rs.alloc_outputs()
        ms = _timed_best(lambda: rs(pts_cfg2, out=os_), stream)
        dd = BevRasterizer(spec2, len(pts_cfg2), device=dev, algo="direct", outputs=("image",))
        entry("configs[1] scan order through LM_ALGO_SWEEP (experimental single-pass kernel)", len(pts_cfg2), spec2, ms,
              _same(os_["image"], dd(pts_cfg2)["image"]), {"sweep": rs.sweep_state()})
        del rs, os_, dd
        torch.cuda.empty_cache()

    # ---- configs[1] shuffled (worst-case ordering of the headline cloud; permuted on the device)
    g = torch.Generator(device=dev)
    g.manual_seed(2021)
    shuf = pts_cfg2[torch.randperm(pts_cfg2.shape[0], device=dev, generator=g)].contiguous()
    r = BevRasterizer(spec2, len(shuf), device=dev, algo=algo, outputs=("image",))
    o = r.alloc_outputs()
    ms = _timed_best(lambda: r(shuf, out=o), stream)
    d = BevRasterizer(spec2, len(shuf), device=dev, algo="direct", outputs=("image",))
    entry("configs[1] shuffled (same 100M points in random order)", len(shuf), spec2, ms, _same(o["image"], d(shuf)["image"]),
          {"sweep": r.sweep_state()} if algo == "sweep" else None)
    del shuf, r, d, o
    torch.cuda.empty_cache()

    # ---- configs[4]: batch-8 on-the-fly proj (8 crops x 10M points -> f32 [8,3,1152,1152]); crop 0 doubles as configs[0]
    spec5, n5 = config_spec(5)
    clouds = [torch.from_numpy(make_cloud(n5, spec5, seed=100 + b, order="scan")).to(dev) for b in range(8)]
    br = BatchRasterizer(spec5, 8, 8 * n5, device=dev, outputs=("proj",))
    ob = br.alloc_outputs()
    ms = _timed_best(lambda: br(clouds, None, out=ob), stream)
    mism = 0
    d5 = BevRasterizer(spec5, n5, device=dev, algo="direct", outputs=("proj",))
    for b in (0, 7):
        mism += _same(ob["proj"][b], d5(clouds[b])["proj"])
    entry("configs[4]: batch-8 on-the-fly proj, 8 x 10M points -> f32 [8,3,1152,1152]", 8 * n5, spec5, ms, mism,
          {"b_alg": float(16 * 8 * n5 + 8 * 3 * 1152 * 1152 * 4)})
    # ---- configs[0]: one 10M-point tile, intensity channel
    spec1, n1 = config_spec(1)
    r1 = BevRasterizer(spec1, n1, device=dev, algo=algo, outputs=("image",), graph=True)   # launch-bound: replay a CUDA graph
    o1 = r1.alloc_outputs()
    ms = _timed_best(lambda: r1(clouds[0], out=o1), stream)
    d1 = BevRasterizer(spec1, n1, device=dev, algo="direct", outputs=("image",))
    entry("configs[0]: one 10M-point tile, 1152x1152, intensity channel", n1, spec1, ms, _same(o1["image"], d1(clouds[0])["image"]),
          {"sweep": r1.sweep_state()} if algo == "sweep" else None)
    del clouds, br, ob, d5, r1, o1, d1
    torch.cuda.empty_cache()

    # ---- configs[3]: 0.02 m, 4 x u8 + u16 count, plus the 1152^2 crop tiling of the image (75 crops)
    spec4, n4 = config_spec(4)
    p4 = torch.from_numpy(make_cloud(n4, spec4, order="scan")).to(dev)
    r4 = BevRasterizer(spec4, n4, device=dev, algo=algo, outputs=("image", "count16"))
    o4 = r4.alloc_outputs()
    crops = {}

    def step4():
        r4(p4, out=o4)
        crops["c"] = crop_tiles(o4["image"], 1152)
    ms = _timed_best(step4, stream)
    ms_crop = _timed_best(lambda: crop_tiles(o4["image"], 1152), stream)
    d4 = BevRasterizer(spec4, n4, device=dev, algo="direct", outputs=("image", "count16"))
    od = d4(p4)
    mism = _same(o4["image"], od["image"]) + _same(o4["count16"], od["count16"])
    mism += _same(crops["c"][1], o4["image"][0:1152, 1152:2304])
    entry("configs[3]: 100M points, 0.02 m, 28800x3456, 4 x u8 + u16 count, + 1152^2 crop tiling (75 crops)", n4, spec4, ms, mism,
          {"crop_tiles_ms": round(ms_crop, 4)})
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    spec, n_pts, name = workload(args.gpus, rank, args.points)
    algo = args.algo if args.gpus == 1 else ("binned" if args.algo == "sweep" else args.algo)

    from lanemapping_b200.synth import make_cloud
    from lanemapping_b200.strips import strip_bounds

    # ---- this rank's synthetic cloud (host, numpy)
    if args.gpus == 1:
        cloud = make_cloud(n_pts, spec, order=args.order)
    else:
        r0, r1 = strip_bounds(spec.height, world, STRIP_ALIGN)[rank]
        # the strip's own points; the +-1 m scan jitter strays up to 20 rows into the neighbours' strips
        cloud = make_cloud(n_pts, spec.window(r0, r1), order=args.order, seed=2021 + rank)

    # ---- CPU baseline first (forks workers: do it before CUDA is initialised), rank 0 at N=1 only
    cpu_baseline = None
    oracle_sub = None
    if args.gpus == 1 and rank == 0:
        # a 2 M-point sub-scene for the parity check against the numpy oracle: the first points of the cloud
        # on the row window they fall in (scan order: the first ~250 rows; any order: the whole raster)
        from oracle import bev_oracle as O
        m = min(n_pts, 2_000_000)
        rows = spec.height if args.order != "scan" else min(spec.height, 128 * (2 + int(m / max(n_pts / spec.height, 1) / 128)))
        sub_spec = spec.window(0, rows)
        oracle_sub = (m, sub_spec, O.rasterize(cloud[:m], sub_spec)["image"])
    if args.gpus == 1 and rank == 0 and not args.no_cpu_baseline:
        P = os.cpu_count() or 1
        n_sample = min(n_pts, args.cpu_points if args.cpu_points > 0 else 20_000_000)
        sub, sample_cloud, what = cpu_sample(spec, n_sample, n_pts)
        n_sample = len(sample_cloud)
        dt = cpu_rasterize_timed(sample_cloud, sub, P, 2)
        cpu_baseline = {"value": round(n_sample / dt / 1e6, 3), "unit": UNIT, "cores": P, "kind": "port",
                        "sample": f"{what}, numpy oracle over multiprocessing.Pool({P}) row strips, best of 2",
                        "sample_fraction": round(n_sample / n_pts, 4)}
        try:
            cpu_baseline["numpy_1thread_mpoints_per_s"] = round(numpy_single_thread(sample_cloud, sub), 3)
            cpu_baseline["png_encode_ms_per_1152_crop"] = round(png_encode_ms(spec), 2)
        except Exception as ex:
            cpu_baseline["extras_error"] = repr(ex)
        try:      # for scale: the plain-C restatement of the same spec, one scalar loop on one core
            from oracle import c_oracle as CO
            m = min(len(sample_cloud), 10_000_000)
            CO.rasterize(sample_cloud[:1000], sub)
            t0 = time.perf_counter()
            CO.accumulate(sample_cloud[:m], sub)
            cpu_baseline["c_port_1core"] = round(m / (time.perf_counter() - t0) / 1e6, 3)
        except Exception:
            pass
        del sample_cloud

    import torch
    import torch.distributed as dist
    from lanemapping_b200 import _cabi
    from lanemapping_b200.bev import BevRasterizer, HostRasterizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_pts = torch.from_numpy(cloud).pin_memory()
    pts = host_pts.to(dev)
    K, Wm = args.steps, max(args.warmup, 3)
    stream = torch.cuda.current_stream(dev)
    sampler = ClockSampler(local_rank).start()

    stage_ms = None
    stage_names = None
    if args.gpus == 1:
        r = BevRasterizer(spec, n_pts, device=dev, algo=algo, outputs=("image",))
        out = r.alloc_outputs()
        step = lambda: r(pts, out=out)
        # the same step, split at the stage boundaries with events between (identical launches)
        if algo == "sweep":
            splits = [_cabi.STAGE_SWEEP, _cabi.STAGE_BIN | _cabi.STAGE_INDEX | _cabi.STAGE_REDUCE]
            stage_names = ["sweep_kernel(+memset, epilogue)", "fall-back kernels (return at once after a good sweep)"]
            launches_per_step = 6      # sweep, epilogue, 4 gated two-pass kernels (+1 memset node)
        else:
            splits = [_cabi.STAGE_BIN, _cabi.STAGE_INDEX, _cabi.STAGE_REDUCE]
            stage_names = ["bin_points(+memset)", "scan+index", "reduce_tiles"]
            launches_per_step = 4      # bin_points, scan_tiles, index_chunks, reduce_tiles (+1 memset node)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(splits) + 1)] for _ in range(K)]

        def staged(i):
            e = evs[i]
            e[0].record(stream)
            for j, sp in enumerate(splits):
                r(pts, out=out, stages=sp)
                e[j + 1].record(stream)
    else:
        from lanemapping_b200.strips import StripRasterizer
        groot = {"root": 0, "all": None, "none": "none"}[args.gather]
        sr = StripRasterizer(spec, n_pts, halo=args.halo, device=dev, align=STRIP_ALIGN, gather_root=groot, time_stages=True,
                             gather_parts=args.gather_parts)
        mosaic_holder = {}

        def step():
            # scene k's halo exchange, merge and mosaic gather (side stream) overlap scene k+1's rasterisation; every
            # one of them completes inside the timed region (sr.flush() before the closing event)
            mosaic_holder["slot"] = sr.step(pts)
        staged = None
        launches_per_step = 4 + 2 * (1 if rank in (0, world - 1) else 2)     # raster + one merge_finalize per neighbour (+ NCCL kernels)

    for _ in range(Wm):
        step()
    barrier()
    if args.gpus > 1:
        sr.reset_stage_times()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for i in range(K):
        if staged is not None:
            staged(i)
        else:
            step()
    if args.gpus > 1:
        sr.flush()
    t_stop.record(stream)
    barrier()
    ms_total = t_start.elapsed_time(t_stop)
    if staged is not None:
        stage_ms = [float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(K)])) for j in range(len(splits))]
    sweep_state = None
    if args.gpus == 1:
        r.check_device_errors()
        n_valid = r.stats()["n_valid"]
        sweep_state = r.sweep_state() if algo == "sweep" else None
    else:
        sr.raster.check_device_errors()
        n_valid = sr.raster.stats()["n_valid"]
        stage_t = sr.stage_times()
        # the same rank's rasterisation alone (no exchange, no gather, no peers' traffic): what the step costs without comm
        barrier()
        ro = sr.raster.alloc_outputs()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(K):
            sr.raster(pts, out=ro)
        a1.record(stream)
        a1.synchronize()
        stage_t["raster_alone"] = a0.elapsed_time(a1) / K
        del ro
        # verify what was timed: every rank checks its finished strip of the last timed scene against the independent
        # DIRECT-algorithm path (strips.verify), rank 0 checks that the gathered mosaic holds the ranks' strips
        slot = mosaic_holder["slot"]
        my_strip = sr.strip_of(slot)
        mism = sr.verify(pts, my_strip)
        sums = torch.zeros(world, dtype=torch.int64, device=dev)
        sums[rank] = my_strip.to(torch.int64).sum()
        dist.all_reduce(sums)
        mos = sr.mosaic(slot)
        if mos is not None:
            rows = [b - a for a, b in sr.plan.bounds]
            off = 0
            for k_, rr in enumerate(rows):
                mism += int(mos[off:off + rr].to(torch.int64).sum().item() != int(sums[k_].item()))
                off += rr
            mism += _same(mos[sr.plan.strip[0]:sr.plan.strip[1]], my_strip)
        mt = torch.tensor([mism], dtype=torch.int64, device=dev)
        dist.all_reduce(mt)
        parity_n = {"parity_checked": True, "mismatches": int(mt.item()),
                    "parity_how": "every rank: finished strip of the last timed scene == LM_ALGO_DIRECT raster of its points with the halo "
                                  "planes exchanged and merged by torch integer ops; gathered mosaic == the ranks' strips (checksums, own strip bytes)"}
        st_t = torch.tensor([stage_t.get(k_, 0.0) for k_ in ("raster", "halo_exchange", "merge_finalize", "gather", "side_stream_total", "raster_alone")],
                            dtype=torch.float64, device=dev)
        dist.all_reduce(st_t, op=dist.ReduceOp.MAX)
        stage_n = dict(zip(("raster", "halo_exchange", "merge_finalize", "gather", "side_stream_total", "raster_alone"),
                           [round(float(v), 4) for v in st_t]))

    # ---- verify what was timed (N = 1): the timed output against the global-atomic algorithm on the same points,
    #      and a 2 M-point sub-scene through the same algo against the numpy oracle
    parity = None
    if args.gpus == 1:
        d = BevRasterizer(spec, n_pts, device=dev, algo="direct", outputs=("image",))
        mism = _same(out["image"], d(pts)["image"])
        del d
        m, sub_spec, want = oracle_sub
        rs = BevRasterizer(sub_spec, m, device=dev, algo=algo, outputs=("image",))
        mism_oracle = int((rs(pts[:m])["image"].cpu().numpy() != want).sum())
        parity = {"parity_checked": True, "mismatches": mism + mism_oracle,
                  "parity_how": f"timed raster == LM_ALGO_DIRECT raster of the same {n_pts} points on the device ({mism} differing bytes); "
                                f"first {m} points on rows [0,{sub_spec.height}) == numpy oracle ({mism_oracle} differing bytes)"}
        del rs
        torch.cuda.empty_cache()

    # ---- e2e through the host-buffer API: pinned H2D + kernels (+ collectives) + D2H
    if args.gpus == 1:
        hr = HostRasterizer(spec, n_pts, device=dev, algo=algo, outputs=("image",))
        e2e_step = lambda: hr(host_pts)
        d2h = spec.cells * spec.n_channels
    else:
        dev_in = torch.empty_like(pts)
        r0, r1 = sr.plan.strip
        # what the step pays for is what the e2e delivers: the gathered mosaic goes to the host on the rank that
        # holds it (rank 0 for --gather root, every rank for all); with --gather none every rank delivers its own strip
        if args.gather == "none":
            host_dst = torch.empty((r1 - r0, spec.width, spec.n_channels), dtype=torch.uint8, pin_memory=True)
        elif args.gather == "all" or rank == 0:
            host_dst = torch.empty((spec.height, spec.width, spec.n_channels), dtype=torch.uint8, pin_memory=True)
        else:
            host_dst = None

        def e2e_step():
            dev_in.copy_(host_pts, non_blocking=True)
            slot = sr.step(dev_in)
            m = sr.mosaic(slot)                                  # the scene's merge + gather are part of the step
            if args.gather == "none":
                host_dst.copy_(sr.strip_of(slot), non_blocking=True)
            elif m is not None:
                host_dst.copy_(m, non_blocking=True)
            stream.synchronize()
        d2h = host_dst.numel() if host_dst is not None else 0
        d2h_t = torch.tensor([d2h], dtype=torch.int64, device=dev)
        dist.all_reduce(d2h_t, op=dist.ReduceOp.MAX)
        d2h = int(d2h_t.item())
    Ke = max(1, min(K, args.e2e_steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    if args.gpus == 1:
        del hr
        torch.cuda.empty_cache()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
    total_pts = n_pts * world
    ms_step = ms_total / K
    value = total_pts / (ms_step * 1e-3) / 1e6
    e2e_value = total_pts / (e2e_s / Ke) / 1e6

    if rank == 0:
        peak, peak_src = load_peaks()
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": name, "order": args.order, "algo": algo,
                       **({"mosaic": {"root": "gathered on rank 0", "all": "all-gathered on every rank",
                                      "none": "left sharded (every rank keeps its strip)"}[args.gather]} if args.gpus > 1 else {}),
                       "points_per_gpu": n_pts, "valid_points_rank0": int(n_valid),
                       "l2_policy": "inputs (1.6+ GB of points per step) exceed the 126 MB L2; no flush needed; the 39.8 MB "
                                    "output image is rewritten every step and partly stays in L2 (2.4 % of the algorithmic bytes)",
                       "timing": "CUDA events on the launch stream, barrier+synchronize both sides, max over ranks"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": n_pts * 16,
                    "d2h_bytes_per_step": int(d2h), "steps": Ke},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
        }
        if parity is not None:
            line.update(parity)
        if args.gpus > 1:
            line.update(parity_n)
            b_rank = float(16 * n_pts + (spec.height // world) * spec.width * spec.n_channels)
            line["roofline"] = {
                "bound": "hbm", "kernel": "bin_points_kernel (per rank; the strip pipeline is bin + index + reduce + merge_finalize)",
                "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                "stage_ms_max_over_ranks": stage_n,
                "halo_bytes_sent_per_rank": sr.halo_bytes, "halo_planes": sr.planes, "halo_rows": args.halo,
                "mosaic_bytes": spec.cells * spec.n_channels,
                "path_algorithmic_bytes_per_rank": b_rank,
                "path_achieved": round(b_rank / (ms_step * 1e-3) / 1e9, 1),
                "path_frac": round(b_rank / (ms_step * 1e-3) / 1e9 / peak, 4),
                "achieved": round(b_rank / (stage_n["raster_alone"] * 1e-3) / 1e9, 1),
                "frac": round(b_rank / (stage_n["raster_alone"] * 1e-3) / 1e9 / peak, 4),
                "note": "achieved/frac: one rank's rasterisation alone (no exchange, no gather); path_*: the whole step incl. "
                        "halo exchange, merge and mosaic gather on the side stream"}
        if sweep_state is not None:
            line["config"]["sweep"] = sweep_state
        if stage_ms is not None:
            b_alg_path = float(spec.algorithmic_bytes(n_pts))
            # dominant kernel: the sweep does the whole path (every point read once, every cell written once);
            # bin_points of the two-pass pipeline reads every point record (16 B) once
            kernel = "sweep_kernel" if algo == "sweep" else "bin_points_kernel"
            b_alg_kernel = b_alg_path if algo == "sweep" else 16.0 * n_pts
            achieved = b_alg_kernel / (stage_ms[0] * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                try:
                    with open(tp) as f:
                        traffic = json.load(f).get(kernel.replace("_kernel", "") + "_dram_bytes_per_launch_100M")
                    if n_pts != 100_000_000 or args.order != "scan":
                        traffic = None
                except Exception:
                    traffic = None
            line["roofline"] = {
                "bound": "hbm", "kernel": kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_alg_kernel,
                "stage_ms": {nm: round(v, 4) for nm, v in zip(stage_names, stage_ms)},
                "path_algorithmic_bytes": b_alg_path,
                "path_achieved": round(b_alg_path / (ms_step * 1e-3) / 1e9, 1),
                "path_frac": round(b_alg_path / (ms_step * 1e-3) / 1e9 / peak, 4),
            }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if args.gpus == 1 and args.configs == "all" and args.points == 0 and args.order == "scan":
            try:
                line["configs"] = other_configs(dev, peak, algo, pts, spec, stream)
            except Exception as ex:      # never lose the headline line over the extras
                line["configs_error"] = repr(ex)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", default="scan", choices=["scan", "shuffled"])
    ap.add_argument("--algo", default="binned", choices=["binned", "sweep", "direct"],
                    help="binned: the two-pass product path; sweep: experimental single-pass kernel with the two-pass kernels "
                         "as fall-back (N = 1); direct: global atomics (cross-check)")
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="N = 1: also measure the other BASELINE configs + the shuffled ordering (adds ~2 min)")
    ap.add_argument("--points", type=int, default=0, help="override points per GPU (debug)")
    ap.add_argument("--halo", type=int, default=32,
                    help="N > 1: halo rows per side (the scan jitter of the synthetic clouds strays <= 20 rows)")
    ap.add_argument("--gather", default="none", choices=["root", "all", "none"],
                    help="N > 1: leave the finished mosaic sharded, one strip per GPU (default: every rank delivers / writes its own "
                         "crops, as the multi-GPU offline converter does), or assemble it on rank 0 (root) / on every rank (all)")
    ap.add_argument("--gather-parts", type=int, default=1,
                    help="N > 1, --gather root: split the gather into this many row blocks on separate communicators")
    ap.add_argument("--cpu-points", type=int, default=0,
                    help="points of the CPU sample (0: the full config for --impl reference, 20 M for the in-line cpu_baseline)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
