#!/usr/bin/env python
"""bench.py -- BEV raster throughput (Mpoints/s) on N B200s, one JSON line.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (numpy port)

A *step* is one pass of the hot path over one synthetic cloud.
  N = 1   BASELINE.json configs[1]: 100 M-point road segment -> 11520 x 1152 x 3 u8 at 0.05 m.
  N > 1   configs[2] geometry, weak scaling: every rank owns one 1440-row strip of an
          (1440 N) x 11520 scene with 1.25e8 points (N = 8 is exactly the 1 B-point scene);
          the step includes the NCCL halo merge and the mosaic gather on rank 0 (--gather all: all-gather
          on every rank); the exchange and the gather of scene k run on a side stream under the
          rasterisation of scene k+1 and all of them finish inside the timed region.
``value``  device-resident throughput (inputs in HBM when the clock starts), CUDA events,
           max over ranks.   ``e2e``: the same metric through the host-buffer API
           (pinned H2D of the points + D2H of the finished raster inside the timed region).
``roofline`` is for the dominant kernel (bin_points), timed live with events between the
           pipeline stages of the same timed steps.
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
                f"{n / 1e6:.0f}M points per GPU, halo merge + mosaic gather in the step")
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec, n_full, name = workload(args.gpus, 0, args.points)
    n_sample = min(n_full, args.cpu_points)
    sub, cloud, what = cpu_sample(spec, n_sample, n_full * args.gpus)
    n_sample = len(cloud)
    P = os.cpu_count() or 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_rasterize_timed(cloud, sub, P, 1)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_rasterize_timed(cloud, sub, P, 1)
    dt = (time.perf_counter() - t0) / steps
    v = n_sample / dt / 1e6
    sample = f"{what}; numpy oracle (floor keys, bincount, maximum.at) over multiprocessing.Pool({P}) row strips"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
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
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    spec, n_pts, name = workload(args.gpus, rank, args.points)

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
    if args.gpus == 1 and rank == 0 and not args.no_cpu_baseline:
        P = os.cpu_count() or 1
        n_sample = min(n_pts, args.cpu_points)
        sub, sample_cloud, what = cpu_sample(spec, n_sample, n_pts)
        n_sample = len(sample_cloud)
        dt = cpu_rasterize_timed(sample_cloud, sub, P, 2)
        cpu_baseline = {"value": round(n_sample / dt / 1e6, 3), "unit": UNIT, "cores": P, "kind": "port",
                        "sample": f"{what}, numpy oracle over multiprocessing.Pool({P}) row strips, best of 2"}
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
    if args.gpus == 1:
        r = BevRasterizer(spec, n_pts, device=dev, algo=args.algo, outputs=("image",))
        out = r.alloc_outputs()
        step = lambda: r(pts, out=out)
        # the same step, split at the stage boundaries with events between (identical work)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]

        def staged(i):
            e = evs[i]
            e[0].record(stream)
            r(pts, out=out, stages=_cabi.STAGE_BIN)
            e[1].record(stream)
            r(pts, out=out, stages=_cabi.STAGE_INDEX)
            e[2].record(stream)
            r(pts, out=out, stages=_cabi.STAGE_REDUCE)
            e[3].record(stream)
        launches_per_step = 4      # bin_points, scan_tiles, index_chunks, reduce_tiles (+1 memset node)
        if args.overlap:
            # EXPERIMENTAL throughput mode: bin_points of scene k+1 under reduce_tiles of
            # scene k.  The grids must leave room for each other on an SM; the staged pass below (same
            # kernels, sequential) only feeds the roofline entry and runs after the timed region.
            os.environ.setdefault("LM_BEV_BIN_CTAS_PER_SM", "2")
            os.environ.setdefault("LM_BEV_RED_CTAS_PER_SM", "1")
            from lanemapping_b200.bev import PipelinedRasterizer
            pr = PipelinedRasterizer(spec, n_pts, device=dev, outputs=("image",))
            step = lambda: pr.submit(pts)
    else:
        from lanemapping_b200.strips import StripRasterizer
        sr = StripRasterizer(spec, n_pts, halo=args.halo, device=dev, align=STRIP_ALIGN,
                             gather_root=0 if args.gather == "root" else None)
        mosaic_holder = {}

        def step():
            # scene k's mosaic all-gather (side stream) overlaps scene k+1's rasterisation; every gather
            # completes inside the timed region (sr.flush() before the closing event)
            mosaic_holder["slot"] = sr.step(pts)
        staged = None
        launches_per_step = 4 + 2 * 2 + 2     # raster + merge/finalize per neighbour + pack copies (interior rank)

    for _ in range(Wm):
        step()
    barrier()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    overlap = args.gpus == 1 and args.overlap
    for i in range(K):
        if staged is not None and not overlap:
            staged(i)
        else:
            step()
    if args.gpus > 1:
        sr.flush()
    if overlap:
        pr.flush()
    t_stop.record(stream)
    barrier()
    ms_total = t_start.elapsed_time(t_stop)
    if overlap:
        pr.check_device_errors()
        for i in range(K):
            staged(i)
        torch.cuda.synchronize()
    if staged is not None:
        stage_ms = [float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(K)])) for j in range(3)]
    if args.gpus == 1:
        r.check_device_errors()
        n_valid = r.stats()["n_valid"]
    else:
        sr.raster.check_device_errors()
        n_valid = sr.raster.stats()["n_valid"]

    # ---- e2e through the host-buffer API: pinned H2D + kernels (+ collectives) + D2H
    if args.gpus == 1:
        hr = HostRasterizer(spec, n_pts, device=dev, algo=args.algo, outputs=("image",))
        e2e_step = lambda: hr(host_pts)
        d2h = spec.cells * spec.n_channels
    else:
        dev_in = torch.empty_like(pts)
        r0, r1 = sr.plan.strip
        host_strip = torch.empty((r1 - r0, spec.width, spec.n_channels), dtype=torch.uint8, pin_memory=True)

        def e2e_step():
            dev_in.copy_(host_pts, non_blocking=True)
            slot = sr.step(dev_in)
            sr.mosaic(slot)                                   # the scene's merge + gather are part of the step
            host_strip.copy_(sr.strip_of(slot), non_blocking=True)
            stream.synchronize()
        d2h = host_strip.numel()
    Ke = max(1, min(K, args.e2e_steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

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
            "config": {"workload": name, "order": args.order, "algo": args.algo,
                       **({"pipeline": "experimental: bin_points(k+1) on one stream under index+reduce_tiles(k) on another; "
                                       "stage_ms from a sequential pass after the timed region"} if args.gpus == 1 and args.overlap else {}),
                       **({"mosaic": "gathered on rank 0" if args.gather == "root" else "all-gathered on every rank"} if args.gpus > 1 else {}),
                       "points_per_gpu": n_pts, "valid_points_rank0": int(n_valid),
                       "l2_policy": "inputs (1.6+ GB of points per step) exceed the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the launch stream, barrier+synchronize both sides, max over ranks"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": n_pts * 16,
                    "d2h_bytes_per_step": int(d2h), "steps": Ke},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
        }
        if stage_ms is not None:
            # dominant kernel = bin_points: every point record (16 B) read exactly once
            b_alg_kernel = 16.0 * n_pts
            achieved = b_alg_kernel / (stage_ms[0] * 1e-3) / 1e9
            b_alg_path = float(spec.algorithmic_bytes(n_pts))
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                try:
                    with open(tp) as f:
                        traffic = json.load(f).get("bin_points_dram_bytes_per_launch_100M")
                    if n_pts != 100_000_000:
                        traffic = None
                except Exception:
                    traffic = None
            line["roofline"] = {
                "bound": "hbm", "kernel": "bin_points_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_alg_kernel,
                "stage_ms": {"bin_points(+memset)": round(stage_ms[0], 4), "scan+index": round(stage_ms[1], 4),
                             "reduce_tiles": round(stage_ms[2], 4)},
                "path_algorithmic_bytes": b_alg_path,
                "path_achieved": round(b_alg_path / (ms_step * 1e-3) / 1e9, 1),
                "path_frac": round(b_alg_path / (ms_step * 1e-3) / 1e9 / peak, 4),
            }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
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
    ap.add_argument("--algo", default="binned", choices=["binned", "direct"])
    ap.add_argument("--points", type=int, default=0, help="override points per GPU (debug)")
    ap.add_argument("--halo", type=int, default=64)
    ap.add_argument("--gather", default="root", choices=["root", "all"],
                    help="N > 1: assemble the mosaic on rank 0 (gather) or on every rank (all-gather)")
    ap.add_argument("--cpu-points", type=int, default=20_000_000, help="bounded sample for the CPU arm")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", action="store_true",
                    help="N = 1, experimental: pipeline consecutive scenes (bin_points of k+1 under reduce_tiles of k)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
