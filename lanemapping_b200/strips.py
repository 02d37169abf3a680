"""Strip-sharded multi-GPU rasterisation.

A large scene is cut along the row / driving axis (reference data/convert_data.py:151-156:
lanes run low-row -> high-row) into ``world`` contiguous strips, one process per GPU.  Every
per-cell reduction (count, sums, max, min) is associative and commutative over integers, so
any partition of the points merges exactly.

* Points arrive pre-bucketed by *coarse* strip (a cheap pass on the along-track coordinate).
  A point whose exact row belongs to a neighbour lands in a **halo** band of ``halo`` rows;
  points farther out than the halo must not be given to this rank (they are dropped).
* Each rank rasterises the integer window ``[r0-halo, r1+halo)`` of the global grid -- same
  float origin, shifted integer window, so results are bit-identical to the one-piece raster.
* One exchange step: the raw u32 accumulators of the halo bands go to the neighbours
  (``batch_isend_irecv`` over NCCL/NVLink) -- only the planes the channel set is derived from, each sent
  straight from its rows of the accumulator buffer (no pack copy) -- and one kernel
  (``lm_bev_merge_finalize``) merges them into the neighbour's edge band and re-finishes that band *before*
  quantising to u8.
* Mosaic gather: the finished strips are gathered on one rank (``gather``, what an offline writer needs)
  or on all of them (``all_gather``) -- 398 MB for config 3.

The arithmetic back end is injected so that the host logic is testable on CPU with gloo
(tests pass an oracle-backed back end); the product back end is ``CudaBackend``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .spec import ACC_PLANES, BevSpec


def strip_bounds(height: int, world: int, align: int = 128) -> List[Tuple[int, int]]:
    """Row ranges of the ``world`` strips.  Interior boundaries are multiples of ``align`` (the
    shared-memory tile height) so that no tile straddles two ranks; strips are as equal as that
    allows (config 3: 11520 rows / 8 = 1440 = 11.25 x 128 -> strips of 1408/1536 rows)."""
    if world < 1 or height < 1:
        raise ValueError("strip_bounds: world and height must be >= 1")
    edges = [0]
    for k in range(1, world):
        e = int(round(height * k / world / align)) * align
        edges.append(min(max(e, edges[-1]), height))
    edges.append(height)
    return [(edges[k], edges[k + 1]) for k in range(world)]


def coarse_strip_of(x_local, spec: BevSpec, bounds: Sequence[Tuple[int, int]]):
    """Coarse bucket of points by along-track coordinate (numpy or torch): index of the strip
    whose row range contains ``floor((x - off)/reso)`` computed in float64/whatever the caller
    has -- it only has to be right to within the halo."""
    import numpy as np
    row = np.floor((np.asarray(x_local, dtype=np.float64) - spec.bev_img_offset[0]) / spec.img_reso[0])
    row = row - spec.row0
    starts = np.array([b[0] for b in bounds[1:]], dtype=np.float64)
    return np.searchsorted(starts, row, side="right")


class CudaBackend:
    """Product back end: the C-ABI kernels."""

    def __init__(self, device):
        self.device = torch.device(device)

    def make(self, spec: BevSpec, max_points: int, outputs, acc_band: int):
        from .bev import BevRasterizer
        return BevRasterizer(spec, max_points, device=self.device, outputs=outputs, acc_band=acc_band)

    def planes(self, spec: BevSpec):
        from .bev import needed_planes
        return needed_planes(spec)

    def merge_finalize(self, spec, acc, r0, r1, recv, planes, out):
        from .bev import merge_finalize_rows
        merge_finalize_rows(spec, acc, r0, r1, recv, planes, out)


@dataclass
class StripPlan:
    rank: int
    world: int
    bounds: List[Tuple[int, int]]
    halo: int
    win0: int          # first global row of this rank's window (strip + halos, clipped to the scene)
    win1: int
    top: int           # halo rows actually present above / below the strip
    bottom: int

    @property
    def strip(self) -> Tuple[int, int]:
        return self.bounds[self.rank]


def make_plan(spec: BevSpec, rank: int, world: int, halo: int, align: int = 128) -> StripPlan:
    bounds = strip_bounds(spec.height, world, align)
    r0, r1 = bounds[rank]
    if world > 1 and halo > 0:
        for k, (a, b) in enumerate(bounds):
            if b - a < halo:
                raise ValueError(f"strip {k} has {b - a} rows < halo {halo}: use fewer ranks or a smaller halo")
    top = halo if rank > 0 else 0
    bottom = halo if rank < world - 1 else 0
    return StripPlan(rank, world, bounds, halo, r0 - top, r1 + bottom, top, bottom)


class StripRasterizer:
    """One rank's part of a strip-sharded rasterisation."""

    def __init__(self, spec: BevSpec, max_points: int, halo: int = 64, group=None, backend=None,
                 device: Optional[torch.device | str] = None, align: int = 128, gather_root=None,
                 time_stages: bool = False, gather_parts: int = 1):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.spec = spec
        self.gather_root = gather_root          # step(): None = mosaic on every rank, r = on rank r only, "none" = stays sharded
        self.time_stages = time_stages          # step(): record events around the stages (stage_times())
        self.gather_parts = max(1, int(gather_parts))   # step(): rooted gather split into row blocks on separate communicators
        self.plan = make_plan(spec, self.rank, self.world, halo, align)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.backend = backend if backend is not None else CudaBackend(self.device)
        p = self.plan
        self.local_spec = spec.window(p.win0, p.win1)
        self.need_acc = self.world > 1 and halo > 0
        outputs = ("image", "acc") if self.need_acc else ("image",)
        # raw accumulators are needed on the halo band and on the strip-edge band it merges into
        self.raster = self.backend.make(self.local_spec, max_points, outputs, 2 * halo if self.need_acc else 0)
        self.out = self.raster.alloc_outputs()
        # pipelined mode (step / flush): two output sets so that the gather of one scene overlaps the
        # rasterisation of the next one
        self._outs = [self.out, None]
        self._mosaics = [None, None]
        self._gather_done = [None, None]
        self._comm_stream = None
        self._gather_stream = None
        self._gather_group = None               # step(): the mosaic gather runs on its own communicator + stream, so that the
        self._step = 0                          # gather of scene k overlaps the halo exchange of scene k+1
        self._part_groups, self._part_streams = [], []
        W = spec.width
        # only the planes the channels are derived from travel (config 3: count, sum_z, max_i = 3 of 6)
        self.planes = list(self.backend.planes(spec)) if self.need_acc else []
        mk = lambda rows: torch.empty((len(self.planes), rows, W), dtype=torch.int32, device=self.device)
        self._recv_up = mk(p.top) if p.top else None
        self._recv_dn = mk(p.bottom) if p.bottom else None
        self.halo_bytes = len(self.planes) * (p.top + p.bottom) * W * 4          # sent per scene by this rank
        self._stage_events = []                                                   # step(): (raster, exchange, merge, gather) events

    # -- one step ---------------------------------------------------------------------------
    def rasterize(self, points: torch.Tensor) -> torch.Tensor:
        """Rasterise this rank's points (coarse bucket incl. halo strays), exchange + merge the
        halo accumulators, return the finished u8 strip [rows, W, C] (a view, no halo rows)."""
        p = self.plan
        out = self.raster(points, out=self.out)
        if self.need_acc:
            self._exchange_and_merge(out)
        hl = self.local_spec.height
        return out["image"][p.top:hl - p.bottom]

    def _exchange_and_merge(self, out: Dict[str, torch.Tensor], events=None) -> None:
        p = self.plan
        acc = out["acc"]
        hl = self.local_spec.height
        ops = []
        for k, pl in enumerate(self.planes):
            if p.top:       # my top halo rows belong to rank-1; its bottom halo covers my first rows
                ops.append(dist.P2POp(dist.isend, acc[pl, :p.top], self._peer(self.rank - 1), self.group))
                ops.append(dist.P2POp(dist.irecv, self._recv_up[k], self._peer(self.rank - 1), self.group))
            if p.bottom:
                ops.append(dist.P2POp(dist.isend, acc[pl, hl - p.bottom:], self._peer(self.rank + 1), self.group))
                ops.append(dist.P2POp(dist.irecv, self._recv_dn[k], self._peer(self.rank + 1), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if events is not None:
            events[0].record(torch.cuda.current_stream(self.device))
        if p.top:       # neighbour's bottom halo = my rows [top, 2*top)
            self.backend.merge_finalize(self.local_spec, acc, p.top, 2 * p.top, self._recv_up, self.planes, {"image": out["image"]})
        if p.bottom:
            self.backend.merge_finalize(self.local_spec, acc, hl - 2 * p.bottom, hl - p.bottom, self._recv_dn, self.planes,
                                        {"image": out["image"]})

    # -- pipelined steps: rasterise scene k while scene k-1's mosaic is still being gathered -----------
    def step(self, points: torch.Tensor) -> int:
        """Enqueue one full scene (rasterisation on the current stream; halo exchange + merge and the
        mosaic all-gather on a side stream) and return its buffer slot; ``mosaic(slot)`` waits for and returns the result.
        At most two scenes are in flight: slot k % 2 is reused by scene k + 2."""
        if self.device.type != "cuda":
            raise RuntimeError("step() needs CUDA streams; use rasterize() + gather() on CPU back ends")
        k = self._step % 2
        self._step += 1
        main = torch.cuda.current_stream(self.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(self.device)
        if self._outs[k] is None:
            self._outs[k] = self.raster.alloc_outputs()
        if self._gather_done[k] is not None:
            main.wait_event(self._gather_done[k])          # the strip buffer is free once its gather has run
        p = self.plan
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if self.time_stages else None
        if ev:
            ev[0].record(main)
        out = self.raster(points, out=self._outs[k])
        hl = self.local_spec.height
        strip = out["image"][p.top:hl - p.bottom]
        ready = torch.cuda.Event(enable_timing=self.time_stages)
        ready.record(main)
        if self._gather_stream is None:
            self._gather_stream = torch.cuda.Stream(self.device)
            if self.world > 1 and self.gather_root != "none":
                ranks = None if self.group is None else dist.get_process_group_ranks(self.group)
                self._gather_group = dist.new_group(ranks=ranks)
                rows = [b - a for a, b in self.plan.bounds]
                if self.gather_parts > 1 and isinstance(self.gather_root, int) and all(r == rows[0] for r in rows):
                    # NCCL's send/recv fan-in runs on a few CTAs per communicator: several communicators, each moving one
                    # row block of every strip on its own stream, put more of them on the links
                    self._part_groups = [self._gather_group] + [dist.new_group(ranks=ranks) for _ in range(self.gather_parts - 1)]
                    self._part_streams = [self._gather_stream] + [torch.cuda.Stream(self.device) for _ in range(self.gather_parts - 1)]
        with torch.cuda.stream(self._comm_stream):
            # after the local rasterisation: halo exchange, merge + re-finish of the edge bands on one side stream ...
            self._comm_stream.wait_event(ready)
            if ev:
                ev[1].record(self._comm_stream)
            if self.need_acc:
                self._exchange_and_merge(out, ev[2:3] if ev else None)
            elif ev:
                ev[2].record(self._comm_stream)
            merged = torch.cuda.Event(enable_timing=self.time_stages)
            merged.record(self._comm_stream)
        with torch.cuda.stream(self._gather_stream):
            # ... and the mosaic gather on another one with its own communicator: all of it under the next scenes' rasterisation
            self._gather_stream.wait_event(merged)
            if self.gather_root == "none":
                self._mosaics[k] = None
            elif self._part_groups:
                self._mosaics[k] = self._gather_in_parts(strip, merged)
            else:
                grp, self.group = self.group, (self._gather_group if self._gather_group is not None else self.group)
                try:
                    self._mosaics[k] = self.gather(strip, self.gather_root)
                finally:
                    self.group = grp
            done = torch.cuda.Event(enable_timing=self.time_stages)
            done.record(self._gather_stream)
        if ev:
            ev[3] = merged
        if ev:
            self._stage_events.append((ev[0], ready, ev[1], ev[2], ev[3], done))
            del self._stage_events[:-64]
        self._gather_done[k] = done
        return k

    def reset_stage_times(self) -> None:
        """Forget the recorded steps (call after warm-up: the first steps carry NCCL's connection set-up)."""
        torch.cuda.synchronize(self.device)
        self._stage_events.clear()

    def stage_times(self) -> Dict[str, float]:
        """Mean ms of the stages of the last ``step()`` calls (``time_stages=True``; synchronises): local
        rasterisation (main stream), then on the side stream halo exchange, merge + re-finish, mosaic gather."""
        if not self._stage_events:
            return {}
        torch.cuda.synchronize(self.device)
        n = len(self._stage_events)
        acc = {"raster": 0.0, "halo_exchange": 0.0, "merge_finalize": 0.0, "gather": 0.0, "side_stream_total": 0.0}
        for e0, ready, c0, c1, c2, done in self._stage_events:
            acc["raster"] += e0.elapsed_time(ready)
            acc["halo_exchange"] += c0.elapsed_time(c1)
            acc["merge_finalize"] += c1.elapsed_time(c2)
            acc["gather"] += c2.elapsed_time(done)
            acc["side_stream_total"] += c0.elapsed_time(done)
        return {k: v / n for k, v in acc.items()}

    def verify(self, points: torch.Tensor, strip: torch.Tensor) -> int:
        """Independent check of a finished strip (bytes that differ): the global-atomic algorithm
        (``LM_ALGO_DIRECT``) on this rank's window, ALL six raw planes of the halo rows exchanged as packed
        copies, merged with plain torch integer ops (the merge law of SURVEY 8e), finished with ``lm_bev_finalize``."""
        from .bev import BevRasterizer, finalize_rows
        p, hl, W = self.plan, self.local_spec.height, self.spec.width
        d = BevRasterizer(self.local_spec, max(int(points.shape[0]), 1), device=self.device, algo="direct", outputs=("image", "acc"))
        o = d(points)
        acc = o["acc"]
        if self.need_acc:
            ops, bufs = [], {}
            for name, rows, peer in (("up", slice(0, p.top), self.rank - 1), ("dn", slice(hl - p.bottom, hl), self.rank + 1)):
                n = p.top if name == "up" else p.bottom
                if not n:
                    continue
                snd = acc[:, rows].contiguous()
                rcv = torch.empty_like(snd)
                bufs[name] = (snd, rcv)
                ops.append(dist.P2POp(dist.isend, snd, self._peer(peer), self.group))
                ops.append(dist.P2POp(dist.irecv, rcv, self._peer(peer), self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for name, (snd, rcv) in bufs.items():
                band = slice(p.top, 2 * p.top) if name == "up" else slice(hl - 2 * p.bottom, hl - p.bottom)
                a = acc[:, band].to(torch.int64) & 0xFFFFFFFF
                b = rcv.to(torch.int64) & 0xFFFFFFFF
                m = torch.empty_like(a)
                m[0:3] = a[0:3] + b[0:3]                       # count, sum_i, sum_z
                m[3] = torch.maximum(a[3], b[3])               # max_i
                m[4] = torch.minimum(a[4], b[4])               # min_z (0xFFFFFFFF where empty)
                m[5] = torch.maximum(a[5], b[5])               # max_z
                acc[:, band] = torch.where(m >= 2 ** 31, m - 2 ** 32, m).to(torch.int32)
                finalize_rows(self.local_spec, acc, band.start, band.stop, {"image": o["image"]})
        want = o["image"][p.top:hl - p.bottom]
        return int((want != strip).sum().item())

    def _gather_in_parts(self, strip: torch.Tensor, merged) -> Optional[torch.Tensor]:
        """Rooted gather of equal strips as ``gather_parts`` row blocks, one communicator + stream each (called on the
        first gather stream, which then waits for the others)."""
        root, P = self.gather_root, len(self._part_groups)
        is_root = self.rank == root
        rows = strip.shape[0]
        W, C = strip.shape[1], strip.shape[2]
        mosaic = torch.empty((self.spec.height, W, C), dtype=strip.dtype, device=strip.device) if is_root else None
        cuts = [rows * q // P for q in range(P + 1)]
        evs = []
        for q in range(P):
            st = self._part_streams[q]
            if q:
                st.wait_event(merged)
                if mosaic is not None:
                    mosaic.record_stream(st)
            with torch.cuda.stream(st):
                a, b = cuts[q], cuts[q + 1]
                parts = [mosaic[r * rows + a:r * rows + b] for r in range(self.world)] if is_root else None
                dist.gather(strip[a:b], parts, dst=self._peer(root), group=self._part_groups[q])
                if q:
                    e = torch.cuda.Event()
                    e.record(st)
                    evs.append(e)
        for e in evs:
            self._gather_stream.wait_event(e)
        return mosaic

    def close(self) -> None:
        """Destroy the gather communicator (before ``dist.destroy_process_group``)."""
        if self._gather_group is not None:
            torch.cuda.synchronize(self.device)
            dist.destroy_process_group(self._gather_group)
            self._gather_group = None

    def mosaic(self, slot: int) -> Optional[torch.Tensor]:
        """The scene's mosaic (None on the ranks a rooted gather leaves empty); the current stream waits
        for the scene's exchange, merge and gather."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._gather_done[slot])
        m = self._mosaics[slot]
        if m is not None:
            m.record_stream(cur)        # allocated on the side stream, read on this one
        return m

    def strip_of(self, slot: int) -> torch.Tensor:
        """This rank's own finished strip of the scene in ``slot`` (halo rows cut off, edge bands merged)."""
        torch.cuda.current_stream(self.device).wait_event(self._gather_done[slot])
        p = self.plan
        return self._outs[slot]["image"][p.top:self.local_spec.height - p.bottom]

    def flush(self) -> None:
        for ev in self._gather_done:
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)

    def _peer(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    # -- mosaic -----------------------------------------------------------------------------
    def gather(self, strip: torch.Tensor, root: Optional[int] = None) -> Optional[torch.Tensor]:
        """Assemble the [H, W, C] mosaic from the finished strips.  ``root=None``: every rank gets it
        (all-gather); ``root=r``: only rank ``r`` does (gather: the other ranks send their strip once and
        return None -- 1/world of the all-gather's traffic, which is what an offline writer needs)."""
        if self.world == 1:
            return strip
        rows = [b - a for a, b in self.plan.bounds]
        mx = max(rows)
        C = strip.shape[2]
        W = strip.shape[1]
        equal = all(r == mx for r in rows)
        if root is None:
            if equal:
                mosaic = torch.empty((self.spec.height, W, C), dtype=strip.dtype, device=strip.device)
                dist.all_gather_into_tensor(mosaic, strip.contiguous(), group=self.group)
                return mosaic
            pad = torch.zeros((mx, W, C), dtype=strip.dtype, device=strip.device)
            pad[:strip.shape[0]] = strip
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            return torch.cat([parts[k][:rows[k]] for k in range(self.world)], dim=0)
        if not 0 <= root < self.world:
            raise ValueError(f"gather: root {root} outside the group of {self.world}")
        is_root = self.rank == root
        if equal:
            mosaic = torch.empty((self.spec.height, W, C), dtype=strip.dtype, device=strip.device) if is_root else None
            # the strips land in place: row blocks of the mosaic are contiguous
            parts = list(mosaic.split(mx, dim=0)) if is_root else None
            dist.gather(strip.contiguous(), parts, dst=self._peer(root), group=self.group)
            return mosaic
        pad = torch.zeros((mx, W, C), dtype=strip.dtype, device=strip.device)
        pad[:strip.shape[0]] = strip
        parts = [torch.empty_like(pad) for _ in range(self.world)] if is_root else None
        dist.gather(pad, parts, dst=self._peer(root), group=self.group)
        return torch.cat([parts[k][:rows[k]] for k in range(self.world)], dim=0) if is_root else None
