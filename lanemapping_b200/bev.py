"""PyTorch-facing host API of the BEV rasteriser.

PyTorch is plumbing here: it owns device memory and streams.  All arithmetic runs in
``csrc/liblm_bev.so`` through the C-ABI of ``include/lm_bev.h``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import _cabi
from .spec import ACC_PLANES, BevSpec

# "binned": the two-pass pipeline (the product path); "direct": global atomics, cross-check; "sweep": EXPERIMENTAL
# single-pass kernel (no record pool in HBM) with the binned kernels queued behind it as the exact fall-back
# (include/lm_bev.h LM_ALGO_SWEEP) -- correct on every input, HBM traffic 1.0 x algorithmic, but slower than
# "binned" today (profiles/README.md)
ALGOS = {"binned": _cabi.ALGO_BINNED, "direct": _cabi.ALGO_DIRECT, "sweep": _cabi.ALGO_SWEEP}
OUTPUT_KEYS = ("image", "count16", "proj", "acc")


def workspace_bytes(spec: BevSpec, n_points: int, algo: str = "binned",
                    outputs: Optional[Iterable[str]] = None, acc_band: int = 0) -> int:
    """Device workspace for one call; ``outputs`` (the buffers that will be requested) tightens it."""
    p = _cabi.make_params(spec)
    out = C.c_size_t(0)
    o = None
    if outputs is not None:
        o = _cabi.LmBevOutputs()
        for k in outputs:                      # only non-NULL-ness (and the band) matters for sizing
            setattr(o, k + "_dev", 1)
        o.acc_band = int(acc_band)
    _cabi.check(_cabi.lib().lm_bev_workspace_bytes(C.byref(p), int(n_points), ALGOS[algo],
                                                   C.byref(o) if o is not None else None, C.byref(out)))
    return int(out.value)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class BevRasterizer:
    """A spec + a reusable device workspace sized for ``max_points``.

    ``outputs`` picks the buffers to produce: ``image`` u8 [H,W,C] (the cropped_tiff pixel
    layout), ``count16`` u16 [H,W], ``proj`` f32 [C,H,W] (= the loader's
    ``to_tensor(img).float()``, reference baseline/datasets/laserlane_proposals.py:88-89)
    and ``acc`` u32 [6,H,W] raw accumulators (strip-halo merges).
    """

    def __init__(self, spec: BevSpec, max_points: int, device: torch.device | str = "cuda",
                 algo: str = "binned", outputs: Iterable[str] = ("image",), acc_band: int = 0,
                 tuning: Optional[Dict[str, int]] = None, graph: bool = False):
        """``tuning``: fields of ``lm_bev_tuning`` (include/lm_bev.h), e.g. ``{"bin_ctas_per_sm": 3}``; ``graph=True``:
        a call that repeats the previous call's buffers and point count replays a captured CUDA graph (one launch
        instead of a memset and 4..6 kernels: what bounds the small configs)."""
        if algo not in ALGOS:
            raise ValueError(f"algo must be one of {sorted(ALGOS)}")
        outputs = tuple(outputs)
        if not outputs or any(o not in OUTPUT_KEYS for o in outputs):
            raise ValueError(f"outputs must be a non-empty subset of {OUTPUT_KEYS}")
        if "count16" in outputs and not spec.count16:
            raise ValueError("count16 output requested but spec.count16 is False")
        self.spec = spec
        self.algo = algo
        self.outputs = outputs
        self.acc_band = int(acc_band)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("lanemapping_b200 runs on CUDA devices only (no CPU fallback)")
        self.max_points = int(max_points)
        self._params = _cabi.make_params(spec)
        self._lib = _cabi.lib()
        self._tuning = _cabi.LmBevTuning()
        for k, v in (tuning or {}).items():
            if not hasattr(self._tuning, k) or k == "reserved":
                raise ValueError(f"unknown tuning field {k!r}")
            setattr(self._tuning, k, int(v))
        if graph:
            self._tuning.use_graph = 1
        self._plan = C.c_void_p()
        o = self._sizing_outputs()
        _cabi.check(self._lib.lm_bev_plan_create(C.byref(self._params), self.max_points, ALGOS[algo], C.byref(o),
                                                 C.byref(self._tuning), C.byref(self._plan)))
        nbytes = C.c_size_t(0)
        _cabi.check(self._lib.lm_bev_plan_workspace_bytes(self._plan, C.byref(nbytes)))
        self.workspace = torch.empty(int(nbytes.value), dtype=torch.uint8, device=self.device)
        if algo == "sweep":
            # the sweep's mailboxes keep state between calls: prepare them once (include/lm_bev.h)
            with torch.cuda.device(self.device):
                _cabi.check(self._lib.lm_bev_plan_init_workspace(self._plan, self.workspace.data_ptr(), self.workspace.numel(),
                                                                 torch.cuda.current_stream(self.device).cuda_stream))

    def __del__(self):
        plan, lib = getattr(self, "_plan", None), getattr(self, "_lib", None)
        if plan and lib is not None:
            try:
                lib.lm_bev_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None

    def _sizing_outputs(self) -> "_cabi.LmBevOutputs":
        o = _cabi.LmBevOutputs()
        for k in self.outputs:                      # only non-NULL-ness (and the band) matters for the layout
            setattr(o, k + "_dev", 1)
        o.acc_band = self.acc_band
        return o

    def sweep_state(self) -> dict:
        """``algo="sweep"`` diagnostics (synchronises): calls done by the sweep / fallen back on this workspace."""
        if self.algo != "sweep":
            return {}
        off = C.c_size_t(0)
        _cabi.check(self._lib.lm_bev_sweep_state_offset(self.workspace.numel(), C.byref(off)))
        v = self.workspace[off.value:off.value + 16].cpu().numpy().view(np.uint32)
        return {"cooldown": int(v[1]), "n_failed": int(v[2]), "n_ok": int(v[3])}

    def sweep_debug(self) -> list:
        """Event counters of a -DLM_SWEEP_DEBUG build (zeros otherwise); see csrc/lm_sweep.cuh SW_DBG."""
        off = C.c_size_t(0)
        _cabi.check(self._lib.lm_bev_sweep_state_offset(self.workspace.numel(), C.byref(off)))
        return [int(x) for x in self.workspace[off.value + 16:off.value + 16 + 96].cpu().numpy().view(np.uint64)]

    # -- buffers ----------------------------------------------------------------------------
    def alloc_outputs(self) -> Dict[str, torch.Tensor]:
        H, W, Cn = self.spec.height, self.spec.width, self.spec.n_channels
        out: Dict[str, torch.Tensor] = {}
        if "image" in self.outputs:
            out["image"] = torch.empty((H, W, Cn), dtype=torch.uint8, device=self.device)
        if "count16" in self.outputs:
            out["count16"] = torch.empty((H, W), dtype=torch.uint16, device=self.device)
        if "proj" in self.outputs:
            out["proj"] = torch.empty((Cn, H, W), dtype=torch.float32, device=self.device)
        if "acc" in self.outputs:
            # rows outside the requested band are never written: start from the empty value
            out["acc"] = torch.zeros((ACC_PLANES, H, W), dtype=torch.int32, device=self.device)
        return out

    # -- the call ---------------------------------------------------------------------------
    def __call__(self, points: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None,
                 stream: Optional[torch.cuda.Stream] = None, stages: int = _cabi.STAGE_ALL) -> Dict[str, torch.Tensor]:
        """Enqueue one rasterisation of ``points`` (f32 [N,4] contiguous, on this device)."""
        if points.device != self.workspace.device and points.device.index != self.workspace.device.index:
            raise ValueError("points must live on the rasteriser's device")
        if points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] != 4 or not points.is_contiguous():
            raise ValueError("points must be a contiguous float32 [N,4] tensor (x, y, z, intensity)")
        n = int(points.shape[0])
        if n > self.max_points:
            raise ValueError(f"{n} points > max_points={self.max_points} this workspace was sized for")
        if out is None:
            out = self.alloc_outputs()
        o = _cabi.LmBevOutputs()
        o.image_dev = _ptr(out.get("image"))
        o.count16_dev = _ptr(out.get("count16"))
        o.proj_dev = _ptr(out.get("proj"))
        o.acc_dev = _ptr(out.get("acc"))
        o.acc_band = self.acc_band
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            if int(stages) == _cabi.STAGE_ALL:
                _cabi.check(self._lib.lm_bev_plan_rasterize(
                    self._plan, points.data_ptr() if n else None, n, self.workspace.data_ptr(), self.workspace.numel(),
                    C.byref(o), st.cuda_stream))
            else:       # stage-split calls (per-kernel timing): the plain entry point, default tuning
                _cabi.check(self._lib.lm_bev_rasterize_stages(
                    C.byref(self._params), points.data_ptr() if n else None, n, ALGOS[self.algo],
                    self.workspace.data_ptr(), self.workspace.numel(), C.byref(o), st.cuda_stream, int(stages)))
        return out

    def rasterize_las(self, records: torch.Tensor, n_points: int, xform, out: Optional[Dict[str, torch.Tensor]] = None,
                      stream: Optional[torch.cuda.Stream] = None) -> Dict[str, torch.Tensor]:
        """Rasterise straight from the point-data block of a LAS file (``lm_bev_rasterize_las``):
        ``records`` is a uint8 device tensor of ``n_points * record_length`` bytes as on disk,
        ``xform`` comes from :func:`las_xform`.  The decode runs inside the first kernel."""
        if self.algo != "binned":
            raise ValueError("rasterize_las runs the binned path only")
        n = int(n_points)
        if records.dtype != torch.uint8 or not records.is_contiguous() or records.numel() < n * xform.record_length:
            raise ValueError("records must be a contiguous uint8 tensor of n_points * record_length bytes")
        if n > self.max_points:
            raise ValueError(f"{n} points > max_points={self.max_points} this workspace was sized for")
        if out is None:
            out = self.alloc_outputs()
        o = _cabi.LmBevOutputs()
        o.image_dev = _ptr(out.get("image"))
        o.count16_dev = _ptr(out.get("count16"))
        o.proj_dev = _ptr(out.get("proj"))
        o.acc_dev = _ptr(out.get("acc"))
        o.acc_band = self.acc_band
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lm_bev_rasterize_las(
                C.byref(self._params), records.data_ptr() if n else None, n, C.byref(xform),
                self.workspace.data_ptr(), self.workspace.numel(), C.byref(o), st.cuda_stream))
        return out

    def stats(self) -> dict:
        """Device-side counters of the last call (synchronises)."""
        raw = self.workspace[:C.sizeof(_cabi.LmBevStats)].cpu().numpy().tobytes()
        s = _cabi.LmBevStats.from_buffer_copy(raw)
        return {"error": int(s.error), "n_chunks": int(s.n_chunks), "n_valid": int(s.n_valid),
                "n_tiles": int(s.n_tiles), "ct_overflow": int(s.ct_overflow)}

    def check_device_errors(self) -> None:
        s = self.stats()
        if s["error"] & _cabi.DEV_ERR_POOL:
            raise RuntimeError("liblm_bev: record chunk pool exhausted (workspace too small)")
        if s["error"] & _cabi.DEV_ERR_CELL_OVERFLOW:
            raise RuntimeError("liblm_bev: a cell received >= 2^24 points; u32 sums may have wrapped")


class BatchRasterizer:
    """B equally-shaped rasters per call (``lm_bev_rasterize_batch``): the samples of a DataLoader
    batch are stacked along the rows inside the library and share one set of launches.

    ``spec`` gives the common shape / resolutions / channels; sample ``b`` takes its origin from
    ``specs[b]`` (``bev_img_offset``, ``local_min_ele``, ``row0``, ``col0``) when given.
    Outputs: ``image`` u8 [B,H,W,C], ``count16`` u16 [B,H,W], ``proj`` f32 [B,C,H,W].
    """

    def __init__(self, spec: BevSpec, batch: int, max_points_total: int, device: torch.device | str = "cuda",
                 outputs: Iterable[str] = ("proj",)):
        outputs = tuple(outputs)
        if not outputs or any(o not in ("image", "count16", "proj") for o in outputs):
            raise ValueError("outputs must be a non-empty subset of ('image', 'count16', 'proj')")
        self.spec, self.batch, self.outputs = spec, int(batch), outputs
        self.max_points_total = int(max_points_total)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("lanemapping_b200 runs on CUDA devices only (no CPU fallback)")
        self._lib = _cabi.lib()
        self._params = _cabi.make_params(spec)
        o = _cabi.LmBevOutputs()
        for k in outputs:
            setattr(o, k + "_dev", 1)
        nbytes = C.c_size_t(0)
        _cabi.check(self._lib.lm_bev_workspace_bytes_batch(C.byref(self._params), self.batch, self.max_points_total,
                                                           C.byref(o), C.byref(nbytes)))
        self.workspace = torch.empty(int(nbytes.value), dtype=torch.uint8, device=self.device)

    def alloc_outputs(self) -> Dict[str, torch.Tensor]:
        B, H, W, Cn = self.batch, self.spec.height, self.spec.width, self.spec.n_channels
        out: Dict[str, torch.Tensor] = {}
        if "image" in self.outputs:
            out["image"] = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=self.device)
        if "count16" in self.outputs:
            out["count16"] = torch.empty((B, H, W), dtype=torch.uint16, device=self.device)
        if "proj" in self.outputs:
            out["proj"] = torch.empty((B, Cn, H, W), dtype=torch.float32, device=self.device)
        return out

    def __call__(self, points, specs=None, out: Optional[Dict[str, torch.Tensor]] = None,
                 stream: Optional[torch.cuda.Stream] = None) -> Dict[str, torch.Tensor]:
        B = len(points)
        if B < 1 or B > self.batch:
            raise ValueError(f"expected 1..{self.batch} clouds, got {B}")
        if specs is not None and len(specs) != B:
            raise ValueError("specs must have one entry per cloud")
        geoms = (_cabi.LmBevSampleGeom * B)()
        ptrs = (C.c_void_p * B)()
        counts = (C.c_int64 * B)()
        total = 0
        for b, pts in enumerate(points):
            if pts.device.type != "cuda" or pts.device.index != self.workspace.device.index:
                raise ValueError("points must live on the rasteriser's device")
            if pts.dtype != torch.float32 or pts.dim() != 2 or pts.shape[1] != 4 or not pts.is_contiguous():
                raise ValueError("points must be contiguous float32 [N,4] tensors (x, y, z, intensity)")
            sp = self.spec if specs is None else specs[b]
            if specs is not None and (sp.height, sp.width, sp.img_reso, sp.ele_reso, sp.channels, sp.inten_min,
                                      sp.inten_max) != (self.spec.height, self.spec.width, self.spec.img_reso,
                                                        self.spec.ele_reso, self.spec.channels, self.spec.inten_min,
                                                        self.spec.inten_max):
                raise ValueError("batched samples must share shape, resolutions, channels and intensity range")
            geoms[b].bev_img_offset[0], geoms[b].bev_img_offset[1] = sp.bev_img_offset
            geoms[b].local_min_ele = sp.local_min_ele
            geoms[b].row0, geoms[b].col0 = sp.row0, sp.col0
            n = int(pts.shape[0])
            ptrs[b] = pts.data_ptr() if n else None
            counts[b] = n
            total += n
        if total > self.max_points_total:
            raise ValueError(f"{total} points > max_points_total={self.max_points_total} this workspace was sized for")
        if out is None:
            out = self.alloc_outputs()
            out = {k: v[:B] for k, v in out.items()}
        o = _cabi.LmBevOutputs()
        o.image_dev = _ptr(out.get("image"))
        o.count16_dev = _ptr(out.get("count16"))
        o.proj_dev = _ptr(out.get("proj"))
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.lm_bev_rasterize_batch(C.byref(self._params), B, geoms, ptrs, counts,
                                                         self.workspace.data_ptr(), self.workspace.numel(),
                                                         C.byref(o), st.cuda_stream))
        return out

    def stats(self) -> dict:
        raw = self.workspace[:C.sizeof(_cabi.LmBevStats)].cpu().numpy().tobytes()
        s = _cabi.LmBevStats.from_buffer_copy(raw)
        return {"error": int(s.error), "n_chunks": int(s.n_chunks), "n_valid": int(s.n_valid),
                "n_tiles": int(s.n_tiles), "ct_overflow": int(s.ct_overflow)}


def las_xform(header, params=None):
    """The decode parameters of one LAS file: ``header`` is a :class:`lanemapping_b200.las.LasHeader`,
    ``params`` the crop's :class:`lanemapping_b200.sidecar.PcImgParams` (None = world frame: no read
    offset, identity rotation).  rot = R(q)^T, the inverse of reference
    baseline/utils/coor_img2pc.py:163-171."""
    if params is None:
        return _cabi.make_las_xform(header.record_length, header.scale, header.offset)
    from .sidecar import quat_to_matrix
    rot = quat_to_matrix(params.las_rotation_trans_quan[3:]).T.reshape(9)
    return _cabi.make_las_xform(header.record_length, header.scale, header.offset, params.las_read_offset,
                                params.las_rotation_trans_quan[:3], rot)


def decode_las(records: torch.Tensor, n_points: int, xform, out: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """LAS point records (uint8 device tensor, as on disk) -> float32 [n,4] (x, y, z, intensity) in the
    raster-local frame (``lm_las_decode``): the GPU replacement of the reference's ``read_las``
    (baseline/datasets/laserlane_proposals.py:618-636) up to its intensity normalisation."""
    n = int(n_points)
    if records.device.type != "cuda":
        raise RuntimeError("lanemapping_b200 runs on CUDA devices only (no CPU fallback)")
    if records.dtype != torch.uint8 or not records.is_contiguous() or records.numel() < n * xform.record_length:
        raise ValueError("records must be a contiguous uint8 tensor of n_points * record_length bytes")
    if out is None:
        out = torch.empty((n, 4), dtype=torch.float32, device=records.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (n, 4) or not out.is_contiguous() or out.device != records.device:
        raise ValueError("out must be a contiguous float32 [n,4] tensor on the records' device")
    st = stream if stream is not None else torch.cuda.current_stream(records.device)
    with torch.cuda.device(records.device):
        _cabi.check(_cabi.lib().lm_las_decode(records.data_ptr() if n else None, n, C.byref(xform),
                                              out.data_ptr() if n else None, st.cuda_stream))
    return out


def rasterize(points: torch.Tensor, spec: BevSpec, algo: str = "binned",
              outputs: Iterable[str] = ("image",)) -> Dict[str, torch.Tensor]:
    """One-shot convenience wrapper (allocates a workspace for this call)."""
    r = BevRasterizer(spec, int(points.shape[0]), device=points.device, algo=algo, outputs=outputs)
    return r(points)


# ---------------------------------------------------------------------------------------------
# helpers on finished rasters
# ---------------------------------------------------------------------------------------------
def acc_merge_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst <- merge(dst, src) for two accumulator sets [6,rows,W] (views into bigger [6,H,W]
    buffers are fine as long as each plane's rows are contiguous)."""
    if dst.shape != src.shape or dst.dim() != 3 or dst.shape[0] != ACC_PLANES:
        raise ValueError("acc_merge_: expected two [6,rows,W] tensors of equal shape")
    for t in (dst, src):
        if t.stride(2) != 1 or t.stride(1) != t.shape[2] or t.dtype != torch.int32:
            raise ValueError("acc_merge_: planes must be int32 with contiguous rows")
    st = torch.cuda.current_stream(dst.device)
    with torch.cuda.device(dst.device):
        _cabi.check(_cabi.lib().lm_bev_acc_merge(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0),
                                                 dst.shape[1], dst.shape[2], st.cuda_stream))
    return dst


def finalize_rows(spec: BevSpec, acc: torch.Tensor, row_begin: int, row_end: int,
                  out: Dict[str, torch.Tensor]) -> None:
    """Accumulators [6,H,W] -> the buffers in ``out`` for rows [row_begin,row_end)."""
    if tuple(acc.shape) != (ACC_PLANES, spec.height, spec.width) or not acc.is_contiguous():
        raise ValueError("finalize_rows: acc must be contiguous [6,H,W]")
    p = _cabi.make_params(spec)
    o = _cabi.LmBevOutputs()
    o.image_dev = _ptr(out.get("image"))
    o.count16_dev = _ptr(out.get("count16"))
    o.proj_dev = _ptr(out.get("proj"))
    st = torch.cuda.current_stream(acc.device)
    with torch.cuda.device(acc.device):
        _cabi.check(_cabi.lib().lm_bev_finalize(C.byref(p), acc.data_ptr(), int(row_begin), int(row_end),
                                                C.byref(o), st.cuda_stream))


def needed_planes(spec: BevSpec) -> List[int]:
    """The raw accumulator planes (ACC_* ids, ascending) the spec's channels are derived from: the only
    ones a strip has to send to its neighbour (``merge_finalize_rows``)."""
    from .spec import (ACC_COUNT, ACC_MAX_I, ACC_MAX_Z, ACC_MIN_Z, ACC_SUM_I, ACC_SUM_Z, CH_DENSITY, CH_MAX_I, CH_MAX_Z,
                       CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z)
    need = {CH_MAX_I: (ACC_MAX_I,), CH_MEAN_I: (ACC_COUNT, ACC_SUM_I), CH_MIN_Z: (ACC_MIN_Z,), CH_MAX_Z: (ACC_MAX_Z,),
            CH_MEAN_Z: (ACC_COUNT, ACC_SUM_Z), CH_DENSITY: (ACC_COUNT,)}
    s = set()
    for c in spec.channels:
        s.update(need[c])
    if spec.count16:
        s.add(ACC_COUNT)
    return sorted(s)


def merge_finalize_rows(spec: BevSpec, acc: torch.Tensor, row_begin: int, row_end: int, recv: torch.Tensor,
                        planes, out: Dict[str, torch.Tensor]) -> None:
    """Halo band in one launch (``lm_bev_merge_finalize``): rows [row_begin,row_end) of ``acc`` [6,H,W] are merged with
    the neighbour's planes ``recv`` [len(planes), rows, W] (ascending ACC_* ids in ``planes``) and finished into ``out``."""
    rows = int(row_end) - int(row_begin)
    if tuple(acc.shape) != (ACC_PLANES, spec.height, spec.width) or not acc.is_contiguous():
        raise ValueError("merge_finalize_rows: acc must be contiguous [6,H,W]")
    planes = list(planes)
    if planes != sorted(set(planes)) or tuple(recv.shape) != (len(planes), rows, spec.width) or not recv.is_contiguous():
        raise ValueError("merge_finalize_rows: recv must be contiguous [len(planes), rows, W], planes ascending")
    mask = 0
    for pl in planes:
        mask |= 1 << int(pl)
    p = _cabi.make_params(spec)
    o = _cabi.LmBevOutputs()
    o.image_dev = _ptr(out.get("image"))
    o.count16_dev = _ptr(out.get("count16"))
    o.proj_dev = _ptr(out.get("proj"))
    st = torch.cuda.current_stream(acc.device)
    with torch.cuda.device(acc.device):
        _cabi.check(_cabi.lib().lm_bev_merge_finalize(C.byref(p), acc.data_ptr(), int(row_begin), int(row_end),
                                                      recv.data_ptr(), mask, C.byref(o), st.cuda_stream))


def crop_tiles(image: torch.Tensor, tile: int = 1152) -> torch.Tensor:
    """u8 [H,W,C] mosaic -> u8 [n_crops,tile,tile,C], row-major crop order, ragged edges zero."""
    if image.dtype != torch.uint8 or image.dim() != 3 or not image.is_contiguous():
        raise ValueError("crop_tiles: image must be contiguous uint8 [H,W,C]")
    H, W, Cn = image.shape
    n = (-(-H // tile)) * (-(-W // tile))
    crops = torch.empty((n, tile, tile, Cn), dtype=torch.uint8, device=image.device)
    st = torch.cuda.current_stream(image.device)
    with torch.cuda.device(image.device):
        _cabi.check(_cabi.lib().lm_bev_crop_tiles(image.data_ptr(), H, W, Cn, tile, crops.data_ptr(), st.cuda_stream))
    return crops


class HostRasterizer:
    """End-to-end entry with HOST buffers: pinned staging, H2D copy, kernels, D2H copy.

    This is the call an offline converter makes per file (see convert_data.py in this
    package) and what bench.py times as ``e2e``.
    """

    def __init__(self, spec: BevSpec, max_points: int, device: torch.device | str = "cuda",
                 algo: str = "binned", outputs: Iterable[str] = ("image",)):
        self.raster = BevRasterizer(spec, max_points, device=device, algo=algo, outputs=outputs)
        self.dev_points = torch.empty((max_points, 4), dtype=torch.float32, device=self.raster.device)
        self.dev_out = self.raster.alloc_outputs()
        self.host_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self.dev_out.items()}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, points_host: torch.Tensor) -> Dict[str, np.ndarray]:
        """points_host: float32 [N,4] CPU tensor (pinned for full copy speed)."""
        n = int(points_host.shape[0])
        dev = self.dev_points[:n]
        dev.copy_(points_host, non_blocking=True)
        self.raster(dev, out=self.dev_out)
        for k, v in self.dev_out.items():
            self.host_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream(self.raster.device).synchronize()
        self.h2d_bytes = n * 16
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.dev_out.values())
        return {k: v.numpy() for k, v in self.host_out.items()}
