"""PCENCODER plug-in: rasterise ``sample['points']`` on the GPU, then run the stock encoder.

Plug-in surface (reference baseline/models/registry.py:20-36): ``@PCENCODER.register_module
class Y(nn.Module): __init__(..., cfg=None)``; ``forward(sample) -> (fea, fea_up, bi_seg, endp)``
consumed at reference baseline/models/net/detector1stage.py:28.  The stock encoder
``PostProjector2.forward`` reads ``sample['proj']`` (reference
baseline/models/pcencoder/postprojector.py:79-82) and the heads read ``batch['proj']`` again for
overlays (reference baseline/models/heads/polyline_fpn_vit_vertex_2.py:956-959), so the raster is
stored back into ``sample['proj']``: f32 [B,3,1152,1152] = u8/255, exactly what
``to_tensor(PNG).float()`` gives (reference baseline/datasets/laserlane_proposals.py:88-89).
"""
from __future__ import annotations

from dataclasses import replace
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .spec import CH_DENSITY, CH_MAX_I, CH_MEAN_Z, TILE, BevSpec


class BatchProjector:
    """points list -> proj [B,C,tile,tile] on the GPU (BASELINE.json configs[4])."""

    def __init__(self, tile: int = TILE, channels: Sequence[int] = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY),
                 img_reso=(0.05, 0.05), ele_reso: float = 0.05):
        self.tile, self.channels, self.img_reso, self.ele_reso = tile, tuple(channels), tuple(img_reso), ele_reso
        self._rasters = {}

    def spec_for(self, geom: Optional[torch.Tensor]) -> BevSpec:
        if geom is None:
            return BevSpec(self.tile, self.tile, img_reso=self.img_reso, ele_reso=self.ele_reso, channels=self.channels)
        g = [float(v) for v in geom]
        row0, col0 = (int(g[6]), int(g[7])) if len(g) >= 8 else (0, 0)
        return BevSpec(self.tile, self.tile, bev_img_offset=(g[0], g[1]), img_reso=(g[2], g[3]), local_min_ele=g[4],
                       ele_reso=g[5], channels=self.channels, row0=row0, col0=col0)

    def __call__(self, points: List[torch.Tensor], geoms=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One batched C-ABI call (``lm_bev_rasterize_batch``): the B clouds share one set of launches."""
        from .bev import BatchRasterizer
        B = len(points)
        dev = points[0].device
        if dev.type != "cuda":
            raise RuntimeError("BatchProjector: points must already be on the GPU (Runner.to_cuda list branch)")
        C = len(self.channels)
        if out is None:
            out = torch.empty((B, C, self.tile, self.tile), dtype=torch.float32, device=dev)
        specs = [self.spec_for(None if geoms is None else geoms[b]) for b in range(B)]
        common = replace(specs[0], bev_img_offset=(0.0, 0.0), local_min_ele=0.0, row0=0, col0=0)
        pts = [p.contiguous() for p in points]
        total = sum(int(p.shape[0]) for p in pts)
        key = (common, dev, B)
        r = self._rasters.get(key)
        if r is None or r.max_points_total < total:
            if len(self._rasters) > 16:
                self._rasters.clear()
            r = BatchRasterizer(common, B, max(total, 1), device=dev, outputs=("proj",))
            self._rasters[key] = r
        r(pts, specs, out={"proj": out})
        return out


class OnTheFlyProjector(nn.Module):
    """Wraps a stock PCENCODER (e.g. PostProjector2): fills ``sample['proj']`` from
    ``sample['points']`` when it is missing, then delegates."""

    def __init__(self, inner: nn.Module, cfg=None, **proj_kwargs):
        super().__init__()
        self.inner = inner
        self.cfg = cfg
        self.projector = BatchProjector(**proj_kwargs)

    @staticmethod
    def clouds_of(sample) -> List[torch.Tensor]:
        """The per-sample [N_i, 4] clouds of a batch, whichever collate produced it:
        ``datasets.PointBatch`` (packed, ``collate_points``), a dense NaN-padded [B, N_max, 4] tensor with
        ``points_count`` (``collate_points_padded``: the form ``nn.DataParallel`` can split), or a list."""
        pts = sample["points"]
        if hasattr(pts, "clouds"):
            return pts.clouds()
        if isinstance(pts, torch.Tensor):
            if pts.ndim != 3 or pts.shape[-1] != 4:
                raise ValueError("sample['points']: a dense batch must be [B, N_max, 4]")
            cnt = sample.get("points_count")
            if cnt is None:
                return [pts[b] for b in range(pts.shape[0])]        # NaN padding is dropped by the rasteriser
            return [pts[b, :int(cnt[b])] for b in range(pts.shape[0])]
        return [p.data if hasattr(p, "data") else p for p in pts]

    def forward(self, sample):
        if "proj" not in sample:
            sample["proj"] = self.projector(self.clouds_of(sample), sample.get("bev_geom"))
        return self.inner(sample)

    def loss(self, *a, **k):
        return self.inner.loss(*a, **k)


def register(PCENCODER, build_from_cfg):
    """Register ``OnTheFlyPostProjector`` in the reference's PCENCODER registry.  Config:
    ``pcencoder = dict(type='OnTheFlyPostProjector', inner=dict(type='PostProjector2', ...))``."""

    class OnTheFlyPostProjector(OnTheFlyProjector):
        def __init__(self, inner, cfg=None, **proj_kwargs):
            module = build_from_cfg(inner, PCENCODER, default_args=dict(cfg=cfg)) if isinstance(inner, dict) else inner
            super().__init__(module, cfg=cfg, **proj_kwargs)

    return PCENCODER.register_module(OnTheFlyPostProjector)
