"""DATASETS plug-in: feed per-crop point records instead of pre-baked PNGs.

The reference's loaders read ``cropped_tiff/<stem>.png`` in forked DataLoader workers
(reference baseline/datasets/laserlane_proposals.py:73-98; ``workers=12, pin_memory=True``,
reference baseline/datasets/registry.py:54-59).  CUDA cannot be used there, so the on-the-fly
path splits the work: this dataset only *loads* each crop's packed point records
(``<data_root>/crop_points/<stem>.npz``: float32 [N,4] + raster geometry) in the worker, and the
rasterisation runs in the main process on the GPU inside the PCENCODER wrapper
(lanemapping_b200/pcencoder.py), which fills ``sample['proj']``.

Plug-in surface (reference baseline/datasets/registry.py:15-25, utils/registry.py:54-80):
``@DATASETS.register_module class X(Dataset): __init__(self, data_root, data_split_file, mode, cfg=None)``.
``make_onthefly_dataset(base)`` derives that class from the reference's ``LaserLaneProposal`` when
it is importable (labels then come from its ``format_gt_column_proposal``); standalone it yields
points + names only (inference).
"""
from __future__ import annotations

import json
import os
import os.path as osp
from typing import Dict, List, Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from .sidecar import read_sidecar

SPLIT_KEYS = {"train": "train", "test": "test", "valid": "valid", "single": "single", "all": "pretrain",
              "infer_only": "pretrain"}      # reference laserlane_proposals.py:500-520


def read_split(data_root: str, data_split_file: str, mode: str) -> List[str]:
    with open(osp.join(data_root, data_split_file), "r") as jf:
        js = json.load(jf)
    if mode not in SPLIT_KEYS:
        raise AssertionError(f"mode {mode!r} not in {sorted(SPLIT_KEYS)}")     # reference :39 asserts too
    stems = list(js[SPLIT_KEYS[mode]])
    if mode == "valid":
        stems = stems[:150]                                                     # reference :512
    return stems


class CropPoints(Dataset):
    """Standalone on-the-fly dataset: ``sample = {'image_name', 'points', 'bev_geom'}``."""

    def __init__(self, data_root: str, data_split_file: str, mode: str, cfg=None,
                 points_dir: str = "crop_points", param_dir: str = "cropped_tiff_param"):
        self.data_root, self.mode, self.cfg = data_root, mode, cfg
        self.points_path = osp.join(data_root, points_dir)
        self.param_path = osp.join(data_root, param_dir)
        self.image_stem_list = read_split(data_root, data_split_file, mode)

    def __len__(self):
        return len(self.image_stem_list)

    def load_points(self, idx: int) -> Dict[str, torch.Tensor]:
        stem = self.image_stem_list[idx]
        npz = osp.join(self.points_path, stem + ".npz")
        if osp.exists(npz):
            # written by convert_data.rasterize_single_file(crop_points_dir=...): mosaic origin + integer
            # window of the crop -> the on-the-fly raster is bit-identical to the crop's PNG
            with np.load(npz) as z:
                pts, geom = z["points"], torch.from_numpy(z["geom"].astype(np.float64))
        else:
            # plain [N,4] records + the crop's own sidecar (its origin is the shifted float origin)
            pts = np.load(osp.join(self.points_path, stem + ".npy"))
            p = read_sidecar(osp.join(self.param_path, stem + ".txt"))
            geom = torch.tensor([p.bev_img_offset[0], p.bev_img_offset[1], p.img_reso[0], p.img_reso[1],
                                 p.local_min_ele, p.ele_reso, 0.0, 0.0], dtype=torch.float64)
        if pts.ndim != 2 or pts.shape[1] != 4:
            raise ValueError(f"{stem}: expected [N,4] (x, y, z, intensity)")
        return {"points": torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)), "bev_geom": geom}

    def __getitem__(self, idx):
        sample = dict()
        sample["image_name"] = self.image_stem_list[idx][0:11]                 # reference :76
        sample.update(self.load_points(idx))
        return sample


def make_onthefly_dataset(base, points_dir: str = "crop_points", param_dir: str = "cropped_tiff_param"):
    """Class factory: subclass a reference dataset (e.g. LaserLaneProposal) so that it returns
    ``points`` instead of ``proj``; label tensors still come from the base class."""

    class LaserLaneProposalOnTheFly(base):          # noqa: D401 - name is what configs refer to
        def __init__(self, data_root, data_split_file, mode, cfg=None):
            super().__init__(data_root, data_split_file, mode, cfg=cfg)
            self._pts = CropPoints.__new__(CropPoints)
            self._pts.points_path = osp.join(data_root, points_dir)
            self._pts.param_path = osp.join(data_root, param_dir)
            self._pts.image_stem_list = self.image_stem_list

        def __getitem__(self, idx):
            sample = dict()
            sample["image_name"] = self.image_stem_list[idx][0:11]
            sample.update(self._pts.load_points(idx))
            if self.mode in {"train", "valid", "test", "single", "all"}:        # reference :79-81
                sample.update(self.format_gt_column_proposal(idx))
            return sample

    return LaserLaneProposalOnTheFly


class PointBatch:
    """The ragged clouds of one batch, packed: ``points`` float32 [sum N_i, 4] + ``offsets`` int64 [B + 1]
    (host).  It is deliberately NOT a list of tensors: the reference's ``Runner.to_cuda`` turns a list of
    tensors into ``torch.cat([t.unsqueeze(0) ...])`` (reference baseline/engine/runner.py:139-142), which
    raises for clouds of different lengths.  A non-list entry goes through ``batch[k].cuda(non_blocking=True)``
    (runner.py:149), which this class implements; ``pin_memory()`` is what ``DataLoader(pin_memory=True)``
    (reference baseline/datasets/registry.py:54-56) calls on custom batch types; ``nn.DataParallel`` on the
    reference's single visible device (``CUDA_VISIBLE_DEVICES='0'``, train_gpu_0.py:6-7) hands a non-tensor
    object to the replica unchanged.  For DataParallel over SEVERAL devices use ``collate_points_padded``."""

    def __init__(self, points: torch.Tensor, offsets):
        self.points = points
        self.offsets = [int(o) for o in offsets]
        if self.offsets[0] != 0 or self.offsets[-1] != int(points.shape[0]) or \
                any(a > b for a, b in zip(self.offsets, self.offsets[1:])):
            raise ValueError("PointBatch: offsets must rise from 0 to len(points)")

    @classmethod
    def from_list(cls, clouds: List[torch.Tensor]) -> "PointBatch":
        offs = [0]
        for c in clouds:
            if c.ndim != 2 or c.shape[1] != 4:
                raise ValueError("PointBatch: every cloud must be [N, 4] (x, y, z, intensity)")
            offs.append(offs[-1] + int(c.shape[0]))
        pts = torch.cat([c.to(torch.float32) for c in clouds], dim=0) if clouds else torch.empty((0, 4))
        return cls(pts.contiguous(), offs)

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i: int) -> torch.Tensor:
        return self.points[self.offsets[i]:self.offsets[i + 1]]

    def clouds(self) -> List[torch.Tensor]:
        return [self[i] for i in range(len(self))]

    @property
    def device(self):
        return self.points.device

    def cuda(self, device=None, non_blocking: bool = False) -> "PointBatch":
        return PointBatch(self.points.cuda(device, non_blocking=non_blocking), self.offsets)

    def to(self, *a, **k) -> "PointBatch":
        return PointBatch(self.points.to(*a, **k), self.offsets)

    def pin_memory(self) -> "PointBatch":
        return PointBatch(self.points.pin_memory(), self.offsets)


def collate_points(batch: List[dict]) -> dict:
    """Collate for variable-length clouds (``collate_fn=`` of the reference's DataLoader, where it
    anticipates ``pseudo_collate``: reference baseline/datasets/registry.py:58).  ``points`` becomes ONE
    ``PointBatch`` -- see there why it must not stay a list of tensors -- and everything else is
    default-collated, so the unmodified ``Runner.to_cuda`` moves the whole batch."""
    from torch.utils.data import default_collate
    pts = PointBatch.from_list([b["points"] for b in batch])
    rest = [{k: v for k, v in b.items() if k != "points"} for b in batch]
    out = default_collate(rest)
    out["points"] = pts
    return out


def collate_points_padded(batch: List[dict]) -> dict:
    """The same batch as dense tensors: ``points`` [B, N_max, 4], padded with NaN records (the rasteriser
    drops NaN points, spec step 1) + ``points_count`` [B].  Costs the padding in H2D bytes, but every entry
    splits along dim 0, which is what ``nn.DataParallel`` over several devices (reference
    baseline/engine/runner.py:103) does to a batch."""
    from torch.utils.data import default_collate
    clouds = [b["points"] for b in batch]
    n_max = max([int(c.shape[0]) for c in clouds] + [1])
    pts = torch.full((len(clouds), n_max, 4), float("nan"), dtype=torch.float32)
    for i, c in enumerate(clouds):
        pts[i, :c.shape[0]] = c
    rest = [{k: v for k, v in b.items() if k != "points"} for b in batch]
    out = default_collate(rest)
    out["points"] = pts
    out["points_count"] = torch.tensor([int(c.shape[0]) for c in clouds], dtype=torch.int64)
    return out


def register(DATASETS, base=None):
    """Register the on-the-fly dataset in the reference's DATASETS registry."""
    cls = make_onthefly_dataset(base) if base is not None else CropPoints
    if base is None:
        cls = type("LaserLaneProposalOnTheFly", (CropPoints,), {})
    return DATASETS.register_module(cls)
