"""Torch-facing API of the stages around the rasteriser (include/lm_post.h): BEV pixel polylines ->
LAS world coordinates, and the label rasters.  The reference runs both as per-vertex / per-pixel
Python loops (baseline/utils/coor_img2pc.py:127-183, data/convert_data.py:248-369); here each is a
few kernel launches on a batch.  PyTorch owns the memory; the arithmetic is in liblm_bev.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _cabi


def _cuda(t: torch.Tensor, name: str) -> None:
    if t.device.type != "cuda":
        raise RuntimeError(f"{name}: lanemapping_b200 runs on CUDA devices only (no CPU fallback)")


def img2pc_params(params: dict) -> np.ndarray:
    """One crop's sidecar dict (as returned by the reference's ``load_pc_2_img_transform_paras``,
    baseline/utils/io_utils.py:125-150) -> the 21 float64 values of ``lm_img2pc_params``.  The inverse
    quaternion is formed exactly as ``rotateByQuanternion3D`` does (coor_img2pc.py:38-48)."""
    quan = np.array(params["las_rotation_trans_quan"][3:], dtype=np.float64)
    quan_norm = np.sqrt(np.sum(np.square(quan)))
    if not quan_norm > 1e-6:
        raise ValueError("img2pc: zero quaternion")
    quan_inv = quan.copy()
    quan_inv[1:4] *= -1.0
    quan_inv /= quan_norm
    return np.concatenate([np.asarray(params["img_reso"], np.float64), np.asarray(params["bev_img_offset"], np.float64),
                           [float(params["ele_reso"]), float(params["local_min_ele"])],
                           np.asarray(params["las_rotation_trans_quan"][:3], np.float64), quan, quan_inv,
                           np.asarray(params["las_read_offset"], np.float64)])


def img2pc(images: torch.Tensor, seqs: torch.Tensor, lens: torch.Tensor, params: Sequence[dict],
           fill_in_place: bool = False) -> torch.Tensor:
    """``transform_coordinate_from_img_2_pc`` for a batch of crops.

    images u8 [B,H,W,C] (cuda), seqs f64 [B,L,V,2] (row, col), lens i32 [B,L], params: B sidecar dicts.
    Returns world coordinates f64 [B,L,V,3].  The hole filling works on a copy of ``images`` unless
    ``fill_in_place`` (the reference also fills a working copy, coor_img2pc.py:141-145)."""
    _cuda(images, "img2pc")
    if images.dtype != torch.uint8 or images.dim() != 4:
        raise ValueError("img2pc: images must be uint8 [B,H,W,C]")
    B, H, W, Cn = images.shape
    if seqs.dtype != torch.float64 or seqs.dim() != 4 or seqs.shape[0] != B or seqs.shape[3] != 2:
        raise ValueError("img2pc: seqs must be float64 [B,L,V,2]")
    L, V = int(seqs.shape[1]), int(seqs.shape[2])
    if lens.dtype != torch.int32 or tuple(lens.shape) != (B, L) or len(params) != B:
        raise ValueError("img2pc: lens must be int32 [B,L] and params must have B entries")
    work = images.contiguous() if fill_in_place else images.clone(memory_format=torch.contiguous_format)
    seqs, lens = seqs.contiguous(), lens.contiguous()
    par = torch.from_numpy(np.stack([img2pc_params(p) for p in params])).to(images.device)
    assert par.shape[1] * 8 == C.sizeof(_cabi.LmImg2PcParams)
    world = torch.empty((B, L, V, 3), dtype=torch.float64, device=images.device)
    st = torch.cuda.current_stream(images.device)
    with torch.cuda.device(images.device):
        _cabi.check(_cabi.lib().lm_bev_img2pc(work.data_ptr(), B, H, W, Cn, seqs.data_ptr(), lens.data_ptr(), L, V,
                                              par.data_ptr(), world.data_ptr(), st.cuda_stream))
    return world


def label_rasters(seqs: torch.Tensor, lens: torch.Tensor, semantic: torch.Tensor, instance: torch.Tensor,
                  orient: torch.Tensor, height: int = 1152, width: int = 1152) -> Dict[str, torch.Tensor]:
    """The four label rasters of one crop (reference data/convert_data.py:319-369).

    seqs f64 [L,V,2] (row, col), lens / semantic / instance i32 [L], orient i32 [L,V] (cuda).
    Returns u8 [H,W] tensors ``semantic``, ``instance``, ``orient``, ``endp``."""
    _cuda(seqs, "label_rasters")
    if seqs.dtype != torch.float64 or seqs.dim() != 3 or seqs.shape[2] != 2:
        raise ValueError("label_rasters: seqs must be float64 [L,V,2]")
    L, V = int(seqs.shape[0]), int(seqs.shape[1])
    for t, shape, name in ((lens, (L,), "lens"), (semantic, (L,), "semantic"), (instance, (L,), "instance"),
                           (orient, (L, V), "orient")):
        if t.dtype != torch.int32 or tuple(t.shape) != shape or t.device != seqs.device:
            raise ValueError(f"label_rasters: {name} must be int32 {shape} on the seqs' device")
    seqs, lens, semantic, instance, orient = (t.contiguous() for t in (seqs, lens, semantic, instance, orient))
    dev = seqs.device
    out = {k: torch.empty((height, width), dtype=torch.uint8, device=dev) for k in ("semantic", "instance", "orient", "endp")}
    scratch = torch.empty((height, width), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    lib = _cabi.lib()
    with torch.cuda.device(dev):
        _cabi.check(lib.lm_label_polylines(seqs.data_ptr(), lens.data_ptr(), semantic.data_ptr(), instance.data_ptr(),
                                           orient.data_ptr(), L, V, height, width, out["semantic"].data_ptr(),
                                           out["instance"].data_ptr(), out["orient"].data_ptr(), scratch.data_ptr(),
                                           st.cuda_stream))
        # first / last vertex of every lane (convert_data.py:336-337)
        if L > 0:
            idx = (lens.clamp(min=1).to(torch.int64) - 1).view(L, 1, 1).expand(L, 1, 2)
            starts = seqs[:, 0, :].contiguous()
            ends = torch.gather(seqs, 1, idx).view(L, 2).contiguous()
        else:
            starts = ends = torch.zeros((1, 2), dtype=torch.float64, device=dev)
        _cabi.check(lib.lm_label_endpoint_map(starts.data_ptr(), ends.data_ptr(), L, height, width,
                                              out["endp"].data_ptr(), st.cuda_stream))
    return out


def color_jitter_normalize_(proj: torch.Tensor, params, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """In place on proj f32 [B,3,H,W] (cuda): torchvision ColorJitter with the given per-sample draws,
    then Normalize(mean, std) -- the reference's ``img_transform``
    (baseline/datasets/laserlane_proposals.py:255-264) moved from the DataLoader worker to the GPU.

    ``params[b]`` = ``(fn_idx, brightness, contrast, saturation)`` as returned (first four values) by
    ``torchvision.transforms.ColorJitter.get_params``; a factor of None skips that operation."""
    _cuda(proj, "color_jitter_normalize_")
    if proj.dtype != torch.float32 or proj.dim() != 4 or proj.shape[1] != 3 or not proj.is_contiguous():
        raise ValueError("color_jitter_normalize_: proj must be contiguous float32 [B,3,H,W]")
    B, _, H, W = proj.shape
    if len(params) != B:
        raise ValueError("color_jitter_normalize_: one parameter tuple per sample")
    tab = (_cabi.LmJitter * max(B, 1))()
    for b, (fn_idx, bf, cf, sf) in enumerate(params):
        for k, v in enumerate(fn_idx):
            tab[b].order[k] = int(v)
        tab[b].brightness = -1.0 if bf is None else float(bf)
        tab[b].contrast = -1.0 if cf is None else float(cf)
        tab[b].saturation = -1.0 if sf is None else float(sf)
    scratch = torch.empty((max(B, 1), _cabi.JITTER_PARTIALS), dtype=torch.float64, device=proj.device)
    st = torch.cuda.current_stream(proj.device)
    with torch.cuda.device(proj.device):
        _cabi.check(_cabi.lib().lm_proj_color_jitter(proj.data_ptr(), B, H, W, tab, float(mean), float(std),
                                                     scratch.data_ptr(), st.cuda_stream))
    return proj


class GpuColorJitter:
    """Drop-in for the loader's ``img_transform``: draws the factors exactly like
    ``torchvision.transforms.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5)`` (same torch RNG
    consumption per sample, so a seeded run augments like the reference) and applies them on the GPU."""

    def __init__(self, brightness=0.5, contrast=0.5, saturation=0.5, mean=0.5, std=0.5):
        import torchvision
        self._cj = torchvision.transforms.ColorJitter(brightness=brightness, contrast=contrast, saturation=saturation)
        self.mean, self.std = mean, std

    def draw(self, batch: int):
        cj = self._cj
        out = []
        for _ in range(batch):
            fn_idx, b, c, s, _h = cj.get_params(cj.brightness, cj.contrast, cj.saturation, cj.hue)
            out.append((fn_idx.tolist(), b, c, s))
        return out

    def __call__(self, proj: torch.Tensor, params=None) -> torch.Tensor:
        return color_jitter_normalize_(proj, self.draw(proj.shape[0]) if params is None else params, self.mean, self.std)
