"""Same-named, same-signature GPU versions of the reference functions on either side of the rasteriser,
so that a maintainer can swap an import and the reference's scripts run unchanged:

    from lanemapping_b200.ref_api import write_instance_orientation_seq        # data/convert_data.py:319
    from lanemapping_b200.ref_api import transform_coordinate_from_img_2_pc    # baseline/utils/coor_img2pc.py:127

Arguments, return values, file formats and error behaviour follow the reference; the per-pixel /
per-vertex loops run in liblm_bev.so (include/lm_post.h).  There is no CPU fallback.
"""
from __future__ import annotations

import json
from typing import Sequence

import numpy as np
import torch

from . import post


class NpEncoder(json.JSONEncoder):
    """reference data/convert_data.py:14-22"""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def save_seq(seqs, seq_lens, seqs_semantic, seqs_instance, seqs_orient, seqs_filename) -> None:
    """reference data/convert_data.py:54-70 (same keys, same order, same encoder)."""
    lines_labeled = []
    for i, seq_len in enumerate(seq_lens):
        lines_labeled.append({"semantic": seqs_semantic[i], "instance": seqs_instance[i], "seq_len": seq_len,
                              "seq": seqs[i, :seq_len, :], "init_vertex": seqs[i, 0, :],
                              "end_vertex": seqs[i, seq_len - 1, :], "seq_orient": seqs_orient[i, :seq_len]})
    with open(seqs_filename, "w") as f:
        json.dump(lines_labeled, f, cls=NpEncoder)


def label_images(new_seqs, new_seq_lens, new_seqs_semantic, new_seqs_instance, new_seqs_orient,
                 device: str | torch.device = "cuda", size: int = 1152):
    """The four uint8 [1152,1152] label images of reference data/convert_data.py:319-361 as numpy arrays
    (semantic, instance, orient, endp)."""
    dev = torch.device(device)
    seqs = np.asarray(new_seqs, dtype=np.float64)
    if seqs.ndim != 3 or seqs.shape[2] != 2:
        raise ValueError("new_seqs must be [n_line, n_vertex, 2] (row, col)")
    L = seqs.shape[0]
    i32 = lambda a, shape: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int64).reshape(shape)
                                                               .astype(np.int32))).to(dev)
    out = post.label_rasters(torch.from_numpy(np.ascontiguousarray(seqs)).to(dev), i32(new_seq_lens, (L,)),
                             i32(new_seqs_semantic, (L,)), i32(new_seqs_instance, (L,)),
                             i32(np.asarray(new_seqs_orient)[:, :seqs.shape[1]], (L, seqs.shape[1])), size, size)
    return tuple(out[k].cpu().numpy() for k in ("semantic", "instance", "orient", "endp"))


def write_instance_orientation_seq(new_seqs, new_seq_lens, new_seqs_semantic, new_seqs_instance, new_seqs_orient,
                                   seqs_filename, semantic_filename, instance_filename, orient_filename, endp_filename,
                                   device: str | torch.device = "cuda") -> None:
    """reference data/convert_data.py:319-369: writes the four label PNGs and the sequence JSON."""
    import cv2
    new_seqs = np.asarray(new_seqs)
    if new_seqs.shape[0] > 0:
        sem, ins, ori, endp = label_images(new_seqs, new_seq_lens, new_seqs_semantic, new_seqs_instance,
                                           new_seqs_orient, device)
    else:                                   # upstream: "to avoid a NULL matrix in endpoints map generation" (:358)
        sem = ins = ori = endp = np.zeros((1152, 1152), dtype=np.uint8)
    img_quality = [cv2.IMWRITE_PNG_COMPRESSION, 9]      # upstream passes 100, which OpenCV clamps to 9
    cv2.imwrite(semantic_filename, sem, img_quality)
    cv2.imwrite(instance_filename, ins, img_quality)
    cv2.imwrite(orient_filename, ori, img_quality)
    cv2.imwrite(endp_filename, endp, img_quality)
    save_seq(new_seqs, new_seq_lens, new_seqs_semantic, new_seqs_instance, np.asarray(new_seqs_orient), seqs_filename)


def transform_coordinate_from_img_2_pc(params: dict, img_seqs, img_seq_lens: Sequence[int], bev_img,
                                       device: str | torch.device = "cuda") -> np.ndarray:
    """reference baseline/utils/coor_img2pc.py:127-183: polylines in BEV pixels -> LAS world coordinates
    [n_line, max_line_len, 3] float64.  ``bev_img`` is a PIL image or an array, as upstream."""
    dev = torch.device(device)
    img = np.array(bev_img)
    if img.ndim != 3 or img.dtype != np.uint8:
        raise ValueError("bev_img must be an 8-bit image with channels (the cropped_tiff PNG)")
    seqs = np.ascontiguousarray(np.asarray(img_seqs, dtype=np.float64))
    n_line = seqs.shape[0]
    lens = np.asarray(list(img_seq_lens), dtype=np.int32).reshape(1, n_line)
    world = post.img2pc(torch.from_numpy(np.ascontiguousarray(img[None])).to(dev), torch.from_numpy(seqs[None]).to(dev),
                        torch.from_numpy(lens).to(dev), [params], fill_in_place=True)
    return world[0].cpu().numpy()
