"""Offline converter: LAS clouds -> ``cropped_tiff/<stem>.png`` + ``cropped_tiff_param/<stem>.txt``.

Written beside (and in the style of) the reference's label converter ``data/convert_data.py``:
one per-file function (``process_single_file``, reference data/convert_data.py:371-396) mapped
over a pool by a driver (``multiprocessing_seqs_files``, :423-436), output stems derived from
the input stem (:378-384), 1152 x 1152 px tiles (:322-324).  The reference's own functions are
untouched; this module adds the raster stage the reference leaves to an external tool
(reference README.md:171-172).

Output contract (what the reference's loaders and inverse map read):
  * PNG, 8-bit, square 1152^2, >= 3 channels, RGB order *as PIL reads it*
    (reference baseline/datasets/laserlane_proposals.py:85-98); index 1 is an elevation channel
    (reference baseline/utils/coor_img2pc.py:150); empty cells are all-zero pixels (:78,106)
  * 14-line sidecar (reference baseline/utils/io_utils.py:125-150)
  * stem ``%06d_%04d`` = 11 chars (reference baseline/datasets/laserlane_proposals.py:76)
"""
from __future__ import annotations

import json
import os
import threading
from functools import partial
from multiprocessing.pool import ThreadPool
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import las as las_io
from .sidecar import PcImgParams, world_to_local, write_sidecar
from .spec import CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_Z, CH_MIN_Z, TILE, BevSpec

ELEVATION_CHANNELS = (CH_MIN_Z, CH_MAX_Z, CH_MEAN_Z)
DEFAULT_CHANNELS = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY)
IDENTITY_POSE = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)
MIN_POINTS = 5       # the reference's read_las bails out below 5 points (laserlane_proposals.py:632-635)


def load_cloud(path: str):
    """-> (xyz_world float64 [N,3], intensity [N], suggested las_read_offset).  ``.las`` or ``.npy`` ([N,4])."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".las":
        xyz, inten, hdr = las_io.read_las(path)
        return xyz, inten, tuple(float(v) for v in hdr.offset)
    if ext == ".npy":
        a = np.load(path)
        if a.ndim != 2 or a.shape[1] < 4:
            raise ValueError(f"{path}: expected [N,>=4] (x, y, z, intensity)")
        xyz = a[:, :3].astype(np.float64)
        off = tuple(float(v) for v in np.floor(xyz.min(axis=0))) if len(a) else (0.0, 0.0, 0.0)
        return xyz, a[:, 3], off
    raise ValueError(f"{path}: unsupported cloud format {ext!r}")


def plan_raster(xyz_world: np.ndarray, las_read_offset: Sequence[float], pose: Sequence[float],
                img_reso: Tuple[float, float], ele_reso: float, tile: int, channels: Sequence[int],
                count16: bool, coor_las_path: str):
    """Choose the local frame, the mosaic extent (whole tiles) and build spec + sidecar template."""
    params = PcImgParams(coor_las_path, tuple(las_read_offset), tuple(pose), (0.0, 0.0), tuple(img_reso), 0.0,
                         float(ele_reso))
    local = world_to_local(xyz_world, params)
    spec = plan_from_extent(local.min(axis=0), local.max(axis=0), img_reso, ele_reso, tile, channels, count16)
    return spec, local, params


def plan_from_extent(lo, hi, img_reso, ele_reso, tile, channels, count16) -> BevSpec:
    """Mosaic of whole tiles covering the local-frame extent [lo, hi] (x, y, z)."""
    off = (float(np.floor(lo[0])), float(np.floor(lo[1])))      # integers: exact in float32
    n_r = max(1, int(np.ceil((hi[0] - off[0]) / img_reso[0] / tile + 1e-9)))
    n_c = max(1, int(np.ceil((hi[1] - off[1]) / img_reso[1] / tile + 1e-9)))
    min_ele = float(np.floor(lo[2] * 10.0) / 10.0)
    return BevSpec(n_r * tile, n_c * tile, bev_img_offset=off, img_reso=tuple(img_reso), local_min_ele=min_ele,
                   ele_reso=float(ele_reso), channels=tuple(channels), count16=count16)


def _stem(seq_id: int, crop_index: int) -> str:
    return "%06d_%04d" % (seq_id % 1_000_000, crop_index % 10_000)


def _write_png(path: str, img: np.ndarray) -> None:
    import cv2
    # cv2 writes BGR(A): reverse so that PIL reads index 0 = channels[0], index 1 = elevation, ...
    if img.ndim == 3 and img.shape[2] == 3:
        img = img[..., ::-1]
    elif img.ndim == 3 and img.shape[2] == 4:
        img = img[..., [2, 1, 0, 3]]
    if not cv2.imwrite(path, np.ascontiguousarray(img), [cv2.IMWRITE_PNG_COMPRESSION, 1]):
        raise IOError(f"cv2.imwrite failed for {path}")


_tls = threading.local()


def rasterize_single_file(las_filename: str, new_tiff_dir: str, new_param_dir: str, seq_id: Optional[int] = None,
                          img_reso: Tuple[float, float] = (0.05, 0.05), ele_reso: float = 0.05, tile: int = TILE,
                          channels: Sequence[int] = DEFAULT_CHANNELS, count16_dir: Optional[str] = None,
                          pose: Sequence[float] = IDENTITY_POSE, first_index: int = 1, device: str = "cuda",
                          skip_existing: bool = True, crop_points_dir: Optional[str] = None,
                          las_decode: str = "gpu") -> List[str]:
    """LAS/NPY cloud -> every non-empty ``tile`` x ``tile`` crop as PNG + sidecar.  Returns the stems.

    ``las_decode='gpu'`` (default): the point-data block of a ``.las`` file goes to the device as it
    lies on disk and is decoded there (``lm_las_decode``, include/lm_las.h) -- the host never scales a
    coordinate; ``'host'`` decodes with numpy first (what ``.npy`` inputs always do)."""
    import torch
    from .bev import BevRasterizer, crop_tiles

    print("las filename", las_filename)
    filetem = os.path.splitext(os.path.basename(las_filename))[0]
    if seq_id is None:
        digits = "".join(ch for ch in filetem if ch.isdigit())
        seq_id = int(digits[:6]) if digits else 0
    manifest = os.path.join(new_param_dir, "%06d.manifest.json" % (seq_id % 1_000_000))
    if skip_existing and os.path.exists(manifest):
        with open(manifest) as f:
            stems = json.load(f)["stems"]
        if all(os.path.exists(os.path.join(new_tiff_dir, s + ".png")) and
               os.path.exists(os.path.join(new_param_dir, s + ".txt")) for s in stems):
            return stems
    if len(channels) < 3:
        raise ValueError("cropped_tiff needs >= 3 channels (reference laserlane_proposals.py:93-94)")
    if channels[1] not in ELEVATION_CHANNELS:
        raise ValueError("channel index 1 must be an elevation channel (reference coor_img2pc.py:150)")

    dev = torch.device(device)
    if las_decode not in ("gpu", "host"):
        raise ValueError("las_decode must be 'gpu' or 'host'")
    if las_decode == "gpu" and os.path.splitext(las_filename)[1].lower() == ".las":
        from .bev import decode_las, las_xform
        raw, hdr = las_io.read_point_block(las_filename)
        if hdr.n_points < MIN_POINTS:
            print("too few lidar pts: ", hdr.n_points, las_filename)
            return []
        read_off = tuple(float(v) for v in hdr.offset)
        params = PcImgParams(las_filename, read_off, tuple(pose), (0.0, 0.0), tuple(img_reso), 0.0, float(ele_reso))
        with torch.cuda.device(dev):
            pts_dev = decode_las(torch.from_numpy(raw).to(dev), hdr.n_points, las_xform(hdr, params))
            ext = torch.stack([pts_dev[:, :3].amin(dim=0), pts_dev[:, :3].amax(dim=0)]).double().cpu().numpy()
        spec = plan_from_extent(ext[0], ext[1], img_reso, ele_reso, tile, channels, count16_dir is not None)
        pts = pts_dev.cpu().numpy() if crop_points_dir is not None else None
        n_points = hdr.n_points
    else:
        xyz, inten, read_off = load_cloud(las_filename)
        if len(xyz) < MIN_POINTS:
            print("too few lidar pts: ", len(xyz), las_filename)
            return []
        spec, local, params = plan_raster(xyz, read_off, pose, img_reso, ele_reso, tile, channels,
                                          count16_dir is not None, las_filename)
        pts = np.empty((len(local), 4), dtype=np.float32)
        pts[:, :3] = local
        pts[:, 3] = inten
        pts_dev = torch.from_numpy(pts).to(dev, non_blocking=True)
        n_points = len(pts)

    key = (spec, dev)
    r = getattr(_tls, "raster", None)
    if r is None or getattr(_tls, "key", None) != key or r.max_points < n_points:
        outputs = ("image", "count16") if spec.count16 else ("image",)
        r = BevRasterizer(spec, n_points, device=dev, outputs=outputs)
        _tls.raster, _tls.key = r, key
    with torch.cuda.device(dev):
        out = r(pts_dev)
        crops = crop_tiles(out["image"], tile).cpu().numpy()
        crops16 = None
        if spec.count16:
            c16 = out["count16"].view(torch.uint8).reshape(spec.height, spec.width, 2)
            crops16 = crop_tiles(c16, tile).cpu().numpy().view(np.uint16)[..., 0]
        r.check_device_errors()

    n_c = spec.width // tile
    crop_of_point = None
    if crop_points_dir is not None:
        # which crop every point falls in: the same float32 keys as the rasteriser (spec.py step 1)
        f32 = np.float32
        rr = np.floor((pts[:, 0] - f32(spec.bev_img_offset[0])) / f32(img_reso[0]))
        cc = np.floor((pts[:, 1] - f32(spec.bev_img_offset[1])) / f32(img_reso[1]))
        ok = (rr >= 0) & (rr < spec.height) & (cc >= 0) & (cc < spec.width)
        crop_of_point = np.where(ok, (rr // tile) * n_c + (cc // tile), -1).astype(np.int64)
    stems = []
    for k in range(crops.shape[0]):
        if not crops[k].any():
            continue                                   # empty crop: nothing to learn from, skip silently
        i, j = divmod(k, n_c)
        stem = _stem(seq_id, first_index + len(stems))
        off = (spec.bev_img_offset[0] + i * tile * img_reso[0], spec.bev_img_offset[1] + j * tile * img_reso[1])
        _write_png(os.path.join(new_tiff_dir, stem + ".png"), crops[k])
        write_sidecar(os.path.join(new_param_dir, stem + ".txt"),
                      PcImgParams(params.coor_las_path, params.las_read_offset, params.las_rotation_trans_quan,
                                  off, tuple(img_reso), spec.local_min_ele, spec.ele_reso))
        if crops16 is not None:
            import cv2
            cv2.imwrite(os.path.join(count16_dir, stem + ".png"), crops16[k])
        if crop_of_point is not None:
            # packed point records of this crop for the on-the-fly dataset (lanemapping_b200/datasets.py),
            # with the MOSAIC origin + the crop's integer window so that re-rasterising is bit-identical
            geom = np.array([spec.bev_img_offset[0], spec.bev_img_offset[1], img_reso[0], img_reso[1],
                             spec.local_min_ele, spec.ele_reso, i * tile, j * tile], dtype=np.float64)
            np.savez(os.path.join(crop_points_dir, stem + ".npz"), points=pts[crop_of_point == k], geom=geom)
        stems.append(stem)
    with open(manifest, "w") as f:
        json.dump({"source": las_filename, "stems": stems, "grid": [spec.height, spec.width],
                   "n_points": int(n_points)}, f)
    return stems


def multiprocessing_las_files(las_filenames: Sequence[str], new_tiff_dir: str, new_param_dir: str,
                              num_process: int = 12, **opts) -> List[str]:
    """Driver in the reference's idiom (``Pool(12).imap_unordered(partial(f, ...), files)`` with a
    tqdm bar, reference data/convert_data.py:429-436).  Workers are threads, not forked
    processes: CUDA contexts do not survive ``fork``; LAS parsing, PNG encoding and the GPU
    calls all release the GIL."""
    import tqdm
    for d in (new_tiff_dir, new_param_dir, opts.get("count16_dir"), opts.get("crop_points_dir")):
        if d and not os.path.exists(d):
            os.makedirs(d)
    stems: List[str] = []
    with ThreadPool(processes=num_process) as p:
        with tqdm.tqdm(total=len(las_filenames)) as pbar:
            for got in p.imap_unordered(partial(rasterize_single_file, new_tiff_dir=new_tiff_dir,
                                                new_param_dir=new_param_dir, **opts), las_filenames):
                stems.extend(got)
                pbar.update()
    return sorted(stems)


if __name__ == "__main__":
    # same shape as the reference's __main__ (data/convert_data.py:440-478): walk a directory
    import sys
    las_dir = sys.argv[1] if len(sys.argv) > 1 else "./data/LaserLane/Test-Area/las"
    parent_dir, _ = os.path.split(las_dir.rstrip("/"))
    all_las_files = []
    for root, dirs, files in os.walk(las_dir):
        for filepath in sorted(files):
            abs_filepath = os.path.join(root, filepath)
            if os.stat(abs_filepath).st_size == 0:
                print("empty filepath: ", abs_filepath)
                continue
            if os.path.splitext(filepath)[1].lower() in (".las", ".npy"):
                all_las_files.append(abs_filepath)
    multiprocessing_las_files(all_las_files, os.path.join(parent_dir, "cropped_tiff"),
                              os.path.join(parent_dir, "cropped_tiff_param"))
