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
  * 14-line sidecar (reference baseline/utils/io_utils.py:125-150), one per crop, with the crop's own
    ``bev_img_offset`` and ``local_min_ele``
  * stem ``%06d_%04d`` = 11 chars (reference baseline/datasets/laserlane_proposals.py:76)

How a file is processed: the cloud goes to the device (``.las`` point blocks as they lie on disk, decoded
there), every point gets its crop id from the same float32 keys as the rasteriser, the cloud is grouped by
crop (one device sort), every crop gets its own ``local_min_ele`` and the crops are rasterised in batches of
equally-shaped 1152^2 windows of ONE global grid (``lm_bev_rasterize_batch``) -- so a crop is bit-identical
to the corresponding window of a one-piece raster of the file.
"""
from __future__ import annotations

import json
import os
import threading
import time
from functools import partial
from multiprocessing.pool import ThreadPool
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import las as las_io
from .sidecar import PcImgParams, world_to_local, write_sidecar
from .spec import CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_Z, CH_MIN_Z, TILE, BevSpec

ELEVATION_CHANNELS = (CH_MIN_Z, CH_MAX_Z, CH_MEAN_Z)
DEFAULT_CHANNELS = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY)
IDENTITY_POSE = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)
MIN_POINTS = 5       # the reference's read_las bails out below 5 points (laserlane_proposals.py:632-635)
BATCH_CROPS = 16     # crops per lm_bev_rasterize_batch call
MIN_ELE_MODES = ("robust", "min", "file")


def data_read_offset(mins: Sequence[float]) -> Tuple[float, float, float]:
    """``las_read_offset`` chosen from the DATA, not from the LAS header's offset field: many writers leave that
    at 0 for UTM / Gauss-Krueger coordinates of 5e5..4e6 m, where float32 has an ulp of 0.03..0.5 m against
    0.05 m cells.  floor(min) keeps every local coordinate in [0, extent) -- exact integers in float64, and the
    inverse map adds the same value back (reference baseline/utils/coor_img2pc.py:175-177)."""
    return tuple(float(np.floor(v)) for v in mins)


def load_cloud(path: str):
    """-> (xyz_world float64 [N,3], intensity [N], las_read_offset from the data).  ``.las`` or ``.npy`` ([N,4])."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".las":
        xyz, inten, hdr = las_io.read_las(path)
        off = data_read_offset(xyz.min(axis=0)) if len(xyz) else (0.0, 0.0, 0.0)
        return xyz, inten, off
    if ext == ".npy":
        a = np.load(path)
        if a.ndim != 2 or a.shape[1] < 4:
            raise ValueError(f"{path}: expected [N,>=4] (x, y, z, intensity)")
        xyz = a[:, :3].astype(np.float64)
        off = data_read_offset(xyz.min(axis=0)) if len(a) else (0.0, 0.0, 0.0)
        return xyz, a[:, 3], off
    raise ValueError(f"{path}: unsupported cloud format {ext!r}")


def plan_from_extent(lo, hi, img_reso, ele_reso, tile, channels, count16) -> BevSpec:
    """Mosaic of whole tiles covering the local-frame extent [lo, hi] (x, y, z)."""
    off = (float(np.floor(lo[0])), float(np.floor(lo[1])))      # integers: exact in float32
    n_r = max(1, int(np.ceil((hi[0] - off[0]) / img_reso[0] / tile + 1e-9)))
    n_c = max(1, int(np.ceil((hi[1] - off[1]) / img_reso[1] / tile + 1e-9)))
    min_ele = float(np.floor(lo[2] * 10.0) / 10.0)
    return BevSpec(n_r * tile, n_c * tile, bev_img_offset=off, img_reso=tuple(img_reso), local_min_ele=min_ele,
                   ele_reso=float(ele_reso), channels=tuple(channels), count16=count16)


def seq_id_of(filename: str) -> int:
    """The 6-digit sequence id a file name suggests (its first six digits; reference stems look like
    ``181013_0190``).  Only a suggestion: ``assign_seq_ids`` makes the ids of one run unique."""
    digits = "".join(ch for ch in os.path.splitext(os.path.basename(filename))[0] if ch.isdigit())
    return int(digits[:6]) if digits else 0


def assign_seq_ids(las_filenames: Sequence[str]) -> Dict[str, int]:
    """One sequence id per input file, unique within the run.  ``181013_0130.las`` and ``181013_0131.las``
    both suggest 181013: the first (in sorted order) keeps it, the others take the next free id, so two files
    never share an output stem.  Deterministic for a given file list."""
    used, out = set(), {}
    for f in sorted(las_filenames):
        sid = seq_id_of(f) % 1_000_000
        while sid in used:
            sid = (sid + 1) % 1_000_000
        used.add(sid)
        out[f] = sid
    return out


def _stem(seq_id: int, crop_index: int) -> str:
    return "%06d_%04d" % (seq_id % 1_000_000, crop_index % 10_000)


def png_params() -> list:
    """zlib level 1 with the run-length strategy: BEV crops are mostly short runs of equal bytes (empty cells, flat
    elevation), and OpenCV drops its RLE default as soon as a level is named -- measured on a 10 M-point crop:
    229 ms / 2.89 MB at level 1 alone, 149 ms / 2.65 MB with the strategy restored (lossless either way)."""
    import cv2
    return [cv2.IMWRITE_PNG_COMPRESSION, 1, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_RLE]


def _write_png(path: str, img: np.ndarray) -> None:
    import cv2
    # cv2 writes BGR(A): reverse so that PIL reads index 0 = channels[0], index 1 = elevation, ...
    if img.ndim == 3 and img.shape[2] == 3:
        img = img[..., ::-1]
    elif img.ndim == 3 and img.shape[2] == 4:
        img = img[..., [2, 1, 0, 3]]
    if not cv2.imwrite(path, np.ascontiguousarray(img), png_params()):
        raise IOError(f"cv2.imwrite failed for {path}")


_encoders = None


def _encode_pool():
    """PNG encoding is the converter's slowest stage by far (~0.1 s per crop against microseconds of GPU time) and
    zlib releases the GIL: the crops of ONE file are encoded by a process-wide pool of host threads (one per core),
    shared by every file the driver has in flight."""
    global _encoders
    with _lock:
        if _encoders is None:
            from concurrent.futures import ThreadPoolExecutor
            _encoders = ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1), thread_name_prefix="lm-png")
        return _encoders


def shutdown_encoders() -> None:
    """Join the encoder threads (a later ``_encode_pool()`` starts new ones): for callers that fork afterwards."""
    global _encoders
    with _lock:
        pool, _encoders = _encoders, None
    if pool is not None:
        pool.shutdown(wait=True)


# ---- process-wide state of a conversion run -------------------------------------------------------------
_lock = threading.Lock()
_claimed: Dict[Tuple[str, str], str] = {}       # (output dir, stem) -> source file that wrote it in this process
_devices: Dict[str, dict] = {}                   # device -> {"lock", "raster", "key"}: ONE workspace per device


def _claim(out_dir: str, stem: str, source: str) -> None:
    """Two inputs must never write the same ``<stem>.png``: raise instead of overwriting."""
    key = (os.path.abspath(out_dir), stem)
    with _lock:
        owner = _claimed.setdefault(key, source)
    if owner != source:
        raise FileExistsError(f"output stem {stem} of {source} was already written for {owner}: give the files "
                              "distinct seq_id values (multiprocessing_las_files does)")


def _device_slot(dev) -> dict:
    with _lock:
        return _devices.setdefault(str(dev), {"lock": threading.Lock(), "raster": None, "key": None})


def new_timings() -> dict:
    return {"files": 0, "points": 0, "crops": 0, "read_s": 0.0, "gpu_s": 0.0, "png_s": 0.0, "other_s": 0.0}


def rasterize_single_file(las_filename: str, new_tiff_dir: str, new_param_dir: str, seq_id: Optional[int] = None,
                          img_reso: Tuple[float, float] = (0.05, 0.05), ele_reso: float = 0.05, tile: int = TILE,
                          channels: Sequence[int] = DEFAULT_CHANNELS, count16_dir: Optional[str] = None,
                          pose: Sequence[float] = IDENTITY_POSE, first_index: int = 1, device: str = "cuda",
                          skip_existing: bool = True, crop_points_dir: Optional[str] = None,
                          las_decode: str = "gpu", min_ele: str = "robust", timings: Optional[dict] = None) -> List[str]:
    """LAS/NPY cloud -> every non-empty ``tile`` x ``tile`` crop as PNG + sidecar.  Returns the stems.

    ``las_decode='gpu'`` (default): the point-data block of a ``.las`` file goes to the device as it
    lies on disk and is decoded there (``lm_las_decode``, include/lm_las.h) -- the host never scales a
    coordinate; ``'host'`` decodes with numpy first (what ``.npy`` inputs always do).
    ``min_ele``: where each crop's ``local_min_ele`` comes from.  The u8 elevation channel spans only
    255 * ele_reso (12.75 m at 0.05 m), so one value per FILE saturates whole crops on long or hilly
    runs.  ``'robust'`` (default): the crop's 0.1 % height quantile minus 0.5 m (a single low outlier does not
    shift the range); ``'min'``: the crop's true minimum; ``'file'``: one value for the file.  All floored to 0.1 m.
    The fraction of occupied cells that still saturate (elevation byte 255) is recorded in the manifest.
    ``timings``: a ``new_timings()`` dict that accumulates seconds spent reading / on the GPU / encoding PNGs."""
    import torch
    from .bev import BatchRasterizer

    t_begin = time.perf_counter()
    print("las filename", las_filename)
    source = os.path.abspath(las_filename)
    filetem = os.path.splitext(os.path.basename(las_filename))[0]
    if seq_id is None:
        seq_id = seq_id_of(las_filename)
    if min_ele not in MIN_ELE_MODES:
        raise ValueError(f"min_ele must be one of {MIN_ELE_MODES}")
    if las_decode not in ("gpu", "host"):
        raise ValueError("las_decode must be 'gpu' or 'host'")
    # the manifest is keyed by the INPUT file's own stem and remembers its source: another file that merely
    # shares the six digits can neither be mistaken for done nor silently overwrite these outputs
    manifest = os.path.join(new_param_dir, filetem + ".manifest.json")
    previous: List[str] = []
    if os.path.exists(manifest):
        with open(manifest) as f:
            man = json.load(f)
        if man.get("source") == source and man.get("seq_id") == seq_id % 1_000_000:
            previous = list(man["stems"])
            if skip_existing and all(os.path.exists(os.path.join(new_tiff_dir, s + ".png")) and
                                     os.path.exists(os.path.join(new_param_dir, s + ".txt")) for s in previous):
                for s in previous:
                    _claim(new_tiff_dir, s, source)
                return previous
    if len(channels) < 3:
        raise ValueError("cropped_tiff needs >= 3 channels (reference laserlane_proposals.py:93-94)")
    if channels[1] not in ELEVATION_CHANNELS:
        raise ValueError("channel index 1 must be an elevation channel (reference coor_img2pc.py:150)")
    ele_index = 1

    dev = torch.device(device)
    t0 = time.perf_counter()
    is_las_gpu = las_decode == "gpu" and os.path.splitext(las_filename)[1].lower() == ".las"
    if is_las_gpu:
        raw, hdr = las_io.read_point_block(las_filename)
        if hdr.n_points < MIN_POINTS:
            print("too few lidar pts: ", hdr.n_points, las_filename)
            return []
        # las_read_offset from the data: the header's min bounds when they are usable, else an exact integer
        # min-reduction of the X/Y/Z fields; never the header's offset field (see data_read_offset)
        read_off = data_read_offset(hdr.mins if hdr.bounds_ok() else las_io.world_min(raw, hdr))
        n_points = hdr.n_points
    else:
        xyz, inten, read_off = load_cloud(las_filename)
        if len(xyz) < MIN_POINTS:
            print("too few lidar pts: ", len(xyz), las_filename)
            return []
        n_points = len(xyz)
    params = PcImgParams(las_filename, tuple(read_off), tuple(pose), (0.0, 0.0), tuple(img_reso), 0.0, float(ele_reso))
    if not is_las_gpu:
        local = world_to_local(xyz, params)
        host_pts = np.empty((n_points, 4), dtype=np.float32)
        host_pts[:, :3] = local
        host_pts[:, 3] = inten
    t_read = time.perf_counter() - t0

    slot = _device_slot(dev)
    t0 = time.perf_counter()
    with slot["lock"], torch.cuda.device(dev):           # one file at a time per device, one workspace per device
        if is_las_gpu:
            from .bev import decode_las, las_xform
            pts_dev = decode_las(torch.from_numpy(raw).to(dev), n_points, las_xform(hdr, params))
        else:
            pts_dev = torch.from_numpy(host_pts).to(dev)
        ext = torch.stack([pts_dev[:, :3].amin(dim=0), pts_dev[:, :3].amax(dim=0)]).double().cpu().numpy()
        mosaic = plan_from_extent(ext[0], ext[1], img_reso, ele_reso, tile, channels, count16_dir is not None)
        n_r, n_c = mosaic.height // tile, mosaic.width // tile
        n_crops = n_r * n_c
        # crop id of every point from the rasteriser's own float32 keys (spec.py step 1): subtract, divide, floor
        f32 = torch.float32
        rr = torch.floor((pts_dev[:, 0] - torch.tensor(mosaic.bev_img_offset[0], dtype=f32, device=dev)) /
                         torch.tensor(img_reso[0], dtype=f32, device=dev))
        cc = torch.floor((pts_dev[:, 1] - torch.tensor(mosaic.bev_img_offset[1], dtype=f32, device=dev)) /
                         torch.tensor(img_reso[1], dtype=f32, device=dev))
        ok = (rr >= 0) & (rr < mosaic.height) & (cc >= 0) & (cc < mosaic.width)      # NaN compares false: dropped
        cid = torch.where(ok, torch.floor(rr / tile) * n_c + torch.floor(cc / tile),
                          torch.full_like(rr, float(n_crops))).to(torch.int32)
        del rr, cc, ok
        # group by crop; for the robust quantile the heights are ascending inside every crop (two stable sorts)
        if min_ele == "robust":
            zorder = torch.argsort(pts_dev[:, 2], stable=True)
            order = zorder[torch.argsort(cid[zorder], stable=True)]
            del zorder
        else:
            order = torch.argsort(cid, stable=True)
        counts = torch.bincount(cid, minlength=n_crops + 1).cpu().numpy()
        del cid
        pts_sorted = pts_dev[order].contiguous()
        del order, pts_dev
        starts = np.concatenate([[0], np.cumsum(counts)])
        live = [k for k in range(n_crops) if counts[k] > 0]
        file_min_ele = float(np.floor(ext[0][2] * 10.0) / 10.0)
        crop_min_ele = {}
        for k in live:
            z = pts_sorted[starts[k]:starts[k + 1], 2]
            if min_ele == "file":
                crop_min_ele[k] = file_min_ele
            elif min_ele == "min":
                crop_min_ele[k] = float(np.floor(float(z.min()) * 10.0) / 10.0)
            else:
                q = float(z[int(0.001 * (len(z) - 1))])
                crop_min_ele[k] = float(np.floor((q - 0.5) * 10.0) / 10.0)
        common = BevSpec(tile, tile, img_reso=tuple(img_reso), ele_reso=float(ele_reso), channels=tuple(channels),
                         count16=count16_dir is not None)
        outputs = ("image", "count16") if common.count16 else ("image",)
        key = (common, outputs)
        r = slot["raster"]
        if r is None or slot["key"] != key or r.max_points_total < n_points:
            r = BatchRasterizer(common, BATCH_CROPS, n_points, device=dev, outputs=outputs)
            slot["raster"], slot["key"] = r, key
        images, counts16, crop_pts = {}, {}, {}
        for g0 in range(0, len(live), BATCH_CROPS):
            group = live[g0:g0 + BATCH_CROPS]
            clouds = [pts_sorted[starts[k]:starts[k + 1]] for k in group]
            # every crop is an integer window (row0, col0) of the ONE mosaic grid with its own local_min_ele
            specs = [BevSpec(tile, tile, bev_img_offset=mosaic.bev_img_offset, img_reso=tuple(img_reso),
                             local_min_ele=crop_min_ele[k], ele_reso=float(ele_reso), channels=tuple(channels),
                             count16=common.count16, row0=(k // n_c) * tile, col0=(k % n_c) * tile) for k in group]
            out = r(clouds, specs)
            img = out["image"].cpu().numpy()
            c16 = out["count16"].cpu().numpy() if common.count16 else None
            for b, k in enumerate(group):
                images[k] = img[b]
                if c16 is not None:
                    counts16[k] = c16[b]
                if crop_points_dir is not None:
                    crop_pts[k] = clouds[b].cpu().numpy()
        st = r.stats()
        if st["error"]:
            raise RuntimeError(f"liblm_bev device error {st['error']} while rasterising {las_filename}")
        del pts_sorted
    t_gpu = time.perf_counter() - t0

    t0 = time.perf_counter()
    stems, saturated = [], {}
    pool, pending = _encode_pool(), []          # file writes of this cloud's crops, encoded in parallel
    for k in live:
        img = images[k]
        if not img.any():
            continue                                   # empty crop: nothing to learn from, skip silently
        i, j = divmod(k, n_c)
        stem = _stem(seq_id, first_index + len(stems))
        _claim(new_tiff_dir, stem, source)
        png_path = os.path.join(new_tiff_dir, stem + ".png")
        if os.path.exists(png_path) and stem not in previous:
            raise FileExistsError(f"{png_path} exists and was not written for {source}: stem collision")
        off = (mosaic.bev_img_offset[0] + i * tile * img_reso[0], mosaic.bev_img_offset[1] + j * tile * img_reso[1])
        pending.append(pool.submit(_write_png, png_path, img))
        write_sidecar(os.path.join(new_param_dir, stem + ".txt"),
                      PcImgParams(params.coor_las_path, params.las_read_offset, params.las_rotation_trans_quan,
                                  off, tuple(img_reso), crop_min_ele[k], float(ele_reso)))
        if common.count16:
            import cv2
            pending.append(pool.submit(cv2.imwrite, os.path.join(count16_dir, stem + ".png"), counts16[k]))
        if crop_points_dir is not None:
            # packed point records of this crop for the on-the-fly dataset (lanemapping_b200/datasets.py),
            # with the MOSAIC origin + the crop's integer window so that re-rasterising is bit-identical
            geom = np.array([mosaic.bev_img_offset[0], mosaic.bev_img_offset[1], img_reso[0], img_reso[1],
                             crop_min_ele[k], float(ele_reso), i * tile, j * tile], dtype=np.float64)
            pending.append(pool.submit(np.savez, os.path.join(crop_points_dir, stem + ".npz"), points=crop_pts[k], geom=geom))
        occ = img.any(axis=2)
        saturated[stem] = round(float((img[..., ele_index][occ] == 255).mean()), 6) if occ.any() else 0.0
        stems.append(stem)
    for fut in pending:
        fut.result()                                   # re-raises an encoder's exception; the manifest is written last
    t_png = time.perf_counter() - t0
    with open(manifest, "w") as f:
        json.dump({"source": source, "seq_id": seq_id % 1_000_000, "stems": stems, "grid": [mosaic.height, mosaic.width],
                   "n_points": int(n_points), "las_read_offset": list(read_off), "min_ele": min_ele,
                   "elevation_saturated_fraction": saturated}, f)
    if timings is not None:
        with _lock:
            timings["files"] += 1
            timings["points"] += int(n_points)
            timings["crops"] += len(stems)
            timings["read_s"] += t_read
            timings["gpu_s"] += t_gpu
            timings["png_s"] += t_png
            timings["other_s"] += time.perf_counter() - t_begin - t_read - t_gpu - t_png
    return stems


def multiprocessing_las_files(las_filenames: Sequence[str], new_tiff_dir: str, new_param_dir: str,
                              num_process: int = 12, devices: Optional[Sequence[str]] = None,
                              stats: Optional[dict] = None, **opts) -> List[str]:
    """Driver in the reference's idiom (``Pool(12).imap_unordered(partial(f, ...), files)`` with a
    tqdm bar, reference data/convert_data.py:429-436).  Workers are threads, not forked
    processes: CUDA contexts do not survive ``fork``; LAS parsing, PNG encoding and the GPU
    calls all release the GIL.

    Files are dealt round-robin over ``devices`` (default: every visible CUDA device) -- the converter's natural
    parallelism is one file per GPU, no exchange -- and every device has ONE workspace that its files take turns
    on, while reading and PNG encoding of other files overlap with it.  Every file gets its own sequence id
    (``assign_seq_ids``).  ``stats`` (a dict) receives files/s, points/s and where the time went."""
    import torch
    import tqdm
    for d in (new_tiff_dir, new_param_dir, opts.get("count16_dir"), opts.get("crop_points_dir")):
        if d and not os.path.exists(d):
            os.makedirs(d)
    if devices is None:
        if "device" in opts:
            devices = [opts.pop("device")]
        else:
            devices = [f"cuda:{i}" for i in range(max(1, torch.cuda.device_count()))]
    opts.pop("device", None)
    files = sorted(las_filenames)
    ids = assign_seq_ids(files)
    if "seq_id" in opts:
        if len(files) > 1:
            raise ValueError("seq_id= names ONE file's id; the driver assigns ids for a list of files")
        ids[files[0]] = opts.pop("seq_id")
    timings = new_timings()
    jobs = [(f, ids[f], devices[k % len(devices)]) for k, f in enumerate(files)]

    def run(job):
        f, sid, dev = job
        return rasterize_single_file(f, new_tiff_dir, new_param_dir, seq_id=sid, device=dev, timings=timings, **opts)

    stems: List[str] = []
    t0 = time.perf_counter()
    with ThreadPool(processes=max(1, min(num_process, len(jobs) or 1))) as p:
        with tqdm.tqdm(total=len(jobs)) as pbar:
            for got in p.imap_unordered(run, jobs):
                stems.extend(got)
                pbar.update()
    wall = time.perf_counter() - t0
    if stats is not None:
        busy = timings["read_s"] + timings["gpu_s"] + timings["png_s"] + timings["other_s"]
        stats.update(timings)
        stats.update({"wall_s": wall, "devices": list(devices), "threads": num_process,
                      "files_per_s": timings["files"] / wall if wall > 0 else 0.0,
                      "mpoints_per_s": timings["points"] / wall / 1e6 if wall > 0 else 0.0,
                      "png_share": timings["png_s"] / busy if busy > 0 else 0.0,
                      "gpu_share": timings["gpu_s"] / busy if busy > 0 else 0.0})
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
    run_stats: dict = {}
    multiprocessing_las_files(all_las_files, os.path.join(parent_dir, "cropped_tiff"),
                              os.path.join(parent_dir, "cropped_tiff_param"), stats=run_stats)
    print(json.dumps(run_stats))
