"""CPU oracle for the BEV rasterisation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``lanemapping_b200``) never does and fails loudly without its CUDA library.

PARITY UNPINNED.  The reference repo contains no forward rasteriser (it defers to
the external, un-vendored MIXIAOXIN/Las2BEV, reference README.md:171-172), no tests
and no golden images.  This file is therefore a numpy *restatement of the frozen
spec* (lanemapping_b200/spec.py), constrained by the in-tree
contracts that do exist and that tests/ check it against:

  * inverse map        reference baseline/utils/coor_img2pc.py:127-183
      X = row*reso0 + off0, Y = col*reso1 + off1 (:136-139)  -> row <-> x, col <-> y
      Z = img[row,col,1]*ele_reso + local_min_ele (:150)      -> channel 1 = elevation
  * empty-cell rule    reference baseline/utils/coor_img2pc.py:78,106 (channel sum < 1)
  * intensity clip     reference baseline/datasets/laserlane_proposals.py:626-628
  * loader contract    reference baseline/datasets/laserlane_proposals.py:85-98
      (uint8, square, >=3 channels, to_tensor => f32 CHW = u8/255)
  * sidecar format     reference baseline/utils/io_utils.py:125-150

numpy idiom mirrors what a numpy rasteriser in the reference's style would do
(floor keys -> np.bincount / np.maximum.at / np.minimum.at / weighted bincount).
"""
from __future__ import annotations

import numpy as np

# channel / plane ids: restated (not imported) so the oracle stands alone
CH_MAX_I, CH_MEAN_I, CH_MIN_Z, CH_MAX_Z, CH_MEAN_Z, CH_DENSITY = range(6)
ACC_COUNT, ACC_SUM_I, ACC_SUM_Z, ACC_MAX_I, ACC_MIN_Z, ACC_MAX_Z = range(6)
MIN_Z_EMPTY = np.uint32(0xFFFFFFFF)

f32 = np.float32


def quantise_points(pts, spec):
    """Per-point integer keys.  pts: [N,4] float32 (x, y, z, intensity-as-u16-value).

    Returns (row, col, iq, zq, valid): row/col are LOCAL window indices (int32),
    iq/zq are int32 in [0,255], valid is the in-window mask.  Every float operation
    is a single IEEE binary32 op (sub, div, floor, rint), as in the CUDA kernel.
    """
    pts = np.ascontiguousarray(pts, dtype=f32).reshape(-1, 4)
    x, y, z, inten = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3]
    off0, off1 = f32(spec.bev_img_offset[0]), f32(spec.bev_img_offset[1])
    r0, r1 = f32(spec.img_reso[0]), f32(spec.img_reso[1])
    with np.errstate(invalid="ignore", over="ignore"):
        # inverse of coor_img2pc.py:136-139
        rf = np.floor((x - off0) / r0)
        cf = np.floor((y - off1) / r1)
        lo_r, hi_r = f32(spec.row0), f32(spec.row0 + spec.height)
        lo_c, hi_c = f32(spec.col0), f32(spec.col0 + spec.width)
        valid = (rf >= lo_r) & (rf < hi_r) & (cf >= lo_c) & (cf < hi_c)   # NaN -> dropped
        row = np.where(valid, rf, lo_r).astype(np.int32) - np.int32(spec.row0)
        col = np.where(valid, cf, lo_c).astype(np.int32) - np.int32(spec.col0)
        # inverse of coor_img2pc.py:150 (round half to even), NaN -> 0 via fmax
        zf = np.rint((z - f32(spec.local_min_ele)) / f32(spec.ele_reso))
        zq = np.fmin(np.fmax(zf, f32(0)), f32(255)).astype(np.int32)
        # laserlane_proposals.py:626-628 clip, then our u8 mapping (truncating integer division)
        ic = np.fmin(np.fmax(inten, f32(spec.inten_min)), f32(spec.inten_max))
        iv = ic.astype(np.int32) - np.int32(spec.inten_min)
    iq = (iv.astype(np.int64) * 255 // (spec.inten_max - spec.inten_min)).astype(np.int32)
    return row, col, iq, zq, valid


def accumulate(pts, spec):
    """Integer accumulator planes, uint32 [6, H, W] in ACC_* order."""
    H, W = spec.height, spec.width
    row, col, iq, zq, valid = quantise_points(pts, spec)
    key = (row[valid].astype(np.int64) * W + col[valid])
    iq = iq[valid]
    zq = zq[valid]
    n = H * W
    acc = np.zeros((6, n), dtype=np.uint32)
    acc[ACC_COUNT] = np.bincount(key, minlength=n).astype(np.uint32)
    acc[ACC_SUM_I] = np.bincount(key, weights=iq, minlength=n).astype(np.uint32)   # exact: sums < 2**53
    acc[ACC_SUM_Z] = np.bincount(key, weights=zq, minlength=n).astype(np.uint32)
    mx = np.zeros(n, dtype=np.int32)
    np.maximum.at(mx, key, iq)
    acc[ACC_MAX_I] = mx
    mx = np.zeros(n, dtype=np.int32)
    np.maximum.at(mx, key, zq)
    acc[ACC_MAX_Z] = mx
    mn = np.full(n, 2**31 - 1, dtype=np.int64)
    np.minimum.at(mn, key, zq)
    acc[ACC_MIN_Z] = np.where(acc[ACC_COUNT] > 0, mn, int(MIN_Z_EMPTY)).astype(np.uint32)
    return acc.reshape(6, H, W)


def merge_acc(a, b):
    """Merge law of two accumulator sets over disjoint point sets (halo merge / strips)."""
    out = np.empty_like(a)
    for p in (ACC_COUNT, ACC_SUM_I, ACC_SUM_Z):
        out[p] = a[p] + b[p]
    for p in (ACC_MAX_I, ACC_MAX_Z):
        out[p] = np.maximum(a[p], b[p])
    out[ACC_MIN_Z] = np.minimum(a[ACC_MIN_Z], b[ACC_MIN_Z])
    return out


def channel_from_acc(acc, ch):
    cnt = acc[ACC_COUNT].astype(np.uint64)
    safe = np.maximum(cnt, 1)
    if ch == CH_MAX_I:
        v = acc[ACC_MAX_I]
    elif ch == CH_MEAN_I:
        v = (acc[ACC_SUM_I].astype(np.uint64) + cnt // 2) // safe
    elif ch == CH_MIN_Z:
        v = np.where(cnt > 0, acc[ACC_MIN_Z], 0)
    elif ch == CH_MAX_Z:
        v = acc[ACC_MAX_Z]
    elif ch == CH_MEAN_Z:
        v = (acc[ACC_SUM_Z].astype(np.uint64) + cnt // 2) // safe
    elif ch == CH_DENSITY:
        v = np.minimum(cnt, 255)
    else:
        raise ValueError(f"unknown channel {ch}")
    return np.asarray(v).astype(np.uint8)


def finalize(acc, spec):
    """acc [6,H,W] -> dict(image=u8 [H,W,C], count16=u16 [H,W] | None)."""
    img = np.stack([channel_from_acc(acc, c) for c in spec.channels], axis=-1)
    out = {"image": np.ascontiguousarray(img), "count16": None}
    if spec.count16:
        out["count16"] = np.minimum(acc[ACC_COUNT], 65535).astype(np.uint16)
    return out


def rasterize(pts, spec):
    """Points -> finished u8 HWC image (+ optional u16 count plane)."""
    return finalize(accumulate(pts, spec), spec)


def proj_from_image(image):
    """u8 [H,W,C] -> f32 [C,H,W] = u8/255, the loader's ``to_tensor(...).float()``
    (reference baseline/datasets/laserlane_proposals.py:88-89)."""
    return np.ascontiguousarray(image.transpose(2, 0, 1)).astype(f32) / f32(255)


def crop_tiles(image, tile=1152):
    """Non-overlapping tile x tile crops, row-major crop order; ragged edges zero-padded
    (an all-zero pixel is an empty cell, coor_img2pc.py:78)."""
    H, W = image.shape[:2]
    nr, nc = -(-H // tile), -(-W // tile)
    out = np.zeros((nr * nc, tile, tile) + image.shape[2:], dtype=image.dtype)
    for i in range(nr):
        for j in range(nc):
            blk = image[i * tile:(i + 1) * tile, j * tile:(j + 1) * tile]
            out[i * nc + j, :blk.shape[0], :blk.shape[1]] = blk
    return out


# ----------------------------------------------------------------------------------------
# parallel variant: the timed CPU baseline.  Mirrors the reference's own offline-script idiom,
# multiprocessing.Pool(P).imap_unordered over independent work items
# (reference data/convert_data.py:423-436); work items here are row strips.
# ----------------------------------------------------------------------------------------
_POOL_PTS = None


def _pool_init(pts):
    global _POOL_PTS
    _POOL_PTS = pts


def _pool_strip(args):
    spec, r0, r1, ranges = args
    sub = spec.window(r0, r1)
    if len(ranges) == 1:
        pts = _POOL_PTS[ranges[0][0]:ranges[0][1]]
    else:
        pts = np.concatenate([_POOL_PTS[lo:hi] for lo, hi in ranges], axis=0)
    return r0, r1, rasterize(pts, sub)


def scan_point_ranges(n, spec, processes, margin_m=2.5, roads=1):
    """Point index ranges per row strip for a scan-ordered cloud (synth.make_cloud, order='scan': the
    cloud is the concatenation of ``roads`` scans, in each of which the along-track position grows
    linearly with the index, +-1 m jitter): per road, the strip's own share of the indices widened by
    ``margin_m`` metres of track on both sides.  -> one list of (lo, hi) per strip."""
    H = spec.height
    P = max(1, int(processes))
    roads = max(1, int(roads))
    edges = [H * k // P for k in range(P + 1)]
    per_road = n / roads
    margin = int(margin_m / (H * spec.img_reso[0]) * per_road) + 1
    out = []
    for k in range(P):
        rs = []
        for r in range(roads):
            base = r * per_road
            lo = max(0, int(base + edges[k] / H * per_road) - margin)
            hi = min(n, int(base + edges[k + 1] / H * per_road) + margin + 2)
            if rs and lo <= rs[-1][1]:          # overlapping windows of neighbouring roads: one interval,
                rs[-1] = (rs[-1][0], max(hi, rs[-1][1]))     # so that no point is taken twice by one strip
            elif hi > lo:
                rs.append((lo, hi))
        out.append(rs)
    return out


def rasterize_pool(pts, spec, processes, point_ranges=None):
    """Row-strip parallel rasterise with ``processes`` forked workers.

    point_ranges: optional, per strip either one (lo, hi) point index range or a list of them (for
    along-track-ordered clouds, with a margin); default = every strip scans all points.
    """
    import multiprocessing as mp
    H = spec.height
    P = max(1, int(processes))
    edges = [H * k // P for k in range(P + 1)]
    jobs = []
    for k in range(P):
        if edges[k + 1] <= edges[k]:
            continue
        rs = [(0, len(pts))] if point_ranges is None else point_ranges[k]
        if rs and not isinstance(rs[0], (tuple, list)):
            rs = [tuple(rs)]
        jobs.append((spec, edges[k], edges[k + 1], [(int(a), int(b)) for a, b in rs]))
    img = np.zeros((H, spec.width, len(spec.channels)), dtype=np.uint8)
    c16 = np.zeros((H, spec.width), dtype=np.uint16) if spec.count16 else None
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=P, initializer=_pool_init, initargs=(pts,)) as pool:
        for r0, r1, out in pool.imap_unordered(_pool_strip, jobs):
            img[r0:r1] = out["image"]
            if c16 is not None:
                c16[r0:r1] = out["count16"]
    return {"image": img, "count16": c16}
