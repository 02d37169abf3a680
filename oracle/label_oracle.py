"""CPU restatement of the reference's label rasterisation and its pixel -> world post-processing
inputs.  TEST INFRASTRUCTURE ONLY (same import rule as bev_oracle.py).

Follows reference data/convert_data.py:
  * gaussian                              :248-254   (math.exp per pixel, cast to float32)
  * get_endpoint_maps_per_batch           :255-317   (merge_endp_map=True branch)
  * write_instance_orientation_seq        :319-369   (cv2.line x3, endpoint map * 255, cv2.imwrite)
PARITY PINNED: tests/golden/labels_*.png were written by running the reference's own
write_instance_orientation_seq in the build container (tests/golden/make_golden.py) on the
polylines in tests/golden/labels_in.json; tests/test_label_oracle.py checks this file against them.
"""
from __future__ import annotations

import math

import numpy as np

CLIP = 20          # clip_width = kernel_size * 5 (:263-264)
SIGMA = 2.0        # kernel_size / 2


def endpoint_map(starts, ends, H=1152, W=1152) -> np.ndarray:
    """uint8 [H, W]: what cv2.imwrite stores for ``label_endp_map * 255`` (:357-361, :368).
    The float64 map is converted with cv2's saturate_cast<uchar> (round half to even)."""
    starts = np.asarray(starts, dtype=np.float64).reshape(-1, 2)
    ends = np.asarray(ends, dtype=np.float64).reshape(-1, 2)
    heat = np.zeros((H, W), dtype=np.float64)
    rr, cc = np.arange(H)[:, None], np.arange(W)[None, :]
    for a, b in zip(starts, ends):
        if abs(b[0] - a[0]) < 1e-3 and abs(b[1] - a[1]) < 1e-3:          # :267-270
            continue
        for p in (a, b):
            if p[0] > CLIP and p[0] < (H - CLIP) and p[1] > CLIP and p[1] < (W - CLIP):   # :274-275
                d2 = (rr - int(p[0])) ** 2 + (cc - int(p[1])) ** 2
                vals = {int(k): np.float32(math.exp(-int(k) / (2 * SIGMA ** 2))) for k in np.unique(d2)}
                lut = np.zeros(int(d2.max()) + 1, dtype=np.float32)
                for k, v in vals.items():
                    lut[k] = v
                heat = np.maximum(heat, lut[d2].astype(np.float64))       # np.max of the two maps, np.amax over lanes
                heat[int(p[0]), int(p[1])] = 1.0                          # :293-298
    return np.clip(np.rint(heat * 255.0), 0, 255).astype(np.uint8)


def polyline_labels(seqs, lens, semantic, instance, orient, H=1152, W=1152):
    """(semantic, instance, orient) uint8 [H, W] rasters, drawn like :326-356."""
    import cv2
    sem = np.zeros((H, W), dtype=np.uint8)
    ins = np.zeros((H, W), dtype=np.uint8)
    ori = np.zeros((H, W), dtype=np.uint8)
    for i, n in enumerate(lens):
        s = 128 if semantic[i] == 1 else 255
        t = int(instance[i])
        for v in range(int(n) - 1):
            p0 = tuple(map(int, seqs[i, v, ::-1]))
            p1 = tuple(map(int, seqs[i, v + 1, ::-1]))
            cv2.line(sem, p0, p1, s)
            cv2.line(ins, p0, p1, t)
            cv2.line(ori, p0, p1, int(orient[i, v]))
    return sem, ins, ori


def line_pixels(x1, y1, x2, y2):
    """The pixel sequence of cv::LineIterator(pt1, pt2, 8, leftToRight=true) for end points inside the
    image (no clipping): what cv2.line draws with thickness 1.  Pure Python; pinned against cv2 itself
    in tests/test_label_oracle.py and used as the model of the CUDA kernel."""
    dx, dy, sx, sy = x2 - x1, y2 - y1, 1, 1
    if dx < 0:
        dx, dy, x1, y1 = -dx, -dy, x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy, sx, sy = dy, dx, sy, sx
    err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
    out, x, y = [], x1, y1
    for _ in range(dx + 1):
        out.append((x, y))
        m = err < 0
        err += minus + (plus if m else 0)
        if vert:
            y += sx
            x += sy if m else 0
        else:
            x += sx
            y += sy if m else 0
    return out
