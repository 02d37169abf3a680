"""CPU restatement of the reference's INVERSE map (BEV pixel -> LAS world).  TEST INFRASTRUCTURE ONLY.

Follows reference baseline/utils/coor_img2pc.py:
  * modify_empty_pixel_elevation, roi branch            :97-122
  * transform_coordinate_from_img_2_pc                  :127-183
  * LeastSuqare (sic)                                   :59-73
  * rotateByQuanternion3D / multiplyQuanternion         :22-53
Pinned by tests/golden/inverse_*.json, which were produced by importing and running the
reference's own functions in the build container (tests/golden/make_golden.py).

This is the only in-tree reference code that fixes the projection geometry, so the forward
rasteriser's spec is *defined* as the map this function inverts (row <-> x, col <-> y,
channel 1 = elevation, all-zero pixel = empty).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-6


def fill_empty_elevation(img, roi_pts, roi_len):
    """coor_img2pc.py:97-122: for every polyline vertex on an (almost) empty pixel, take the mean
    G value over the smallest window [pt-step, pt+step) that contains a non-empty pixel."""
    img = img.copy()
    H, W, _ = img.shape
    for l in range(roi_pts.shape[0]):
        for k in range(int(roi_len[l])):
            ph, pw = int(roi_pts[l, k, 0]), int(roi_pts[l, k, 1])
            if (ph == 0 and pw == 0) or int(img[ph, pw, :].astype(np.int64).sum()) > 1:
                continue
            step = 1
            while True:
                win = img[max(ph - step, 0):min(ph + step, H), max(pw - step, 0):min(pw + step, W), :]
                if int(win.astype(np.int64).sum()) > 0:
                    valid = int((win.astype(np.int64).sum(axis=2) > 0).sum())
                    img[ph, pw, 1] = win[:, :, 1].astype(np.int64).sum() / valid   # uint8 store truncates
                    break
                step += 1
    return img


def least_square(X, Y):
    """coor_img2pc.py:59-73.  Upstream uses Python's builtin sum(): strictly left-to-right float64
    additions -- np.sum's pairwise order rounds differently from 8 elements on, so it must not be used
    here (tests/golden/inverse_io3.npz, random data run through the reference, pins this)."""
    n = len(Y)
    p = n * sum(X * Y) - sum(X) * sum(Y)
    q = n * sum(X * X) - sum(X) * sum(X)
    w = 0.0 if abs(q) < EPS else p / q
    b = sum(Y - w * X) / n
    return w, b


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def rotate(q, v):
    """coor_img2pc.py:38-53 (note: the inverse is divided by |q|, not |q|^2, as upstream)."""
    n = np.sqrt(np.sum(np.square(q)))
    qi = np.array(q, dtype=np.float64)
    qi[1:] *= -1.0
    qi /= n
    return quat_mul(quat_mul(q, np.array([0.0, v[0], v[1], v[2]])), qi)[1:]


def img2pc(params, img_seqs, img_seq_lens, bev_img):
    """coor_img2pc.py:127-183.  params: dict as returned by load_pc_2_img_transform_paras."""
    n_line, max_len, _ = img_seqs.shape
    out = np.zeros((n_line, max_len, 3))
    out[:, :, 0] = img_seqs[:, :, 0] * params["img_reso"][0] + params["bev_img_offset"][0]     # :136-139
    out[:, :, 1] = img_seqs[:, :, 1] * params["img_reso"][1] + params["bev_img_offset"][1]
    img = fill_empty_elevation(np.array(bev_img), img_seqs, img_seq_lens)                     # :145
    out[:, :, 2] = img[img_seqs[:, :, 0].astype(int), img_seqs[:, :, 1].astype(int), 1] * params["ele_reso"] \
        + params["local_min_ele"]                                                               # :150
    for l in range(n_line):                                                                     # :154-159
        k = int(img_seq_lens[l])
        idx = np.arange(k)
        w, b = least_square(idx, np.array(out[l, :k, 2]))
        out[l, :k, 2] = w * idx + b
    t = np.array(params["las_rotation_trans_quan"][0:3])                                        # :163-172
    q = np.array(params["las_rotation_trans_quan"][3:])
    for l in range(n_line):
        for v in range(max_len):
            out[l, v, :] = rotate(q, out[l, v, :]) + t
    return out + np.array(params["las_read_offset"])                                            # :175-177
