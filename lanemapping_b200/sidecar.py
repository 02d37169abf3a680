"""``cropped_tiff_param/<stem>.txt`` sidecars: the 14-line text format parsed by the
reference's ``load_pc_2_img_transform_paras`` (reference baseline/utils/io_utils.py:125-150).

Values sit on the odd 0-based lines 1,3,...,13, space-separated, in this order:
coor_las_path, las_read_offset (3), las_rotation_trans_quan (7 = t_xyz + q_wxyz, used at
reference baseline/utils/coor_img2pc.py:163-165), bev_img_offset (2), img_reso (2),
local_min_ele, ele_reso.  The even lines are free-text labels the parser skips.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np

LABELS = ("coor_las_path:", "las_read_offset:", "las_rotation_trans_quan:", "bev_img_offset:",
          "img_reso:", "local_min_ele:", "ele_reso:")


@dataclass(frozen=True)
class PcImgParams:
    coor_las_path: str
    las_read_offset: Tuple[float, float, float]
    las_rotation_trans_quan: Tuple[float, ...]      # (tx, ty, tz, qw, qx, qy, qz)
    bev_img_offset: Tuple[float, float]
    img_reso: Tuple[float, float]
    local_min_ele: float
    ele_reso: float


def _fmt(values: Sequence[float]) -> str:
    # repr keeps the float64 value exactly; single spaces only (the reference splits on ' ')
    return " ".join(repr(float(v)) for v in values)


def format_sidecar(p: PcImgParams) -> str:
    if len(p.las_read_offset) != 3 or len(p.las_rotation_trans_quan) != 7:
        raise ValueError("sidecar: las_read_offset needs 3 values, las_rotation_trans_quan 7")
    if "\n" in p.coor_las_path:
        raise ValueError("sidecar: newline in path")
    vals = (p.coor_las_path, _fmt(p.las_read_offset), _fmt(p.las_rotation_trans_quan), _fmt(p.bev_img_offset),
            _fmt(p.img_reso), repr(float(p.local_min_ele)), repr(float(p.ele_reso)))
    lines = []
    for label, v in zip(LABELS, vals):
        lines += [label, v]
    return "\n".join(lines) + "\n"


def write_sidecar(path: str, p: PcImgParams) -> None:
    with open(path, "w") as f:
        f.write(format_sidecar(p))


def read_sidecar(path: str) -> PcImgParams:
    """Same line positions as the reference parser (io_utils.py:136-149)."""
    with open(path) as f:
        lines = f.read().split("\n")
    if len(lines) < 14:
        raise ValueError(f"{path}: expected 14 lines, got {len(lines)}")
    fl = lambda s: tuple(float(x) for x in s.split(" "))
    return PcImgParams(lines[1], fl(lines[3]), fl(lines[5]), fl(lines[7]), fl(lines[9]), float(lines[11]),
                       float(lines[13]))


# ---------------------------------------------------------------------------------------------
# the rigid part of the map (float64, host): world <-> raster-local frame
# ---------------------------------------------------------------------------------------------
def quat_to_matrix(q_wxyz: Sequence[float]) -> np.ndarray:
    """Rotation matrix of v -> q v q^-1 (reference coor_img2pc.py:38-53) for a unit quaternion."""
    w, x, y, z = (float(v) for v in q_wxyz)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    if n < 1e-12:
        raise ValueError("zero quaternion")
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def world_to_local(xyz_world: np.ndarray, p: PcImgParams) -> np.ndarray:
    """Inverse of reference coor_img2pc.py:163-177:  p_world = R(q) p_local + t + las_read_offset."""
    t = np.asarray(p.las_rotation_trans_quan[:3], dtype=np.float64)
    R = quat_to_matrix(p.las_rotation_trans_quan[3:])
    d = np.asarray(xyz_world, dtype=np.float64) - np.asarray(p.las_read_offset, dtype=np.float64) - t
    return d @ R          # row-vector form of R^T d


def local_to_world(xyz_local: np.ndarray, p: PcImgParams) -> np.ndarray:
    t = np.asarray(p.las_rotation_trans_quan[:3], dtype=np.float64)
    R = quat_to_matrix(p.las_rotation_trans_quan[3:])
    return np.asarray(xyz_local, dtype=np.float64) @ R.T + t + np.asarray(p.las_read_offset, dtype=np.float64)
