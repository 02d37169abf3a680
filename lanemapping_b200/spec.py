"""Frozen forward specification of the MLS point cloud -> BEV raster stage.

The upstream repo does not ship this stage (it points at an external tool,
reference README.md:171-172), so the forward map is obtained by inverting the
one piece of in-tree code that pins the geometry, the BEV-pixel -> LAS-world map
``transform_coordinate_from_img_2_pc`` (reference baseline/utils/coor_img2pc.py:127-183):

* image **row** <-> local x, image **col** <-> local y      (coor_img2pc.py:136-139)
* channel index 1 (G as PIL reads the PNG) is elevation,
  ``z = G * ele_reso + local_min_ele``                       (coor_img2pc.py:150)
* an all-zero pixel means "empty cell"                       (coor_img2pc.py:78,106)
* intensity is the LAS u16 value clipped to [800, 33000]     (baseline/datasets/laserlane_proposals.py:626-628)

Everything here is plain data; the arithmetic lives in ``csrc/`` (CUDA) and is
restated on the CPU in ``oracle/`` for the tests only.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Tuple

# u8 image channel ids (values are part of the C-ABI, see include/lm_bev.h)
CH_MAX_I = 0    # max quantised intensity in the cell
CH_MEAN_I = 1   # (sum_i + count//2) // count
CH_MIN_Z = 2    # min quantised height
CH_MAX_Z = 3    # max quantised height
CH_MEAN_Z = 4   # (sum_z + count//2) // count
CH_DENSITY = 5  # min(count, 255)
CHANNEL_NAMES = ("max_i", "mean_i", "min_z", "max_z", "mean_z", "density")

# raw accumulator planes (u32) used for strip-halo merges (order is C-ABI)
ACC_COUNT, ACC_SUM_I, ACC_SUM_Z, ACC_MAX_I, ACC_MIN_Z, ACC_MAX_Z = range(6)
ACC_PLANES = 6
ACC_NAMES = ("count", "sum_i", "sum_z", "max_i", "min_z", "max_z")
MIN_Z_EMPTY = 0xFFFFFFFF  # value of the min_z plane where count == 0

TILE = 1152  # crop edge in px (reference configs/*:38, data/convert_data.py:322-324)


@dataclass(frozen=True)
class BevSpec:
    """Geometry + channel list of one raster (one mosaic, one strip or one crop).

    Field names follow the sidecar keys of ``cropped_tiff_param/<stem>.txt``
    (reference baseline/utils/io_utils.py:125-150).
    """
    height: int                      # rows   (local x axis)
    width: int                       # cols   (local y axis)
    bev_img_offset: Tuple[float, float] = (0.0, 0.0)
    img_reso: Tuple[float, float] = (0.05, 0.05)
    local_min_ele: float = 0.0
    ele_reso: float = 0.05
    inten_min: int = 800
    inten_max: int = 33000
    channels: Tuple[int, ...] = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY)
    count16: bool = False            # also emit a u16 plane min(count, 65535)
    # integer window into the global grid: a point's GLOBAL cell is computed from
    # bev_img_offset/img_reso, then (row0, col0) is subtracted.  Strips and crops are
    # therefore bit-exact sub-windows of the one-piece raster (no shifted float origin).
    row0: int = 0
    col0: int = 0

    def __post_init__(self):
        if self.height <= 0 or self.width <= 0:
            raise ValueError("BevSpec: height/width must be positive")
        if not (1 <= len(self.channels) <= 4):
            raise ValueError("BevSpec: 1..4 u8 channels")
        if any(c < 0 or c > CH_DENSITY for c in self.channels):
            raise ValueError("BevSpec: unknown channel id")
        if self.img_reso[0] <= 0 or self.img_reso[1] <= 0 or self.ele_reso <= 0:
            raise ValueError("BevSpec: resolutions must be positive")
        if not (0 <= self.inten_min < self.inten_max <= 65535):
            raise ValueError("BevSpec: need 0 <= inten_min < inten_max <= 65535")

    @property
    def n_channels(self) -> int:
        return len(self.channels)

    @property
    def cells(self) -> int:
        return self.height * self.width

    def out_bytes_per_cell(self) -> int:
        return self.n_channels + (2 if self.count16 else 0)

    def algorithmic_bytes(self, n_points: int) -> int:
        """every point read once as a packed float4, every
        output cell written once; no credit for intermediate traffic."""
        return 16 * int(n_points) + self.cells * self.out_bytes_per_cell()

    def window(self, row0: int, row1: int, col0: int = 0, col1: int | None = None) -> "BevSpec":
        """Spec of the window [row0,row1) x [col0,col1) of this raster, in this raster's
        own (local) indices.  Same float origin, shifted integer window."""
        if col1 is None:
            col1 = self.width
        if not (0 <= row0 < row1 and 0 <= col0 < col1):
            raise ValueError("BevSpec.window: empty or negative window")
        return replace(self, height=row1 - row0, width=col1 - col0,
                       row0=self.row0 + row0, col0=self.col0 + col0)
