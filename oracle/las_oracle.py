"""CPU oracle of the LAS-record decode (include/lm_las.h).  TEST INFRASTRUCTURE ONLY -- same import
rule as bev_oracle.py: tests/, __graft_entry__.smoke() and bench.py's CPU legs.

What it restates.  The reference reads clouds with ``laspy.read`` and stacks ``las.x, las.y, las.z``
(scaled float64) with ``las.intensity`` (reference baseline/datasets/laserlane_proposals.py:618-636);
laspy's scaled dimensions are ``X * scale + offset`` in float64 (ASPRS LAS 1.x: every point data
record format starts with X, Y, Z int32 and intensity u16, little endian).  The world -> raster-local
step inverts reference baseline/utils/coor_img2pc.py:163-177 (``p_world = R(q) p + t +
las_read_offset``).  laspy is not installed here and the reference has no test for this step:
PARITY UNPINNED upstream; pinned against lanemapping_b200/las.py's independent reader and a
hand-computed record in tests/test_las_oracle.py.

Every step is one IEEE binary64 operation in the order the header documents (numpy ufuncs never
fuse a multiply with an add), so the CUDA kernels reproduce it bit for bit.
"""
from __future__ import annotations

import numpy as np


def split_records(raw: np.ndarray, record_length: int):
    """raw: uint8 [n * record_length] -> (X, Y, Z int32 [n], intensity uint16 [n])."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if record_length < 14 or raw.size % record_length:
        raise ValueError("las_oracle: bad record length / truncated block")
    rec = raw.reshape(-1, record_length)
    ixyz = np.ascontiguousarray(rec[:, :12]).view("<i4").reshape(-1, 3)
    inten = np.ascontiguousarray(rec[:, 12:14]).view("<u2").reshape(-1)
    return ixyz[:, 0], ixyz[:, 1], ixyz[:, 2], inten


def decode_records(raw, record_length, scale, offset, las_read_offset=(0.0, 0.0, 0.0),
                   translation=(0.0, 0.0, 0.0), rot=(1, 0, 0, 0, 1, 0, 0, 0, 1)) -> np.ndarray:
    """-> float32 [n, 4] (x, y, z, intensity) in the raster-local frame."""
    X, Y, Z, inten = split_records(raw, record_length)
    sc, of = np.asarray(scale, np.float64), np.asarray(offset, np.float64)
    ro, t = np.asarray(las_read_offset, np.float64), np.asarray(translation, np.float64)
    m = np.asarray(rot, np.float64).reshape(9)
    world = [c.astype(np.float64) * sc[k] + of[k] for k, c in enumerate((X, Y, Z))]      # laspy's las.x/y/z
    d = [(world[k] - ro[k]) - t[k] for k in range(3)]
    out = np.empty((len(X), 4), dtype=np.float32)
    for k in range(3):
        out[:, k] = ((m[3 * k] * d[0] + m[3 * k + 1] * d[1]) + m[3 * k + 2] * d[2]).astype(np.float32)
    out[:, 3] = inten.astype(np.float32)
    return out
