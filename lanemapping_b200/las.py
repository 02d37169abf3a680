"""Minimal uncompressed LAS reader/writer (numpy only).

The reference reads clouds with ``laspy.read`` and uses ``las.x/y/z`` (scaled float64) and
``las.intensity`` (u16) -- reference baseline/datasets/laserlane_proposals.py:618-636.  laspy is
not installed here, so this module parses the public LAS 1.0-1.4 header and the fixed leading
fields shared by every point data record format (X, Y, Z int32 + intensity u16).  ``.laz`` is
not supported.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Tuple

import numpy as np


@dataclass(frozen=True)
class LasHeader:
    version: Tuple[int, int]
    offset_to_points: int
    point_format: int
    record_length: int
    n_points: int
    scale: Tuple[float, float, float]
    offset: Tuple[float, float, float]
    mins: Tuple[float, float, float] = (float("nan"),) * 3      # header bounds (min x, y, z); NaN if absent
    maxs: Tuple[float, float, float] = (float("nan"),) * 3

    def bounds_ok(self) -> bool:
        """Are the header's min/max bounds usable (finite, ordered, not the all-zero of a lazy writer)?"""
        import math
        vals = list(self.mins) + list(self.maxs)
        if not all(math.isfinite(v) for v in vals):
            return False
        if any(lo > hi for lo, hi in zip(self.mins, self.maxs)):
            return False
        return any(v != 0.0 for v in vals)


def read_header(buf: bytes) -> LasHeader:
    if buf[:4] != b"LASF":
        raise ValueError("not a LAS file (missing LASF signature)")
    major, minor = buf[24], buf[25]
    offset_to_points, = struct.unpack_from("<I", buf, 96)
    fmt = buf[104] & 0x3F                       # top bits flag compression (LAZ)
    if buf[104] & 0xC0:
        raise ValueError("compressed LAS (.laz) is not supported")
    reclen, = struct.unpack_from("<H", buf, 105)
    n, = struct.unpack_from("<I", buf, 107)
    scale = struct.unpack_from("<3d", buf, 131)
    offset = struct.unpack_from("<3d", buf, 155)
    if (major, minor) >= (1, 4) and len(buf) >= 255:
        n64, = struct.unpack_from("<Q", buf, 247)
        if n64:
            n = n64
    if reclen < 14:
        raise ValueError("LAS record length < 14")
    mins = maxs = (float("nan"),) * 3
    if len(buf) >= 227:                         # max x, min x, max y, min y, max z, min z
        b = struct.unpack_from("<6d", buf, 179)
        maxs, mins = (b[0], b[2], b[4]), (b[1], b[3], b[5])
    return LasHeader((major, minor), offset_to_points, fmt, reclen, int(n), scale, offset, mins, maxs)


def world_min(raw: np.ndarray, hdr: LasHeader) -> Tuple[float, float, float]:
    """Exact minimum of the world coordinates from the int32 X/Y/Z fields of the point block (one
    strided integer pass, then ONE scale+offset in float64): what a reader falls back to when the
    header bounds are missing."""
    n = hdr.n_points
    ixyz = np.ndarray((n, 3), dtype="<i4", buffer=raw, strides=(hdr.record_length, 4))
    lo = ixyz.min(axis=0).astype(np.float64) if n else np.zeros(3)
    return tuple(float(v) for v in lo * np.asarray(hdr.scale) + np.asarray(hdr.offset))


def read_las(path: str):
    """-> (xyz float64 [N,3] in world coordinates, intensity uint16 [N], header)."""
    with open(path, "rb") as f:
        head = f.read(375)
        hdr = read_header(head)
        f.seek(hdr.offset_to_points)
        raw = np.fromfile(f, dtype=np.uint8, count=hdr.n_points * hdr.record_length)
    if raw.size != hdr.n_points * hdr.record_length:
        raise ValueError(f"{path}: truncated point data")
    rec = raw.reshape(hdr.n_points, hdr.record_length)
    ixyz = np.ascontiguousarray(rec[:, :12]).view("<i4").reshape(-1, 3)
    inten = np.ascontiguousarray(rec[:, 12:14]).view("<u2").reshape(-1)
    xyz = ixyz.astype(np.float64) * np.asarray(hdr.scale) + np.asarray(hdr.offset)
    return xyz, inten, hdr


def write_las(path: str, xyz: np.ndarray, intensity: np.ndarray, scale=(0.001, 0.001, 0.001), offset=None) -> None:
    """LAS 1.2, point data record format 0 (20-byte records).  For tests and synthetic data."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = len(xyz)
    intensity = np.asarray(intensity).reshape(-1).astype("<u2")
    if offset is None:
        offset = np.floor(xyz.min(axis=0)) if n else np.zeros(3)
    offset = np.asarray(offset, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    ixyz = np.rint((xyz - offset) / scale).astype("<i4")
    head = bytearray(227)
    head[0:4] = b"LASF"
    head[24], head[25] = 1, 2
    head[26:58] = b"lanemapping_b200".ljust(32, b"\0")
    head[58:90] = b"lanemapping_b200.las".ljust(32, b"\0")
    struct.pack_into("<H", head, 94, 227)       # header size
    struct.pack_into("<I", head, 96, 227)       # offset to point data
    struct.pack_into("<I", head, 100, 0)        # VLRs
    head[104] = 0
    struct.pack_into("<H", head, 105, 20)
    struct.pack_into("<I", head, 107, n)
    struct.pack_into("<I", head, 111, n)        # points by return [0]
    struct.pack_into("<3d", head, 131, *scale)
    struct.pack_into("<3d", head, 155, *offset)
    if n:
        w = ixyz.astype(np.float64) * scale + offset
        struct.pack_into("<6d", head, 179, w[:, 0].max(), w[:, 0].min(), w[:, 1].max(), w[:, 1].min(),
                         w[:, 2].max(), w[:, 2].min())
    rec = np.zeros((n, 20), dtype=np.uint8)
    rec[:, :12] = ixyz.view(np.uint8).reshape(n, 12)
    rec[:, 12:14] = intensity.view(np.uint8).reshape(n, 2)
    rec[:, 14] = 0x09                            # return 1 of 1
    with open(path, "wb") as f:
        f.write(bytes(head))
        f.write(rec.tobytes())


def read_point_block(path: str):
    """-> (uint8 [n_points * record_length] exactly as on disk, header): the input of the GPU
    decode (``bev.decode_las`` / ``BevRasterizer.rasterize_las``); nothing is scaled on the host."""
    with open(path, "rb") as f:
        hdr = read_header(f.read(375))
        f.seek(hdr.offset_to_points)
        raw = np.fromfile(f, dtype=np.uint8, count=hdr.n_points * hdr.record_length)
    if raw.size != hdr.n_points * hdr.record_length:
        raise ValueError(f"{path}: truncated point data")
    return raw, hdr
