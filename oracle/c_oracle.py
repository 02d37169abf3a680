"""ctypes loader for the plain-C oracle (oracle/bev_oracle.c).  TEST INFRASTRUCTURE ONLY --
same access rule as oracle/bev_oracle.py: tests/, smoke() and bench.py's CPU leg."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblm_oracle.so")


class _Spec(ctypes.Structure):
    _fields_ = [("height", ctypes.c_int32), ("width", ctypes.c_int32), ("row0", ctypes.c_int32), ("col0", ctypes.c_int32),
                ("off0", ctypes.c_float), ("off1", ctypes.c_float), ("reso0", ctypes.c_float), ("reso1", ctypes.c_float),
                ("local_min_ele", ctypes.c_float), ("ele_reso", ctypes.c_float),
                ("inten_min", ctypes.c_int32), ("inten_max", ctypes.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "bev_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        L = ctypes.CDLL(LIB)
        L.lmo_accumulate.restype = ctypes.c_int64
        L.lmo_accumulate.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(_Spec), ctypes.c_void_p]
        L.lmo_finalize.restype = ctypes.c_int
        L.lmo_finalize.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Spec), ctypes.c_void_p, ctypes.c_int32,
                                   ctypes.c_void_p, ctypes.c_void_p]
        assert L.lmo_abi_version() == 1
        _lib = L
    return _lib


def _cspec(spec):
    return _Spec(spec.height, spec.width, spec.row0, spec.col0, spec.bev_img_offset[0], spec.bev_img_offset[1],
                 spec.img_reso[0], spec.img_reso[1], spec.local_min_ele, spec.ele_reso, spec.inten_min, spec.inten_max)


def accumulate(pts, spec):
    """-> (uint32 [6, H, W] accumulator planes, number of in-window points)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 4)
    acc = np.empty((6, spec.height, spec.width), dtype=np.uint32)
    cs = _cspec(spec)
    kept = lib().lmo_accumulate(pts.ctypes.data, len(pts), ctypes.byref(cs), acc.ctypes.data)
    return acc, int(kept)


def finalize(acc, spec):
    acc = np.ascontiguousarray(acc, dtype=np.uint32)
    ch = np.asarray(spec.channels, dtype=np.int32)
    img = np.empty((spec.height, spec.width, len(ch)), dtype=np.uint8)
    c16 = np.empty((spec.height, spec.width), dtype=np.uint16) if spec.count16 else None
    cs = _cspec(spec)
    rc = lib().lmo_finalize(acc.ctypes.data, ctypes.byref(cs), ch.ctypes.data, len(ch), img.ctypes.data,
                            c16.ctypes.data if c16 is not None else None)
    if rc != 0:
        raise ValueError("lmo_finalize: unknown channel id")
    return {"image": img, "count16": c16}


def rasterize(pts, spec):
    acc, kept = accumulate(pts, spec)
    out = finalize(acc, spec)
    out["n_valid"] = kept
    return out
