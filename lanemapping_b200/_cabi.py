"""ctypes binding of include/lm_bev.h.  There is NO fallback: if the shared library is
missing the import of any product entry point raises, and calls on a box without a GPU fail
with the CUDA error the library reports."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

ABI_VERSION = 3
ALGO_BINNED, ALGO_DIRECT, ALGO_SWEEP = 0, 1, 2
STAGE_BIN, STAGE_INDEX, STAGE_REDUCE, STAGE_SWEEP, STAGE_ALL = 1, 2, 4, 8, 15
DEV_ERR_POOL, DEV_ERR_CELL_OVERFLOW = 1, 2


class LmBevParams(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32),
        ("row0", C.c_int32), ("col0", C.c_int32),
        ("bev_img_offset", C.c_float * 2),
        ("img_reso", C.c_float * 2),
        ("local_min_ele", C.c_float),
        ("ele_reso", C.c_float),
        ("inten_min", C.c_int32), ("inten_max", C.c_int32),
        ("n_channels", C.c_int32),
        ("channels", C.c_int32 * 4),
    ]


class LmBevOutputs(C.Structure):
    _fields_ = [
        ("image_dev", C.c_void_p),
        ("count16_dev", C.c_void_p),
        ("proj_dev", C.c_void_p),
        ("acc_dev", C.c_void_p),
        ("acc_band", C.c_int32),
        ("reserved", C.c_int32),
    ]


class LmBevTuning(C.Structure):
    _fields_ = [("bin_ctas_per_sm", C.c_int32), ("red_ctas_per_sm", C.c_int32), ("tile_h_log2", C.c_int32),
                ("max_tiles", C.c_int32), ("stream_hint", C.c_int32), ("use_graph", C.c_int32), ("bin_compact_table", C.c_int32),
                ("reserved", C.c_int32)]


class LmBevSampleGeom(C.Structure):
    _fields_ = [
        ("bev_img_offset", C.c_float * 2),
        ("local_min_ele", C.c_float),
        ("row0", C.c_int32), ("col0", C.c_int32),
        ("reserved", C.c_int32),
    ]


class LmLasXform(C.Structure):
    _fields_ = [
        ("record_length", C.c_int32), ("reserved", C.c_int32),
        ("scale", C.c_double * 3), ("offset", C.c_double * 3),
        ("las_read_offset", C.c_double * 3), ("translation", C.c_double * 3),
        ("rot", C.c_double * 9),
    ]


class LmImg2PcParams(C.Structure):
    _fields_ = [
        ("img_reso", C.c_double * 2), ("bev_img_offset", C.c_double * 2),
        ("ele_reso", C.c_double), ("local_min_ele", C.c_double),
        ("translation", C.c_double * 3), ("quat", C.c_double * 4), ("quat_inv", C.c_double * 4),
        ("las_read_offset", C.c_double * 3),
    ]


class LmJitter(C.Structure):
    _fields_ = [("order", C.c_int32 * 4), ("brightness", C.c_float), ("contrast", C.c_float),
                ("saturation", C.c_float), ("reserved", C.c_float)]


JITTER_PARTIALS = 64


class LmBevStats(C.Structure):
    _fields_ = [
        ("error", C.c_uint32), ("n_chunks", C.c_uint32),
        ("n_valid", C.c_uint64),
        ("n_tiles", C.c_uint32), ("ct_overflow", C.c_uint32), ("reserved", C.c_uint32 * 2),
    ]


SYMBOLS = {
    "lm_bev_abi_version": (C.c_int, []),
    "lm_bev_last_error": (C.c_char_p, []),
    "lm_bev_workspace_bytes": (C.c_int, [C.POINTER(LmBevParams), C.c_int64, C.c_int, C.POINTER(LmBevOutputs),
                                         C.POINTER(C.c_size_t)]),
    "lm_bev_workspace_init": (C.c_int, [C.POINTER(LmBevParams), C.c_int64, C.c_int, C.POINTER(LmBevOutputs), C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "lm_bev_sweep_state_offset": (C.c_int, [C.c_size_t, C.POINTER(C.c_size_t)]),
    "lm_bev_plan_create": (C.c_int, [C.POINTER(LmBevParams), C.c_int64, C.c_int, C.POINTER(LmBevOutputs),
                                     C.POINTER(LmBevTuning), C.POINTER(C.c_void_p)]),
    "lm_bev_plan_workspace_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "lm_bev_plan_init_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lm_bev_plan_rasterize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.POINTER(LmBevOutputs),
                                        C.c_void_p]),
    "lm_bev_plan_destroy": (C.c_int, [C.c_void_p]),
    "lm_bev_rasterize": (C.c_int, [C.POINTER(LmBevParams), C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_size_t,
                                   C.POINTER(LmBevOutputs), C.c_void_p]),
    "lm_bev_rasterize_stages": (C.c_int, [C.POINTER(LmBevParams), C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_size_t,
                                          C.POINTER(LmBevOutputs), C.c_void_p, C.c_int]),
    "lm_bev_workspace_bytes_batch": (C.c_int, [C.POINTER(LmBevParams), C.c_int32, C.c_int64, C.POINTER(LmBevOutputs),
                                               C.POINTER(C.c_size_t)]),
    "lm_bev_rasterize_batch": (C.c_int, [C.POINTER(LmBevParams), C.c_int32, C.POINTER(LmBevSampleGeom),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p, C.c_size_t,
                                         C.POINTER(LmBevOutputs), C.c_void_p]),
    "lm_las_decode": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(LmLasXform), C.c_void_p, C.c_void_p]),
    "lm_bev_rasterize_las": (C.c_int, [C.POINTER(LmBevParams), C.c_void_p, C.c_int64, C.POINTER(LmLasXform), C.c_void_p,
                                       C.c_size_t, C.POINTER(LmBevOutputs), C.c_void_p]),
    "lm_bev_img2pc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lm_label_endpoint_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lm_label_polylines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lm_proj_color_jitter": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(LmJitter), C.c_float, C.c_float,
                                       C.c_void_p, C.c_void_p]),
    "lm_bev_acc_merge": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "lm_bev_finalize": (C.c_int, [C.POINTER(LmBevParams), C.c_void_p, C.c_int32, C.c_int32,
                                  C.POINTER(LmBevOutputs), C.c_void_p]),
    "lm_bev_merge_finalize": (C.c_int, [C.POINTER(LmBevParams), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                        C.POINTER(LmBevOutputs), C.c_void_p]),
    "lm_bev_selftest_div": (C.c_int, [C.c_float, C.c_void_p, C.c_void_p]),
    "lm_bev_selftest_mean": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lm_bev_crop_tiles": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


class LmBevError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liblm_bev error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load liblm_bev.so (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m lanemapping_b200.build` "
                "(nvcc, sm_100a).  lanemapping_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        got = handle.lm_bev_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"liblm_bev ABI {got} != binding ABI {ABI_VERSION}: rebuild")
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise LmBevError(code, lib().lm_bev_last_error().decode("utf-8", "replace"))


def make_las_xform(record_length, scale, offset, las_read_offset=(0.0, 0.0, 0.0), translation=(0.0, 0.0, 0.0),
                   rot=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)) -> LmLasXform:
    x = LmLasXform()
    x.record_length = int(record_length)
    for k in range(3):
        x.scale[k], x.offset[k] = float(scale[k]), float(offset[k])
        x.las_read_offset[k], x.translation[k] = float(las_read_offset[k]), float(translation[k])
    for k in range(9):
        x.rot[k] = float(rot[k])
    return x


def make_params(spec) -> LmBevParams:
    p = LmBevParams()
    p.height, p.width, p.row0, p.col0 = spec.height, spec.width, spec.row0, spec.col0
    p.bev_img_offset[0], p.bev_img_offset[1] = spec.bev_img_offset
    p.img_reso[0], p.img_reso[1] = spec.img_reso
    p.local_min_ele, p.ele_reso = spec.local_min_ele, spec.ele_reso
    p.inten_min, p.inten_max = spec.inten_min, spec.inten_max
    p.n_channels = len(spec.channels)
    for i, c in enumerate(spec.channels):
        p.channels[i] = int(c)
    return p
