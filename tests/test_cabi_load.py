"""CPU: the C-ABI library builds, loads, exports every symbol include/lm_bev.h declares, and
rejects bad arguments before touching the GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from lanemapping_b200 import BevSpec, _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        text = open(os.path.join(inc, h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_exported_and_bound(native_lib):
    names = declared_symbols()
    assert len(names) >= 7
    for name in names:
        assert hasattr(native_lib, name), f"{name} declared in lm_bev.h but not exported"
    assert sorted(_cabi.SYMBOLS) == names, "ctypes binding and header disagree"
    assert native_lib.lm_bev_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.LmBevParams) == 4 * 4 + 4 * 4 + 2 * 4 + 2 * 4 + 4 + 4 * 4   # 68 bytes, no padding
    assert C.sizeof(_cabi.LmBevOutputs) == 4 * 8 + 8
    assert C.sizeof(_cabi.LmBevStats) == 32 and _cabi.LmBevStats.n_valid.offset == 8
    assert C.sizeof(_cabi.LmBevSampleGeom) == 24
    assert C.sizeof(_cabi.LmLasXform) == 8 + 21 * 8 and _cabi.LmLasXform.rot.offset == 8 + 12 * 8


def test_workspace_bytes_and_argument_errors(native_lib):
    spec = BevSpec(11520, 1152)
    p = _cabi.make_params(spec)
    out = C.c_size_t(0)
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 100_000_000, _cabi.ALGO_BINNED, None, C.byref(out)) == 0
    binned = out.value
    assert 4e8 < binned < 12e9 and binned % 256 == 0      # N*4 B of records + open-chunk slack (upper bound)
    o = _cabi.LmBevOutputs()
    o.image_dev = 1
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 100_000_000, _cabi.ALGO_BINNED, C.byref(o), C.byref(out)) == 0
    assert 4e8 < out.value < binned                       # the real output set needs fewer tiles
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 100_000_000, _cabi.ALGO_DIRECT, None, C.byref(out)) == 0
    assert out.value >= 6 * 4 * spec.cells
    # errors: negative codes + a message, nothing launched
    assert native_lib.lm_bev_workspace_bytes(None, 10, 0, None, C.byref(out)) == -1
    assert b"NULL" in native_lib.lm_bev_last_error()
    p.n_channels = 5
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 10, 0, None, C.byref(out)) == -1
    p = _cabi.make_params(spec)
    p.inten_min = 40000
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 10, 0, None, C.byref(out)) == -1
    p = _cabi.make_params(spec)
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 10, 7, None, C.byref(out)) == -1
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), -1, 0, None, C.byref(out)) == -1
    # a tall raster runs as row windows inside the call; only a raster too WIDE for one tile row is refused
    tall = _cabi.make_params(BevSpec(3_000_000, 1152))
    assert native_lib.lm_bev_workspace_bytes(C.byref(tall), 10, 0, None, C.byref(out)) == 0
    wide = _cabi.make_params(BevSpec(64, 2_000_000))
    assert native_lib.lm_bev_workspace_bytes(C.byref(wide), 10, 0, None, C.byref(out)) == -3
    o = _cabi.LmBevOutputs()
    assert native_lib.lm_bev_rasterize(C.byref(p), None, 0, 0, None, 0, C.byref(o), None) == -1  # no outputs
    o.image_dev = 256
    assert native_lib.lm_bev_rasterize(C.byref(p), None, 0, 0, None, 0, C.byref(o), None) == -2  # no workspace
    assert native_lib.lm_bev_rasterize(C.byref(p), 8, 4, 0, 256, 1 << 30, C.byref(o), None) == -1  # misaligned pts
    assert native_lib.lm_bev_crop_tiles(None, 1, 1, 3, 1152, None, None) == -1
    assert native_lib.lm_bev_acc_merge(None, 0, None, 0, 1, 1, None) == -1
    assert native_lib.lm_bev_finalize(C.byref(p), None, 0, 1, C.byref(o), None) == -1


def test_product_path_has_no_cpu_fallback():
    """The package must not import the oracle, and must refuse CPU tensors."""
    import torch
    from lanemapping_b200 import bev
    pkg = os.path.join(ROOT, "lanemapping_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
    with pytest.raises(RuntimeError, match="CUDA"):
        bev.BevRasterizer(BevSpec(8, 8), 16, device="cpu")


def test_batch_and_las_argument_errors(native_lib):
    """The newer entry points reject bad arguments before any launch (no GPU needed)."""
    spec = BevSpec(1152, 1152)
    p = _cabi.make_params(spec)
    out = C.c_size_t(0)
    o = _cabi.LmBevOutputs()
    o.proj_dev = 1
    assert native_lib.lm_bev_workspace_bytes_batch(C.byref(p), 8, 80_000_000, C.byref(o), C.byref(out)) == 0
    assert out.value % 256 == 0 and out.value > 80_000_000 * 4
    assert native_lib.lm_bev_workspace_bytes_batch(C.byref(p), 0, 10, C.byref(o), C.byref(out)) == -1
    assert native_lib.lm_bev_workspace_bytes_batch(C.byref(p), 8, 10, None, C.byref(out)) == -1
    x = _cabi.make_las_xform(10, (0.001,) * 3, (0.0,) * 3)
    assert native_lib.lm_las_decode(None, 0, C.byref(x), None, None) == -1
    assert b"record_length" in native_lib.lm_bev_last_error()
    x = _cabi.make_las_xform(20, (0.001,) * 3, (0.0,) * 3)
    assert native_lib.lm_las_decode(None, 0, C.byref(x), None, None) == 0          # nothing to do
    assert native_lib.lm_las_decode(None, 5, C.byref(x), None, None) == -1
    assert native_lib.lm_las_decode(None, 0, None, None, None) == -1


def test_plan_api_without_a_gpu(native_lib):
    """lm_bev_plan_*: creation, sizing and argument errors need no device; the tuning knobs are plan fields (the
    library reads no environment variables) and the workspace bound without an output set covers every output set."""
    spec = BevSpec(2304, 1152)
    p = _cabi.make_params(spec)
    o = _cabi.LmBevOutputs()
    o.image_dev = 1
    plan = C.c_void_p()
    assert native_lib.lm_bev_plan_create(C.byref(p), 1_000_000, _cabi.ALGO_BINNED, None, None, C.byref(plan)) == -1   # no output set
    assert native_lib.lm_bev_plan_create(C.byref(p), 1_000_000, 9, C.byref(o), None, C.byref(plan)) == -1           # unknown algo
    assert native_lib.lm_bev_plan_create(C.byref(p), 1_000_000, _cabi.ALGO_BINNED, C.byref(o), None, C.byref(plan)) == 0
    nbytes, legacy = C.c_size_t(0), C.c_size_t(0)
    assert native_lib.lm_bev_plan_workspace_bytes(plan, C.byref(nbytes)) == 0
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 1_000_000, _cabi.ALGO_BINNED, C.byref(o), C.byref(legacy)) == 0
    assert nbytes.value == legacy.value and nbytes.value % 256 == 0
    assert native_lib.lm_bev_plan_rasterize(plan, None, 2_000_000, None, 0, C.byref(o), None) == -1                 # > max_points
    assert native_lib.lm_bev_plan_rasterize(plan, None, 0, None, 0, C.byref(o), None) == -2                         # no workspace
    assert native_lib.lm_bev_plan_destroy(plan) == 0 and native_lib.lm_bev_plan_destroy(None) == 0
    # a tuning field changes the layout the plan sizes for: fewer tiles per launch -> a smaller pool reservation
    t = _cabi.LmBevTuning()
    t.max_tiles = 9
    plan2 = C.c_void_p()
    assert native_lib.lm_bev_plan_create(C.byref(p), 1_000_000, _cabi.ALGO_BINNED, C.byref(o), C.byref(t), C.byref(plan2)) == 0
    small = C.c_size_t(0)
    assert native_lib.lm_bev_plan_workspace_bytes(plan2, C.byref(small)) == 0 and small.value < nbytes.value
    native_lib.lm_bev_plan_destroy(plan2)
    # the sweep algorithm reserves its mailboxes at the END of the workspace
    sw = C.c_size_t(0)
    assert native_lib.lm_bev_workspace_bytes(C.byref(p), 1_000_000, _cabi.ALGO_SWEEP, C.byref(o), C.byref(sw)) == 0
    off = C.c_size_t(0)
    assert native_lib.lm_bev_sweep_state_offset(sw.value, C.byref(off)) == 0
    assert legacy.value <= off.value < sw.value and sw.value - off.value > 296 * 148 * 8 * 32
    assert native_lib.lm_bev_sweep_state_offset(1000, C.byref(off)) == -1
    # the bound without an output set holds for every output set (advisor, round 1): try the three tile heights
    bound = C.c_size_t(0)
    for h, w in ((2304, 1152), (28800, 3456), (1440, 11520), (700, 300)):
        pp = _cabi.make_params(BevSpec(h, w, count16=True))
        assert native_lib.lm_bev_workspace_bytes(C.byref(pp), 5_000_000, _cabi.ALGO_BINNED, None, C.byref(bound)) == 0
        for outs in (("image",), ("image", "count16"), ("acc",), ("proj", "acc")):
            oo = _cabi.LmBevOutputs()
            for k in outs:
                setattr(oo, k + "_dev", 1)
            need = C.c_size_t(0)
            assert native_lib.lm_bev_workspace_bytes(C.byref(pp), 5_000_000, _cabi.ALGO_BINNED, C.byref(oo), C.byref(need)) == 0
            assert need.value <= bound.value, (h, w, outs)
    # compact-table pass (tuning bin_compact_table): automatic on a raster whose tile count leaves direct indexing fewer
    # than three bin CTAs per SM -- the plan then sizes for both layouts; -1 sizes for the direct-indexed kernels alone,
    # +1 asks for the pass on a raster that would not take it (no more open chunks per CTA than the raster has tiles)
    fine, coarse = _cabi.make_params(BevSpec(28800, 3456, count16=True)), _cabi.make_params(BevSpec(2304, 1152))
    oo = _cabi.LmBevOutputs()
    oo.image_dev = 1
    sizes = {}
    for name, pp in (("fine", fine), ("coarse", coarse)):
        for mode in (0, 1, -1):
            tt = _cabi.LmBevTuning()
            tt.bin_compact_table = mode
            pl = C.c_void_p()
            assert native_lib.lm_bev_plan_create(C.byref(pp), 5_000_000, _cabi.ALGO_BINNED, C.byref(oo), C.byref(tt), C.byref(pl)) == 0
            nb = C.c_size_t(0)
            assert native_lib.lm_bev_plan_workspace_bytes(pl, C.byref(nb)) == 0
            sizes[name, mode] = nb.value
            native_lib.lm_bev_plan_destroy(pl)
    assert sizes["fine", 0] == sizes["fine", 1] >= sizes["fine", -1]
    assert sizes["coarse", 0] == sizes["coarse", -1] <= sizes["coarse", 1]
