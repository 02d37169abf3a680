"""GPU parity: the CUDA path (through the C-ABI) against the oracle on the same seeded
inputs.  Bar: bit-exact on every integer output (cells, counts, sums, max/min, u8 channels,
u16 count plane); f32 proj bit-exact too (it is a single IEEE division of a u8 by 255)."""
import numpy as np
import pytest
import torch

from lanemapping_b200 import (BevSpec, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z)
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O

pytestmark = pytest.mark.gpu

CFG2_CH = (CH_MAX_I, CH_MEAN_Z, CH_DENSITY)
CFG4_CH = (CH_MAX_I, CH_MEAN_I, CH_MIN_Z, CH_MAX_Z)


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    assert torch.cuda.is_available()
    return B


def run_gpu(bev, cloud, spec, algo, outputs):
    pts = torch.from_numpy(np.ascontiguousarray(cloud)).cuda()
    r = bev.BevRasterizer(spec, max(1, len(cloud)), algo=algo, outputs=outputs)
    out = r(pts)
    torch.cuda.synchronize()
    st = r.stats()
    assert st["error"] == 0, st
    return {k: v.cpu().numpy() for k, v in out.items()}, st


def assert_matches_oracle(bev, cloud, spec, algo, with_acc=False):
    outputs = ["image", "proj"] + (["count16"] if spec.count16 else []) + (["acc"] if with_acc else [])
    got, st = run_gpu(bev, cloud, spec, algo, outputs)
    acc = O.accumulate(cloud, spec)
    want = O.finalize(acc, spec)
    assert st["n_valid"] == int(acc[O.ACC_COUNT].sum())
    assert np.array_equal(got["image"], want["image"]), "u8 image differs"
    if spec.count16:
        assert np.array_equal(got["count16"], want["count16"]), "count16 differs"
    assert np.array_equal(got["proj"], O.proj_from_image(want["image"])), "f32 proj differs"
    if with_acc:
        assert np.array_equal(got["acc"].view(np.uint32), acc), "raw accumulators differ"
    return got


@pytest.mark.parametrize("algo", ["direct", "binned"])
@pytest.mark.parametrize("order", ["scan", "shuffled"])
def test_cfg2_like_segment(bev, algo, order):
    # 2 x 1 crops of the cfg-2 geometry, 2 M points: oracle finishes in seconds
    spec = BevSpec(2304, 1152, channels=CFG2_CH, local_min_ele=default_min_ele(BevSpec(2304, 1152)))
    cloud = make_cloud(2_000_000, spec, order=order)
    assert_matches_oracle(bev, cloud, spec, algo)


@pytest.mark.parametrize("algo", ["direct", "binned"])
def test_cfg1_single_channel_tile(bev, algo):
    spec = BevSpec(1152, 1152, channels=(CH_MAX_I,), local_min_ele=default_min_ele(BevSpec(1152, 1152)))
    cloud = make_cloud(1_000_000, spec, order="scan")
    assert_matches_oracle(bev, cloud, spec, algo)


@pytest.mark.parametrize("algo", ["direct", "binned"])
@pytest.mark.parametrize("order", ["scan", "shuffled"])
def test_cfg4_fine_resolution_five_outputs(bev, algo, order):
    base = BevSpec(2304, 1152, img_reso=(0.02, 0.02), ele_reso=0.02, channels=CFG4_CH, count16=True)
    spec = BevSpec(2304, 1152, img_reso=(0.02, 0.02), ele_reso=0.02, channels=CFG4_CH, count16=True,
                   local_min_ele=default_min_ele(base))
    cloud = make_cloud(1_500_000, spec, order=order)
    assert_matches_oracle(bev, cloud, spec, algo)


@pytest.mark.parametrize("algo", ["direct", "binned"])
def test_raw_accumulators_all_planes(bev, algo):
    spec = BevSpec(300, 260, img_reso=(0.1, 0.1), channels=CFG2_CH, count16=True, local_min_ele=-2.0,
                   bev_img_offset=(12.5, -3.25))
    cloud = make_cloud(400_000, spec, seed=5, order="shuffled")
    assert_matches_oracle(bev, cloud, spec, algo, with_acc=True)


@pytest.mark.parametrize("algo", ["direct", "binned"])
@pytest.mark.parametrize("channels", [(CH_DENSITY,), (CH_MIN_Z,), (CH_MEAN_I, CH_MAX_Z), (CH_MEAN_Z, CH_MIN_Z, CH_MAX_I, CH_DENSITY),
                                      (CH_MAX_Z, CH_MAX_Z, CH_MAX_I)])
def test_channel_sets_on_ragged_grid(bev, algo, channels):
    # dimensions that are not multiples of the 128/64 shared-memory tile nor of 4 bytes
    spec = BevSpec(201, 333, img_reso=(0.07, 0.13), ele_reso=0.03, channels=channels, local_min_ele=-1.5,
                   bev_img_offset=(-5.0, 2.0), count16=(len(channels) % 2 == 0))
    cloud = make_cloud(300_000, spec, seed=9, order="shuffled")
    assert_matches_oracle(bev, cloud, spec, algo)


@pytest.mark.parametrize("algo", ["direct", "binned"])
def test_window_into_global_grid(bev, algo):
    full = BevSpec(700, 500, bev_img_offset=(100.0, -40.0), channels=CFG2_CH, local_min_ele=-2.0)
    cloud = make_cloud(500_000, full, seed=21, order="scan")
    sub = full.window(130, 577, 64, 450)
    got = assert_matches_oracle(bev, cloud, sub, algo)
    whole = O.rasterize(cloud, full)["image"]
    assert np.array_equal(got["image"], whole[130:577, 64:450])


@pytest.mark.parametrize("algo", ["direct", "binned"])
def test_edge_inputs(bev, algo):
    spec = BevSpec(130, 129, img_reso=(1.0, 1.0), ele_reso=0.1, channels=CFG2_CH, count16=True)
    # empty cloud -> all-zero raster
    got, st = run_gpu(bev, np.zeros((0, 4), np.float32), spec, algo, ["image", "count16"])
    assert not got["image"].any() and not got["count16"].any() and st["n_valid"] == 0
    # every point outside / NaN -> dropped, never clamped
    bad = np.array([[-1.0, 5, 0, 900], [5, 1e9, 0, 900], [np.nan, 5, 0, 900], [5, np.nan, 0, 900],
                    [np.inf, 5, 0, 900], [130.0, 5, 0, 900], [5, 129.0, 0, 900]], np.float32)
    got, st = run_gpu(bev, bad, spec, algo, ["image"])
    assert not got["image"].any() and st["n_valid"] == 0
    # NaN z / NaN intensity on an inside point, exact cell boundaries, negative zero
    edge = np.array([[5.5, 5.5, np.nan, np.nan], [129.99999, 128.99999, 1.0, 1e9], [-0.0, 0.0, 0.05, 800.0],
                     [64.0, 64.0, 0.25, 33000.0], [63.999996, 63.999996, 0.35, 32999.0]], np.float32)
    assert_matches_oracle(bev, edge, spec, algo, with_acc=True)
    # 300k points in ONE cell: hot-cell atomics, density and count16 saturation
    hot = np.tile(np.array([[7.5, 9.5, 1.23, 20000.0]], np.float32), (300_000, 1))
    hot[::3, 2] = 4.0
    hot[1::3, 3] = 2500.0
    got = assert_matches_oracle(bev, hot, spec, algo, with_acc=True)
    assert got["image"][7, 9, 2] == 255 and got["count16"][7, 9] == 65535


@pytest.mark.parametrize("channels,count16", [(CFG2_CH, False), (CFG4_CH, True), ((CH_MEAN_I,), False)])
@pytest.mark.parametrize("hot", [4095, 4096, 4097, 8192, 70_000])
def test_packed_count_wrap_is_found_and_redone(bev, channels, count16, hot):
    """reduce_tiles packs [count:12 | sum:20] in one word and checks conservation per tile: cells with
    exactly 4095 points stay packed, 4096 and beyond must trigger the unpacked redo (also several cells
    wrapping in one tile, and tiles without a wrap next to them)."""
    spec = BevSpec(200, 300, img_reso=(1.0, 1.0), ele_reso=0.1, channels=channels, count16=count16)
    rng = np.random.default_rng(hot)
    bg = np.stack([rng.uniform(0, 200, 40_000), rng.uniform(0, 300, 40_000), rng.uniform(0, 25, 40_000),
                   rng.uniform(0, 40000, 40_000)], axis=1).astype(np.float32)
    cells = [(7.5, 9.5), (70.5, 140.5), (71.5, 140.5)] if hot > 4096 else [(7.5, 9.5)]
    parts = [bg]
    for k, (x, y) in enumerate(cells):
        h = np.tile(np.array([[x, y, 1.0, 20000.0]], np.float32), (hot + k, 1))
        h[:, 2] = rng.uniform(0, 25, hot + k)
        h[:, 3] = rng.uniform(0, 40000, hot + k)
        parts.append(h)
    cloud = np.concatenate(parts)
    rng.shuffle(cloud)
    assert_matches_oracle(bev, cloud, spec, "binned")


def test_binned_equals_direct_and_is_deterministic(bev):
    spec = BevSpec(1152, 1152, channels=CFG2_CH, local_min_ele=default_min_ele(BevSpec(1152, 1152)), count16=True)
    cloud = make_cloud(3_000_000, spec, seed=77, order="shuffled")
    a, _ = run_gpu(bev, cloud, spec, "binned", ["image", "count16", "acc"])
    b, _ = run_gpu(bev, cloud, spec, "binned", ["image", "count16", "acc"])
    d, _ = run_gpu(bev, cloud, spec, "direct", ["image", "count16", "acc"])
    for k in a:
        assert np.array_equal(a[k], b[k]), f"binned run-to-run difference in {k}"
        assert np.array_equal(a[k], d[k]), f"binned != direct in {k}"


def test_workspace_reuse_across_calls(bev):
    spec = BevSpec(640, 384, channels=CFG2_CH, local_min_ele=-2.0)
    r = bev.BevRasterizer(spec, 600_000, outputs=["image"])
    for seed, n in ((1, 600_000), (2, 1234), (3, 0), (4, 450_001)):
        cloud = make_cloud(n, spec, seed=seed, order="scan")
        out = r(torch.from_numpy(cloud).cuda())
        torch.cuda.synchronize()
        r.check_device_errors()
        assert np.array_equal(out["image"].cpu().numpy(), O.rasterize(cloud, spec)["image"])


def test_merge_finalize_and_crops(bev):
    spec = BevSpec(500, 300, img_reso=(0.1, 0.1), channels=CFG4_CH, count16=True, local_min_ele=-2.0)
    cloud = make_cloud(400_000, spec, seed=13, order="shuffled")
    half = len(cloud) // 3
    ra = bev.BevRasterizer(spec, len(cloud), outputs=["acc"])
    a = ra(torch.from_numpy(cloud[:half]).cuda())["acc"]
    b = bev.BevRasterizer(spec, len(cloud), algo="direct", outputs=["acc"])(torch.from_numpy(cloud[half:]).cuda())["acc"]
    # halo-merge law on a row band (views into the full planes), then the whole raster
    bev.acc_merge_(a[:, 100:228], b[:, 100:228])
    want = O.accumulate(cloud[:half], spec)
    other = O.accumulate(cloud[half:], spec)
    want[:, 100:228] = O.merge_acc(want[:, 100:228], other[:, 100:228])
    torch.cuda.synchronize()
    assert np.array_equal(a.cpu().numpy().view(np.uint32), want)
    bev.acc_merge_(a[:, :100], b[:, :100])
    bev.acc_merge_(a[:, 228:], b[:, 228:])
    out = {"image": torch.zeros((500, 300, 4), dtype=torch.uint8, device="cuda"),
           "count16": torch.zeros((500, 300), dtype=torch.uint16, device="cuda")}
    bev.finalize_rows(spec, a, 37, 411, out)
    full = O.rasterize(cloud, spec)
    torch.cuda.synchronize()
    assert np.array_equal(out["image"].cpu().numpy()[37:411], full["image"][37:411])
    assert np.array_equal(out["count16"].cpu().numpy()[37:411], full["count16"][37:411])
    assert not out["image"][:37].any() and not out["image"][411:].any()
    crops = bev.crop_tiles(torch.from_numpy(full["image"]).cuda(), tile=128)
    assert np.array_equal(crops.cpu().numpy(), O.crop_tiles(full["image"], tile=128))


def test_host_rasterizer_e2e(bev):
    spec = BevSpec(1152, 1152, channels=CFG2_CH, local_min_ele=default_min_ele(BevSpec(1152, 1152)))
    cloud = make_cloud(500_000, spec, seed=3)
    hr = bev.HostRasterizer(spec, len(cloud))
    pinned = torch.from_numpy(cloud).pin_memory()
    img = hr(pinned)["image"]
    assert np.array_equal(img, O.rasterize(cloud, spec)["image"])
    assert hr.h2d_bytes == len(cloud) * 16 and hr.d2h_bytes == spec.cells * 3


@pytest.mark.parametrize("divisor", [0.05, 0.02, 0.1, 0.07, 0.13, 0.03, 1.0, 0.25, 0.5, 1.0 / 3.0, 0.001, 1000.0,
                                     0.0499999, 1.9999999, 3.0e-6, 7.0e5])
def test_exact_division_exhaustive(bev, native_lib, divisor):
    """The kernels divide by img_reso / ele_reso with a 3-operation sequence (FMUL, FFMA, FFMA).
    It must be bit-identical to the IEEE division the spec (and the numpy oracle) uses: compare
    against __fdiv_rn for every one of the 2^32 binary32 dividends."""
    import ctypes as C
    from lanemapping_b200 import _cabi
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    _cabi.check(native_lib.lm_bev_selftest_div(C.c_float(divisor), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    bad, fast = (int(v) for v in out.cpu())
    assert fast == 104 * 2**24          # dividends with exponent in [2^-40, 2^64), both signs
    assert bad == 0, f"{bad} dividends where the fast division differs from __fdiv_rn for divisor {divisor}"


@pytest.mark.parametrize("band", [64, 100])
def test_banded_raw_accumulators(bev, band):
    """Halo mode: raw accumulators only on the first/last `band` rows (6-plane kernel on the band tiles,
    light kernel elsewhere); image everywhere.  Band rows of acc and the whole image must be exact."""
    spec = BevSpec(900, 700, img_reso=(0.1, 0.1), channels=CFG2_CH, local_min_ele=-2.0)
    cloud = make_cloud(800_000, spec, seed=41, order="scan")
    r = bev.BevRasterizer(spec, len(cloud), outputs=["image", "acc"], acc_band=band)
    out = r(torch.from_numpy(cloud).cuda())
    torch.cuda.synchronize()
    r.check_device_errors()
    acc = O.accumulate(cloud, spec)
    got = out["acc"].cpu().numpy().view(np.uint32)
    assert np.array_equal(out["image"].cpu().numpy(), O.finalize(acc, spec)["image"])
    for plane in (O.ACC_COUNT, O.ACC_SUM_Z, O.ACC_MAX_I):      # the planes max_i / mean_z / density need
        assert np.array_equal(got[plane, :band], acc[plane, :band]) and np.array_equal(got[plane, -band:], acc[plane, -band:])


def test_exact_mean_exhaustive(bev, native_lib):
    """reduce_tiles derives mean channels with a float reciprocal when count <= 4095; it must equal
    the integer (sum + count/2) // count for every possible (count, sum)."""
    from lanemapping_b200 import _cabi
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    _cabi.check(native_lib.lm_bev_selftest_mean(out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    bad, n = (int(v) for v in out.cpu())
    assert n == sum(255 * c + 1 for c in range(1, 4096)) and bad == 0


def test_plan_graph_replay_and_tuning_fields(bev):
    """BevRasterizer runs on the plan API (include/lm_bev.h): with ``graph=True`` a call that repeats the previous
    call's buffers replays a captured CUDA graph, a call with other buffers re-captures; the tuning knobs are plan
    fields (the library reads no environment variables) and never change a result."""
    from oracle import c_oracle as CO
    spec = BevSpec(1152, 1152, local_min_ele=default_min_ele(BevSpec(1152, 1152)))
    clouds = [make_cloud(600_000 + 10_000 * i, spec, seed=70 + i, order="scan" if i % 2 else "shuffled") for i in range(3)]
    want = [CO.rasterize(c, spec)["image"] for c in clouds]
    dev = [torch.from_numpy(c).cuda() for c in clouds]
    r = bev.BevRasterizer(spec, max(len(c) for c in clouds), graph=True)
    out = r.alloc_outputs()
    for rep in range(3):                                   # same arguments: capture once, replay twice
        out["image"].fill_(9)
        r(dev[0], out=out)
        assert np.array_equal(out["image"].cpu().numpy(), want[0])
    for i in (1, 2, 0):                                    # other points / sizes: re-captured, still exact
        got = r(dev[i], out=out)
        assert np.array_equal(got["image"].cpu().numpy(), want[i])
    r.check_device_errors()
    for tuning in ({"bin_ctas_per_sm": 2, "red_ctas_per_sm": 1}, {"tile_h_log2": 5, "stream_hint": 1}, {"max_tiles": 20}):
        rt = bev.BevRasterizer(spec, len(clouds[1]), tuning=tuning)
        assert np.array_equal(rt(dev[1])["image"].cpu().numpy(), want[1]), tuning
    with pytest.raises(ValueError):
        bev.BevRasterizer(spec, 10, tuning={"no_such_knob": 1})


def _ct_run(bev, cloud, spec, outputs, tuning):
    r = bev.BevRasterizer(spec, max(1, len(cloud)), outputs=outputs, tuning=tuning)
    out = r(torch.from_numpy(np.ascontiguousarray(cloud)).cuda())
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, r.stats()


@pytest.mark.parametrize("order", ["scan", "shuffled"])
@pytest.mark.parametrize("channels,count16", [(CFG2_CH, False), (CFG4_CH, True)])
def test_compact_table_bin_pass(bev, order, channels, count16):
    """bin_points' compact-table variant (tuning bin_compact_table = 1; automatic on rasters whose tile count leaves
    direct indexing fewer than three CTAs per SM): same bits as the oracle, a scan-ordered cloud never overflows the
    table, and the stats count every point once."""
    from oracle import c_oracle as CO
    spec = BevSpec(2304, 1152, channels=channels, count16=count16, local_min_ele=default_min_ele(BevSpec(2304, 1152)))
    cloud = make_cloud(1_500_000, spec, order=order, seed=5)
    outputs = ["image", "proj"] + (["count16"] if count16 else [])
    want = CO.rasterize(cloud, spec)
    for tuning in ({"bin_compact_table": 1}, {"bin_compact_table": 1, "tile_h_log2": 6}):
        got, st = _ct_run(bev, cloud, spec, outputs, tuning)
        assert st["error"] == 0 and st["ct_overflow"] == 0, st      # 162..648 tiles: the table holds them all
        assert st["n_valid"] == int(want["n_valid"])
        assert np.array_equal(got["image"], want["image"])
        if count16:
            assert np.array_equal(got["count16"], want["count16"])
    off, st_off = _ct_run(bev, cloud, spec, outputs, {"bin_compact_table": -1})
    assert np.array_equal(off["image"], want["image"]) and st_off["n_valid"] == st["n_valid"]


@pytest.mark.parametrize("max_tiles", [0, 700])
def test_compact_table_overflow_falls_back_exactly(bev, max_tiles):
    """A cloud in no spatial order on a raster of 4608 tiles: every bin CTA meets more tiles than its 1024-slot table
    holds, raises stats.ct_overflow, and the direct-indexed kernels queued behind the pass (row windows included)
    redo the raster -- same bits, every point counted once.  The same raster in scan order stays on the pass."""
    spec = BevSpec(4608, 4096, channels=CFG4_CH, count16=True, local_min_ele=default_min_ele(BevSpec(4608, 4096)))
    outputs = ["image", "count16"]
    tuning = {"bin_compact_table": 1}
    if max_tiles:
        tuning["max_tiles"] = max_tiles
    for order, overflow in (("shuffled", 1), ("scan", 0)):
        cloud = make_cloud(3_000_000, spec, order=order, seed=11)
        ref, st_ref = run_gpu(bev, cloud, spec, "direct", outputs)
        got, st = _ct_run(bev, cloud, spec, outputs, tuning)
        assert st["error"] == 0 and st["ct_overflow"] == overflow, (order, st)
        assert st["n_valid"] == st_ref["n_valid"]
        assert np.array_equal(got["image"], ref["image"]) and np.array_equal(got["count16"], ref["count16"])


def test_compact_table_with_banded_accumulators(bev):
    spec = BevSpec(1440, 2304, channels=CFG2_CH, local_min_ele=default_min_ele(BevSpec(1440, 2304)))
    cloud = make_cloud(1_000_000, spec, order="scan", seed=3)
    pts = torch.from_numpy(cloud).cuda()
    outs = {}
    for mode in (1, -1):
        r = bev.BevRasterizer(spec, len(cloud), outputs=("image", "acc"), acc_band=64, tuning={"bin_compact_table": mode})
        o = r(pts)
        torch.cuda.synchronize()
        assert r.stats()["ct_overflow"] == 0
        outs[mode] = {k: v.cpu().numpy() for k, v in o.items()}
    assert np.array_equal(outs[1]["image"], outs[-1]["image"])
    band = np.r_[0:64, 1440 - 64:1440]
    for plane in (O.ACC_COUNT, O.ACC_SUM_Z, O.ACC_MAX_I):      # the planes max_i / mean_z / density need
        assert np.array_equal(outs[1]["acc"][plane][band], outs[-1]["acc"][plane][band])


def test_shorter_last_row_window_fits_the_pool(bev):
    """Row windows: the record pool is sized for the first window's tile count.  A shorter last window has fewer
    tiles, fits more bin CTAs per SM and would reserve MORE open chunks than that (200 tile rows x 27 = 5400 tiles run
    two CTAs per SM, the last 192 x 27 = 5184 three): the launch then takes fewer CTAs instead of failing."""
    spec = BevSpec(12544, 3456, img_reso=(0.02, 0.02), channels=(CH_MAX_I,), local_min_ele=default_min_ele(BevSpec(12544, 3456)))
    cloud = make_cloud(2_000_000, spec, order="scan", seed=21)
    ref, st_ref = run_gpu(bev, cloud, spec, "direct", ["image"])
    got, st = _ct_run(bev, cloud, spec, ["image"], {"max_tiles": 5400, "tile_h_log2": 5, "bin_compact_table": -1})
    assert st["error"] == 0 and st["n_valid"] == st_ref["n_valid"]
    assert np.array_equal(got["image"], ref["image"])
