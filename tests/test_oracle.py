"""CPU tests of the oracle itself: hand-computed known answers + the algebraic laws the
multi-GPU merge relies on.  Parity is unpinned upstream (no reference rasteriser, SURVEY.md
section 8c), so these are the pins the oracle gets: KATs, properties, and (test_reference_contracts)
the reference's own inverse map / loader contracts."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z
from lanemapping_b200.synth import make_cloud, default_min_ele
from oracle import bev_oracle as O

ALL6 = (CH_MAX_I, CH_MEAN_I, CH_MIN_Z, CH_MAX_Z)


def pts(*rows):
    return np.array(rows, dtype=np.float32).reshape(-1, 4)


def test_kat_single_cell_by_hand():
    # 4x4 grid at 1 m, z step 0.1 m from 0, default intensity clip [800, 33000]
    spec = BevSpec(4, 4, img_reso=(1.0, 1.0), ele_reso=0.1, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    cloud = pts(
        (1.5, 2.5, 0.50, 800.0),     # cell (1,2): zq=5,  iq=0
        (1.2, 2.9, 1.00, 33000.0),   # cell (1,2): zq=10, iq=255
        (1.9, 2.0, 0.25, 16900.0),   # cell (1,2): zq=rint(2.5)=2 (half-even), iq=(16100*255)//32200=127
        (0.0, 0.0, 30.0, 100.0),     # cell (0,0): zq clamps to 255, intensity clips to 800 -> 0
        (3.999, 3.999, -1.0, 70000.0),  # cell (3,3): zq clamps to 0, iq 255
        (4.0, 0.0, 0.0, 5000.0),     # row 4: outside, dropped
        (-0.001, 1.0, 0.0, 5000.0),  # row -1: outside, dropped
    )
    row, col, iq, zq, valid = O.quantise_points(cloud, spec)
    assert valid.tolist() == [True] * 5 + [False] * 2
    assert row[:5].tolist() == [1, 1, 1, 0, 3] and col[:5].tolist() == [2, 2, 2, 0, 3]
    assert zq[:5].tolist() == [5, 10, 2, 255, 0]
    assert iq[:5].tolist() == [0, 255, 127, 0, 255]
    acc = O.accumulate(cloud, spec)
    assert acc[O.ACC_COUNT, 1, 2] == 3 and acc[O.ACC_SUM_Z, 1, 2] == 17 and acc[O.ACC_SUM_I, 1, 2] == 382
    assert acc[O.ACC_MAX_I, 1, 2] == 255 and acc[O.ACC_MIN_Z, 1, 2] == 2 and acc[O.ACC_MAX_Z, 1, 2] == 10
    assert acc[O.ACC_COUNT].sum() == 5
    assert acc[O.ACC_MIN_Z, 2, 2] == 0xFFFFFFFF       # empty cell
    img = O.rasterize(cloud, spec)["image"]
    assert img.shape == (4, 4, 3) and img.dtype == np.uint8
    assert img[1, 2].tolist() == [255, (17 + 1) // 3, 3]     # max_i, mean_z=(17+3//2)//3=6, density
    assert img[0, 0].tolist() == [0, 255, 1]
    assert img[3, 3].tolist() == [255, 0, 1]
    assert img[2, 2].tolist() == [0, 0, 0]                    # empty cell is all-zero (coor_img2pc.py:78)


def test_kat_all_channels_and_count16():
    spec = BevSpec(2, 3, img_reso=(0.5, 0.5), ele_reso=0.05, local_min_ele=-1.0, channels=ALL6, count16=True)
    cloud = pts((0.1, 0.6, -0.5, 1000.0), (0.4, 0.9, 0.0, 30000.0), (0.3, 0.7, -0.75, 2000.0))  # all cell (0,1)
    out = O.rasterize(cloud, spec)
    zq = [10, 20, 5]
    iq = [(200 * 255) // 32200, (29200 * 255) // 32200, (1200 * 255) // 32200]
    assert out["image"][0, 1].tolist() == [max(iq), (sum(iq) + 1) // 3, min(zq), max(zq)]
    assert out["count16"][0, 1] == 3 and out["count16"].sum() == 3
    assert out["image"][1].sum() == 0


def test_density_saturates_and_count16():
    spec = BevSpec(1, 1, img_reso=(1.0, 1.0), channels=(CH_DENSITY,), count16=True)
    cloud = np.tile(pts((0.5, 0.5, 0.0, 900.0)), (70000, 1))
    out = O.rasterize(cloud, spec)
    assert out["image"][0, 0, 0] == 255 and out["count16"][0, 0] == 65535


def test_nan_and_inf_points():
    spec = BevSpec(2, 2, img_reso=(1.0, 1.0), channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    cloud = pts((np.nan, 0.5, 0.0, 900.0), (0.5, np.inf, 0.0, 900.0), (0.5, 0.5, np.nan, np.nan),
                (-np.inf, 0.5, 0.0, 900.0))
    out = O.rasterize(cloud, spec)
    # only the third point is inside; NaN z -> 0, NaN intensity -> inten_min -> 0
    assert out["image"][0, 0].tolist() == [0, 0, 1] and out["image"].sum() == 1


def test_window_is_a_bit_exact_subwindow():
    spec = BevSpec(96, 80, bev_img_offset=(3.25, -7.5), local_min_ele=-2.0, channels=ALL6, count16=True)
    cloud = make_cloud(20000, spec, seed=7, order="shuffled")
    full = O.rasterize(cloud, spec)
    sub = O.rasterize(cloud, spec.window(17, 61, 5, 70))
    assert np.array_equal(sub["image"], full["image"][17:61, 5:70])
    assert np.array_equal(sub["count16"], full["count16"][17:61, 5:70])


def test_intensity_magic_division_is_exact():
    # the CUDA kernel replaces n // d by umulhi(n << 8, m) >> l, l = ceil(log2 d), m = ceil(2^(24+l) / d)
    for d in (1, 2, 3, 255, 256, 32200, 65535, 12345, 32768, 32769):
        l = 0
        while (1 << l) < d:
            l += 1
        m = ((1 << (24 + l)) + d - 1) // d
        assert m < 2**32
        n = np.arange(0, d + 1, dtype=np.uint64) * np.uint64(255)
        assert int(n.max()) < 2**24
        got = ((n << np.uint64(8)) * np.uint64(m) >> np.uint64(32)) >> np.uint64(l)
        assert np.array_equal(got, n // np.uint64(d))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(0, 3000), h=st.integers(1, 40), w=st.integers(1, 40))
def test_permutation_invariance_and_merge_law(seed, n, h, w):
    spec = BevSpec(h, w, img_reso=(0.25, 0.5), ele_reso=0.05, local_min_ele=default_min_ele(BevSpec(h, w)),
                   channels=ALL6, count16=True)
    cloud = make_cloud(n, spec, seed=seed, order="shuffled")
    rng = np.random.default_rng(seed)
    acc = O.accumulate(cloud, spec)
    assert np.array_equal(acc, O.accumulate(cloud[rng.permutation(n)], spec))
    k = int(rng.integers(0, n + 1))
    merged = O.merge_acc(O.accumulate(cloud[:k], spec), O.accumulate(cloud[k:], spec))
    assert np.array_equal(acc, merged)          # the halo-merge law
    # empty cells are all-zero pixels, occupied cells carry count >= 1
    out = O.finalize(acc, spec)
    empty = acc[O.ACC_COUNT] == 0
    assert not out["image"][empty].any() and (out["count16"][~empty] >= 1).all()
    assert int(acc[O.ACC_COUNT].sum()) == int(O.quantise_points(cloud, spec)[4].sum())


def test_synth_cloud_properties():
    spec = BevSpec(1152, 1152, local_min_ele=default_min_ele(BevSpec(1152, 1152)))
    a = make_cloud(200_000, spec, order="scan")
    b = make_cloud(200_000, spec, order="scan")
    assert a.dtype == np.float32 and a.shape == (200_000, 4) and np.array_equal(a, b)
    valid = O.quantise_points(a, spec)[4]
    frac_out = 1.0 - valid.mean()
    assert 0.002 < frac_out < 0.08           # some points fall outside and are dropped
    acc = O.accumulate(a, spec)
    assert (acc[O.ACC_COUNT] == 0).mean() > 0.2      # sparse verges -> empty cells
    assert acc[O.ACC_COUNT].max() >= 4               # dense under the trajectory
    assert (a[:, 3] < 800).any() and (a[:, 3] > 33000).any()   # clip is exercised
    # scan order is spatially coherent along-track
    assert np.abs(np.diff(a[:, 0])).mean() < 2.0
    s = make_cloud(200_000, spec, order="shuffled")
    assert np.abs(np.diff(s[:, 0])).mean() > 10.0


def test_crop_tiles_and_proj():
    spec = BevSpec(30, 50, img_reso=(1.0, 1.0))
    cloud = make_cloud(5000, spec, seed=3, order="shuffled")
    img = O.rasterize(cloud, spec)["image"]
    crops = O.crop_tiles(img, tile=16)
    assert crops.shape == (2 * 4, 16, 16, 3)
    assert np.array_equal(crops[5][:14, :16], img[16:30, 16:32]) and not crops[5][14:].any()
    proj = O.proj_from_image(img)
    assert proj.shape == (3, 30, 50) and proj.dtype == np.float32
    assert proj.max() <= 1.0 and np.array_equal(proj[1] * 255, img[..., 1].astype(np.float32))


def test_pool_variant_matches_single_thread():
    spec = BevSpec(64, 48, img_reso=(0.5, 0.5), channels=ALL6, count16=True, local_min_ele=-2.0)
    cloud = make_cloud(30000, spec, seed=11, order="scan")
    one = O.rasterize(cloud, spec)
    par = O.rasterize_pool(cloud, spec, processes=3)
    assert np.array_equal(one["image"], par["image"]) and np.array_equal(one["count16"], par["count16"])


@pytest.mark.parametrize("height,width,n,procs", [(600, 4608, 300_001, 7), (2304, 1152, 400_003, 5)])
def test_pool_with_scan_point_ranges_matches_single_thread(height, width, n, procs):
    # the bounded index ranges the CPU arm of bench.py hands to its strips (one or several roads
    # scanned one after the other) must lose no point and take none twice
    spec = BevSpec(height, width, local_min_ele=-8.0)
    cloud = make_cloud(n, spec, seed=3, order="scan")
    roads = max(1, int(round(width * spec.img_reso[1] / 57.6)))
    ranges = O.scan_point_ranges(n, spec, procs, roads=roads)
    assert len(ranges) == procs and all(a < b for rs in ranges for a, b in rs)
    assert all(rs[i][1] < rs[i + 1][0] for rs in ranges for i in range(len(rs) - 1))     # disjoint within a strip
    par = O.rasterize_pool(cloud, spec, procs, point_ranges=ranges)
    assert np.array_equal(O.rasterize(cloud, spec)["image"], par["image"])
