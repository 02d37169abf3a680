"""The two CPU restatements of the forward spec -- numpy (oracle/bev_oracle.py) and plain C
(oracle/bev_oracle.c) -- were written independently and must agree bit for bit, including on the
inputs where float handling is easy to get wrong (NaN, infinities, cell edges, ties, clip limits)."""
import numpy as np
import pytest

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O
from oracle import c_oracle as C


def _same(cloud, spec):
    acc = O.accumulate(cloud, spec)
    cacc, kept = C.accumulate(cloud, spec)
    assert kept == int(acc[O.ACC_COUNT].sum())
    assert np.array_equal(cacc, acc)
    a, b = O.finalize(acc, spec), C.finalize(cacc, spec)
    assert np.array_equal(a["image"], b["image"])
    if spec.count16:
        assert np.array_equal(a["count16"], b["count16"])


@pytest.mark.parametrize("order", ["scan", "shuffled"])
def test_c_equals_numpy_on_synthetic_cloud(order):
    spec = BevSpec(1152, 1152, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), local_min_ele=default_min_ele(BevSpec(1152, 1152)))
    _same(make_cloud(1_000_000, spec, order=order, seed=5), spec)


def test_c_equals_numpy_on_window_and_all_channels():
    full = BevSpec(700, 500, bev_img_offset=(100.0, -40.0), img_reso=(0.07, 0.13), ele_reso=0.03,
                   channels=(CH_MEAN_I, CH_MIN_Z, CH_MAX_Z, CH_MEAN_Z), count16=True, local_min_ele=-1.5)
    cloud = make_cloud(400_000, full, seed=11, order="shuffled")
    _same(cloud, full)
    _same(cloud, full.window(130, 431, 17, 402))


def test_c_equals_numpy_on_edge_inputs():
    spec = BevSpec(130, 129, img_reso=(1.0, 1.0), ele_reso=0.1, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), count16=True)
    rng = np.random.default_rng(4)
    n = 200_000
    pts = np.empty((n, 4), dtype=np.float32)
    pts[:, 0] = rng.integers(-2, 133, n) + rng.choice([0.0, 0.5, np.nextafter(np.float32(1), np.float32(0))], n)
    pts[:, 1] = rng.integers(-2, 132, n) + rng.choice([0.0, 0.25, 0.999999], n)
    pts[:, 2] = rng.integers(-10, 300, n) * 0.05          # ties at .5 of ele_reso: round-half-even
    pts[:, 3] = rng.choice([0, 799, 800, 801, 32999, 33000, 33001, 65535, 12345.75], n)
    special = rng.integers(0, n, 4000)
    pts[special[:1000], 0] = np.nan
    pts[special[1000:2000], 2] = np.nan
    pts[special[2000:2500], 3] = np.nan
    pts[special[2500:3000], 1] = np.inf
    pts[special[3000:3500], 2] = -np.inf
    pts[special[3500:], 3] = np.inf
    _same(pts, spec)
    _same(pts[:0], spec)                                   # empty input: all-zero raster
    out = C.rasterize(pts[:0], spec)
    assert out["n_valid"] == 0 and not out["image"].any() and not out["count16"].any()
