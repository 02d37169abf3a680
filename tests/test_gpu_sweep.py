"""LM_ALGO_AUTO: the single-pass sweep (csrc/lm_sweep.cuh) against the oracle, and its fall-back.

The sweep is exact by construction for row-ordered clouds and hands any other cloud to the two-pass kernels
queued behind it, so ``algo="sweep"`` must equal the oracle bit for bit on EVERY input; the persistent counters
(``sweep_state``) tell which of the two did the work."""
import numpy as np
import pytest
import torch

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_I, CH_MEAN_Z
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    return B


def spec_of(h, w, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), **kw):
    return BevSpec(h, w, channels=channels, local_min_ele=default_min_ele(BevSpec(1152, 1152)), **kw)


def check(r, cloud, spec, outs=("image",)):
    got = r(torch.from_numpy(cloud).cuda())
    torch.cuda.synchronize()
    want = O.rasterize(cloud, spec)
    st = r.stats()
    assert st["error"] == 0 and st["n_valid"] == int(O.accumulate(cloud, spec)[O.ACC_COUNT].sum())
    assert np.array_equal(got["image"].cpu().numpy(), want["image"])
    if "proj" in outs:
        assert np.array_equal(got["proj"].cpu().numpy(), O.proj_from_image(want["image"]))
    if "count16" in outs:
        assert np.array_equal(got["count16"].cpu().numpy(), want["count16"])


@pytest.mark.parametrize("h,w,n", [(2304, 1152, 6_000_000), (1152, 1152, 1_000_000), (700, 333, 300_000), (4000, 1000, 50_000)])
def test_sweep_is_bit_exact_on_scan_ordered_clouds(bev, h, w, n):
    spec = spec_of(h, w)
    cloud = make_cloud(n, spec, seed=5, order="scan")
    r = bev.BevRasterizer(spec, n, algo="sweep", outputs=("image", "proj"))
    for rep in range(3):                                        # the mailboxes carry their state from call to call
        check(r, cloud, spec, ("image", "proj"))
    s = r.sweep_state()
    assert s == {"cooldown": 0, "n_failed": 0, "n_ok": 3}, s    # the sweep did all three


@pytest.mark.parametrize("channels,count16", [((CH_MAX_I,), False), ((CH_MAX_I, CH_DENSITY), False),
                                              ((CH_DENSITY, CH_MEAN_Z, CH_MAX_I, CH_MAX_I), True)])
def test_sweep_channel_sets(bev, channels, count16):
    spec = spec_of(1500, 1100, channels=channels, count16=count16)
    cloud = make_cloud(2_000_000, spec, seed=9, order="scan")
    outs = ("image", "count16") if count16 else ("image",)
    r = bev.BevRasterizer(spec, len(cloud), algo="sweep", outputs=outs)
    check(r, cloud, spec, outs)
    assert r.sweep_state()["n_ok"] == 1


def test_unordered_cloud_falls_back_and_stays_exact(bev):
    spec = spec_of(2304, 1152)
    shuffled = make_cloud(2_000_000, spec, seed=6, order="shuffled")
    ordered = make_cloud(2_000_000, spec, seed=7, order="scan")
    r = bev.BevRasterizer(spec, 2_000_000, algo="sweep")
    check(r, shuffled, spec)                                    # sweep gives up -> two-pass kernels
    s = r.sweep_state()
    assert s["n_failed"] == 1 and s["n_ok"] == 0 and s["cooldown"] > 0
    check(r, ordered, spec)                                     # switched off for a while: still exact
    s2 = r.sweep_state()
    assert s2["cooldown"] == s["cooldown"] - 1 and s2["n_ok"] == 0
    for _ in range(s2["cooldown"]):
        check(r, ordered, spec)
    check(r, ordered, spec)                                     # cooled down: the sweep is back
    assert r.sweep_state()["n_ok"] == 1


def test_inputs_the_sweep_does_not_take(bev):
    """Channels outside {max_i, mean_z, density}, wide rasters, raw accumulators: plain two-pass, same numbers."""
    spec = spec_of(800, 1152, channels=(CH_MAX_I, CH_MEAN_I, CH_DENSITY))
    cloud = make_cloud(500_000, spec, seed=8, order="scan")
    r = bev.BevRasterizer(spec, len(cloud), algo="sweep")
    check(r, cloud, spec)
    assert r.sweep_state()["n_ok"] == 0 and r.sweep_state()["n_failed"] == 0
    wide = spec_of(600, 2400)
    cloud = make_cloud(500_000, wide, seed=8, order="scan")
    check(bev.BevRasterizer(wide, len(cloud), algo="sweep"), cloud, wide)


def test_sweep_edge_cases(bev):
    spec = spec_of(1152, 1152)
    r = bev.BevRasterizer(spec, 1_000_000, algo="sweep")
    # empty cloud: every cell is written (zero)
    out = r.alloc_outputs()
    out["image"].fill_(7)
    got = r(torch.empty((0, 4), dtype=torch.float32, device="cuda"), out=out)
    torch.cuda.synchronize()
    assert int(got["image"].max()) == 0 and r.sweep_state()["n_ok"] == 1
    # NaN / inf / out-of-grid points, a hot cell with > 4095 points (packed count wraps -> exact fall-back)
    cloud = make_cloud(300_000, spec, seed=3, order="scan")
    cloud[::7, 0] = np.nan
    cloud[5::11, 1] = np.inf
    cloud[3::13, 0] = -5.0
    hot = np.tile(np.array([[20.02, 30.01, 0.1, 5000.0]], dtype=np.float32), (5000, 1))
    srt = np.concatenate([cloud, hot])
    srt = srt[np.argsort(np.nan_to_num(srt[:, 0], nan=1e9, posinf=1e9), kind="stable")]
    check(r, srt, spec)
    assert r.sweep_state()["n_failed"] == 1                     # the 5000-point cell does not fit 12 bits
    # one point
    r2 = bev.BevRasterizer(spec, 10, algo="sweep")
    check(r2, np.array([[1.0, 2.0, 0.0, 900.0]], dtype=np.float32), spec)
