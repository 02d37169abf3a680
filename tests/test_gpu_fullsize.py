"""GPU parity at BASELINE.json's FULL sizes (configs[1] and configs[3]: 10^8 points each).

The numpy oracle needs minutes for 10^8 points on one core, so the whole raster is compared bit for
bit with the plain-C restatement (oracle/bev_oracle.c, seconds; tests/test_c_oracle.py holds the two
restatements equal).  Beside that, size-independent properties: conservation of the in-grid point
count (kernel statistics == the numpy oracle's recount of the keys == the C oracle's) and agreement
of the two independent CUDA algorithms (binned pipeline == global-atomics path)."""
import os

import numpy as np
import pytest
import torch

from lanemapping_b200.synth import config_spec, make_cloud
from oracle import bev_oracle as O
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu

FULL = os.environ.get("LM_SKIP_FULLSIZE", "0") != "1"


def _host_valid_count(cloud, spec, chunk=1 << 23):
    n = 0
    for lo in range(0, len(cloud), chunk):
        row, col, _, _, keep = O.quantise_points(cloud[lo:lo + chunk], spec)
        n += int(keep.sum())
    return n


def _run(bev, pts, spec, algo, outputs, n):
    r = bev.BevRasterizer(spec, n, algo=algo, outputs=outputs)
    out = r(pts)
    torch.cuda.synchronize()
    st = r.stats()
    assert st["error"] == 0, st
    return {k: v.cpu().numpy() for k, v in out.items()}, st


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    assert torch.cuda.is_available()
    return B


@pytest.mark.skipif(not FULL, reason="LM_SKIP_FULLSIZE=1")
@pytest.mark.parametrize("cfg", [2, 4])
def test_full_size_config_is_bit_exact(bev, cfg):
    spec, n = config_spec(cfg)
    cloud = make_cloud(n, spec, order="scan")
    want = C.rasterize(cloud, spec)
    n_valid = _host_valid_count(cloud, spec)
    assert want["n_valid"] == n_valid
    pts = torch.from_numpy(cloud).cuda()
    outputs = ("image", "count16") if spec.count16 else ("image",)
    got, st = _run(bev, pts, spec, "binned", outputs, n)
    assert st["n_valid"] == n_valid
    assert np.array_equal(got["image"], want["image"]), "full-size u8 raster differs from the oracle"
    if spec.count16:
        assert np.array_equal(got["count16"], want["count16"])
        # every in-grid point is counted exactly once (no cell of this cloud reaches 65535)
        assert int(want["count16"].max()) < 65535 and int(got["count16"].sum(dtype=np.int64)) == n_valid
    direct, st2 = _run(bev, pts, spec, "direct", outputs, n)
    assert st2["n_valid"] == n_valid
    for k in outputs:
        assert np.array_equal(direct[k], got[k]), f"binned and direct differ in {k}"
