"""CPU, world_size 2 and 3 over gloo: strip-sharded rasterisation == one-piece rasterisation,
bit for bit, after the halo merge and the mosaic gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z
from lanemapping_b200.strips import StripRasterizer, coarse_strip_of, make_plan, strip_bounds
from lanemapping_b200.synth import make_cloud


def test_strip_bounds_are_aligned_and_cover():
    b = strip_bounds(11520, 8, 128)
    assert b[0][0] == 0 and b[-1][1] == 11520 and all(b[k][1] == b[k + 1][0] for k in range(7))
    assert all(e % 128 == 0 for _, e in b[:-1]) and max(r1 - r0 for r0, r1 in b) - min(r1 - r0 for r0, r1 in b) <= 128
    assert strip_bounds(1000, 1) == [(0, 1000)]
    assert strip_bounds(300, 2, 128) == [(0, 128), (128, 300)]
    p = make_plan(BevSpec(11520, 1152), 3, 8, 64)
    assert (p.win0, p.win1) == (p.strip[0] - 64, p.strip[1] + 64) and p.top == p.bottom == 64
    p0 = make_plan(BevSpec(11520, 1152), 0, 8, 64)
    assert p0.top == 0 and p0.win0 == 0
    with pytest.raises(ValueError):
        make_plan(BevSpec(256, 64), 0, 2, 200)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, halo, q):
    from oracle_backend import OracleBackend
    from oracle import bev_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = BevSpec(640, 96, img_reso=(0.25, 0.25), ele_reso=0.05, local_min_ele=-2.0,
                       channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
        cloud = make_cloud(60_000, spec, seed=5, order="scan")
        sr = StripRasterizer(spec, len(cloud), halo=halo, backend=OracleBackend(), device="cpu", align=32)
        # coarse bucketing with a deliberately sloppy along-track key: +-(halo-1) rows of error
        jitter = np.random.default_rng(17).integers(-(halo - 1), halo, len(cloud)) * spec.img_reso[0] if halo > 1 else 0.0
        bucket = coarse_strip_of(cloud[:, 0] + jitter, spec, sr.plan.bounds)
        mine = torch.from_numpy(np.ascontiguousarray(cloud[bucket == rank]))
        strip = sr.rasterize(mine)
        mosaic = sr.gather(strip)
        want = O.rasterize(cloud, spec)["image"]
        r0, r1 = sr.plan.strip
        ok_strip = np.array_equal(strip.numpy(), want[r0:r1])
        ok_mosaic = np.array_equal(mosaic.numpy(), want)
        # gather on one rank only (what an offline writer needs): the others send and get None
        root = world - 1
        rooted = sr.gather(strip, root=root)
        ok_mosaic = ok_mosaic and (np.array_equal(rooted.numpy(), want) if rank == root else rooted is None)
        q.put((rank, ok_strip, ok_mosaic, int(mine.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, 24), (3, 16)])
def test_strips_equal_one_piece(world, halo):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), f"strip mismatch: {res}"
    assert all(r[2] for r in res), f"mosaic mismatch: {res}"
    assert sum(r[3] for r in res) == 60_000


def test_single_rank_is_plain_rasterisation():
    from oracle_backend import OracleBackend
    from oracle import bev_oracle as O
    spec = BevSpec(100, 40, img_reso=(0.5, 0.5), local_min_ele=-2.0)
    cloud = make_cloud(5000, spec, seed=1)
    sr = StripRasterizer(spec, len(cloud), halo=8, backend=OracleBackend(), device="cpu")
    strip = sr.rasterize(torch.from_numpy(cloud))
    assert np.array_equal(sr.gather(strip).numpy(), O.rasterize(cloud, spec)["image"])
