"""GPU, >= 2 devices: strip-sharded rasterisation over NCCL (plane-wise halo exchange, fused merge + re-finish,
mosaic gather) equals the ORACLE's one-piece raster bit for bit (and the single-GPU CUDA raster, and the
independent DIRECT-algorithm check of ``StripRasterizer.verify``).  Skipped on single-GPU boxes (the gloo
tests cover the host logic there; bench.py --gpus N runs the same verification on every scaling run)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z
    from lanemapping_b200.bev import BevRasterizer
    from lanemapping_b200.strips import StripRasterizer, coarse_strip_of
    from lanemapping_b200.synth import make_cloud
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        spec = BevSpec(2560, 1152, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), local_min_ele=-2.0)
        cloud = make_cloud(4_000_000, spec, seed=9, order="scan")
        sr = StripRasterizer(spec, len(cloud), halo=64)
        jitter = np.random.default_rng(3).integers(-60, 61, len(cloud)) * spec.img_reso[0]
        bucket = coarse_strip_of(cloud[:, 0] + jitter, spec, sr.plan.bounds)
        mine = torch.from_numpy(np.ascontiguousarray(cloud[bucket == rank])).cuda()
        strip = sr.rasterize(mine)
        mosaic = sr.gather(strip)
        one = BevRasterizer(spec, len(cloud), outputs=("image",))(torch.from_numpy(cloud).cuda())["image"]
        from oracle import c_oracle as CO
        want = torch.from_numpy(CO.rasterize(cloud, spec)["image"]).cuda()          # the oracle, not only CUDA vs CUDA
        ok_oracle = bool(torch.equal(mosaic, want)) and bool(torch.equal(one, want)) and sr.verify(mine, strip) == 0
        assert sr.planes == [0, 2, 3] and sr.halo_bytes == 3 * 64 * 1152 * 4       # count, sum_z, max_i only
        # pipelined form: three scenes in flight through the two buffer slots
        slots = [sr.step(mine) for _ in range(3)]
        piped = sr.mosaic(slots[-1]).clone()
        sr.flush()
        # mosaic gathered on one rank only, direct and through the pipelined form
        rooted = sr.gather(strip, root=1)
        ok_root = bool(torch.equal(rooted, one)) if rank == 1 else rooted is None
        sr2 = StripRasterizer(spec, len(cloud), halo=64, align=32, gather_root=0)
        mine2 = torch.from_numpy(np.ascontiguousarray(cloud[coarse_strip_of(cloud[:, 0] + jitter, spec, sr2.plan.bounds) == rank])).cuda()
        s2 = [sr2.step(mine2) for _ in range(3)]
        m2 = sr2.mosaic(s2[-1])
        a, b = sr2.plan.strip
        ok_root = ok_root and (bool(torch.equal(m2, one)) if rank == 0 else m2 is None) and bool(torch.equal(sr2.strip_of(s2[-1]), one[a:b]))
        sr2.flush()
        torch.cuda.synchronize()
        ok = bool(torch.equal(mosaic, one)) and bool(torch.equal(piped, one)) and slots == [0, 1, 0] and ok_root and ok_oracle
        q.put((rank, ok, int(mine.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_strips_equal_one_piece():
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == 4_000_000
