"""GPU parity of the LAS front end (include/lm_las.h): decode == oracle bit for bit for every
record length / alignment, and the fused LAS -> BEV call == decode + rasterise == oracle."""
import numpy as np
import pytest
import torch

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z, _cabi, las, sidecar
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O
from oracle import las_oracle as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    assert torch.cuda.is_available()
    return B


def random_records(n, reclen, seed):
    """Random bytes everywhere (every int32 X/Y/Z, every intensity, arbitrary trailing fields)."""
    return np.random.default_rng(seed).integers(0, 256, n * reclen, dtype=np.uint8)


ROT = sidecar.quat_to_matrix((np.cos(0.3), 0.1 * np.sin(0.3), -0.2 * np.sin(0.3), np.sqrt(0.95) * np.sin(0.3))).T.reshape(9)


@pytest.mark.parametrize("reclen", [14, 15, 20, 26, 28, 34, 37, 67, 100])
@pytest.mark.parametrize("n", [1, 1023, 1024, 1025, 70_001])
def test_decode_matches_oracle(bev, reclen, n):
    raw = random_records(n, reclen, seed=reclen * 131 + n)
    args = ((0.001, 0.002, 0.0005), (533000.0, 3380000.0, 20.0), (533010.5, 3380020.25, 19.0), (3.0, -2.0, 1.0), ROT)
    x = _cabi.make_las_xform(reclen, *args)
    got = bev.decode_las(torch.from_numpy(raw).cuda(), n, x).cpu().numpy()
    want = L.decode_records(raw, reclen, *args)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_decode_identity_is_the_reader(bev, tmp_path):
    rng = np.random.default_rng(5)
    xyz = rng.random((30_000, 3)) * [120, 60, 6] + [533000.0, 3380000.0, 20.0]
    path = str(tmp_path / "a.las")
    las.write_las(path, xyz, rng.integers(0, 65536, 30_000))
    world, inten, hdr = las.read_las(path)
    raw, _ = las.read_point_block(path)
    got = bev.decode_las(torch.from_numpy(raw).cuda(), hdr.n_points, bev.las_xform(hdr)).cpu().numpy()
    assert np.array_equal(got[:, :3], world.astype(np.float32)) and np.array_equal(got[:, 3], inten.astype(np.float32))


def las_scene(tmp_path, n, spec, reclen_pad=0, order="scan"):
    """A synthetic road cloud written as a LAS file in a rotated, offset world frame."""
    cloud = make_cloud(n, spec, order=order)
    ang = np.deg2rad(25.0)
    p = sidecar.PcImgParams("scene.las", (533000.0, 3380000.0, 20.0),
                            (1.5, -0.5, 0.25, np.cos(ang / 2), 0.0, 0.0, np.sin(ang / 2)),
                            spec.bev_img_offset, spec.img_reso, spec.local_min_ele, spec.ele_reso)
    world = sidecar.local_to_world(cloud[:, :3].astype(np.float64), p)
    path = str(tmp_path / "scene.las")
    las.write_las(path, world, cloud[:, 3].astype(np.uint16), scale=(0.0005, 0.0005, 0.0005))
    raw, hdr = las.read_point_block(path)
    if reclen_pad:                                   # longer records (other point formats): pad every record
        rec = raw.reshape(-1, hdr.record_length)
        rec = np.concatenate([rec, np.full((len(rec), reclen_pad), 0xA5, np.uint8)], axis=1)
        raw = np.ascontiguousarray(rec).reshape(-1)
        hdr = las.LasHeader(hdr.version, hdr.offset_to_points, 1, hdr.record_length + reclen_pad, hdr.n_points,
                            hdr.scale, hdr.offset)
    return raw, hdr, p


@pytest.mark.parametrize("pad,n", [(0, 1_500_000), (8, 300_000), (14, 300_001), (3, 2047)])
def test_fused_las_raster_matches_decode_then_raster_and_oracle(bev, tmp_path, pad, n):
    spec = BevSpec(2304, 1152, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), local_min_ele=default_min_ele(BevSpec(2304, 1152)))
    raw, hdr, p = las_scene(tmp_path, n, spec, pad)
    x = bev.las_xform(hdr, p)
    rot = sidecar.quat_to_matrix(p.las_rotation_trans_quan[3:]).T.reshape(9)
    pts = L.decode_records(raw, hdr.record_length, hdr.scale, hdr.offset, p.las_read_offset,
                           p.las_rotation_trans_quan[:3], rot)
    want = O.rasterize(pts, spec)["image"]
    assert (want[..., 2] > 0).sum() > min(n, 100_000) // 10           # the scene really lands in the window
    dev = torch.from_numpy(raw).cuda()
    r = bev.BevRasterizer(spec, n, outputs=("image", "proj"))
    fused = r.rasterize_las(dev, n, x)
    torch.cuda.synchronize()
    assert r.stats()["error"] == 0
    assert np.array_equal(fused["image"].cpu().numpy(), want)
    assert np.array_equal(fused["proj"].cpu().numpy(), O.proj_from_image(want))
    two_step = r(bev.decode_las(dev, n, x))
    assert torch.equal(two_step["image"], fused["image"])


def test_fused_las_errors(bev):
    spec = BevSpec(1152, 1152)
    r = bev.BevRasterizer(spec, 1000)
    x = _cabi.make_las_xform(20, (0.001,) * 3, (0.0,) * 3)
    with pytest.raises(ValueError):
        r.rasterize_las(torch.zeros(100, dtype=torch.uint8, device="cuda"), 1000, x)      # too few bytes
    with pytest.raises(ValueError):
        bev.decode_las(torch.zeros(100, dtype=torch.uint8, device="cuda"), 1000, x)
