"""GPU tests of the callers either side of the hot path: the row-window loop for rasters with more
shared-memory tiles than one launch holds, the offline converter (LAS -> cropped_tiff PNG + sidecar),
and the on-the-fly DATASETS / PCENCODER plug-ins."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z, sidecar
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O
from oracle import inverse_oracle as INV

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    return B


@pytest.mark.parametrize("channels,count16", [((CH_MAX_I, CH_MEAN_Z, CH_DENSITY), False),
                                              ((CH_MAX_I, CH_MIN_Z, CH_MEAN_I, CH_MAX_Z), True)])
def test_row_window_loop_is_bit_exact(bev, channels, count16):
    """Force the in-call row-window loop (as config 4's 24 300 tiles would) on a small raster."""
    spec = BevSpec(700, 300, img_reso=(0.1, 0.1), channels=channels, count16=count16, local_min_ele=-2.0)
    cloud = make_cloud(600_000, spec, seed=31, order="scan")
    want = O.rasterize(cloud, spec)
    outs = ["image", "proj", "acc"] + (["count16"] if count16 else [])
    r = bev.BevRasterizer(spec, len(cloud), outputs=outs, tuning={"max_tiles": 7})   # 3 tiles per tile row -> 2 tile rows per window
    got = r(torch.from_numpy(cloud).cuda())
    torch.cuda.synchronize()
    st = r.stats()
    assert st["error"] == 0 and st["n_valid"] == int(O.accumulate(cloud, spec)[O.ACC_COUNT].sum())
    assert np.array_equal(got["image"].cpu().numpy(), want["image"])
    assert np.array_equal(got["proj"].cpu().numpy(), O.proj_from_image(want["image"]))
    assert np.array_equal(got["acc"].cpu().numpy().view(np.uint32), O.accumulate(cloud, spec))
    if count16:
        assert np.array_equal(got["count16"].cpu().numpy(), want["count16"])


def test_cfg4_geometry_slice(bev):
    """Config 4's geometry (0.02 m, 3456 columns, 4 x u8 + u16 count) on a 2304-row slice."""
    base = BevSpec(2304, 3456, img_reso=(0.02, 0.02), ele_reso=0.02)
    spec = BevSpec(2304, 3456, img_reso=(0.02, 0.02), ele_reso=0.02, local_min_ele=default_min_ele(base),
                   channels=(CH_MAX_I, CH_MIN_Z, CH_MEAN_I, CH_MAX_Z), count16=True)
    cloud = make_cloud(3_000_000, spec, order="scan")
    want = O.rasterize(cloud, spec)
    r = bev.BevRasterizer(spec, len(cloud), outputs=["image", "count16"])
    got = r(torch.from_numpy(cloud).cuda())
    torch.cuda.synchronize()
    r.check_device_errors()
    assert np.array_equal(got["image"].cpu().numpy(), want["image"])
    assert np.array_equal(got["count16"].cpu().numpy(), want["count16"])
    crops = bev.crop_tiles(got["image"], 1152)
    assert crops.shape == (2 * 3, 1152, 1152, 4)
    assert np.array_equal(crops.cpu().numpy(), O.crop_tiles(want["image"], 1152))


def test_offline_converter_writes_reference_compatible_files(bev, tmp_path):
    """LAS file -> rasterize_single_file -> PNG + sidecar that pass the reference's loader op
    sequence and are inverted by the reference's inverse map (restated in oracle/inverse_oracle.py,
    pinned against the reference's own outputs in tests/test_reference_contracts.py)."""
    from lanemapping_b200 import las
    from lanemapping_b200.convert_data import multiprocessing_las_files
    rng = np.random.default_rng(5)
    n = 400_000
    # a 60 m x 30 m patch of road in world coordinates (UTM-like magnitudes)
    world = np.stack([533000.0 + rng.random(n) * 60.0, 3380000.0 + rng.random(n) * 30.0,
                      21.0 + 0.01 * rng.standard_normal(n)], axis=1)
    inten = rng.integers(500, 40000, n)
    las_path = str(tmp_path / "181013_road.las")
    las.write_las(las_path, world, inten, offset=(533000.0, 3380000.0, 0.0))
    tiff, param = str(tmp_path / "cropped_tiff"), str(tmp_path / "cropped_tiff_param")
    cpts = str(tmp_path / "crop_points")
    stems = multiprocessing_las_files([las_path], tiff, param, num_process=2, crop_points_dir=cpts)
    assert stems == ["181013_0001", "181013_0002"] and all(len(s) == 11 for s in stems)   # 60 m -> 2 crops of 57.6 m
    assert multiprocessing_las_files([las_path], tiff, param, num_process=1) == stems       # resumable: manifest hit
    # the default decodes the LAS records on the GPU; the host (numpy) decode gives the same files
    tiff_h, param_h = str(tmp_path / "tiff_host"), str(tmp_path / "param_host")
    assert multiprocessing_las_files([las_path], tiff_h, param_h, num_process=1, las_decode="host") == stems
    for s_ in stems:
        assert open(os.path.join(tiff, s_ + ".png"), "rb").read() == open(os.path.join(tiff_h, s_ + ".png"), "rb").read()
        assert open(os.path.join(param, s_ + ".txt")).read() == open(os.path.join(param_h, s_ + ".txt")).read()
    xyz, inten_rd, hdr = las.read_las(las_path)
    for k, stem in enumerate(stems):
        img = np.array(Image.open(os.path.join(tiff, stem + ".png")), dtype=np.uint8)        # loader :87-88
        assert img.shape == (1152, 1152, 3)
        p = sidecar.read_sidecar(os.path.join(param, stem + ".txt"))
        assert p.coor_las_path == las_path and p.img_reso == (0.05, 0.05)
        # the crop equals the oracle's raster of the same points on the crop's own window
        spec = BevSpec(1152, 1152, bev_img_offset=(0.0, 0.0), local_min_ele=p.local_min_ele, ele_reso=p.ele_reso,
                       row0=k * 1152)
        local = sidecar.world_to_local(xyz, p)
        pts = np.concatenate([local, inten_rd[:, None]], axis=1).astype(np.float32)
        assert np.array_equal(img, O.rasterize(pts, spec)["image"])
        # inverse map (reference coor_img2pc.py:127-183) on an occupied pixel row recovers world x, y within a cell
        rows = np.arange(100, 1000, 100, dtype=np.float64) if k == 0 else np.arange(5, 45, 5, dtype=np.float64)
        seqs = np.zeros((1, len(rows), 2))
        seqs[0, :, 0], seqs[0, :, 1] = rows, 300.0
        params = {"img_reso": p.img_reso, "bev_img_offset": p.bev_img_offset, "ele_reso": p.ele_reso,
                  "local_min_ele": p.local_min_ele, "las_rotation_trans_quan": list(p.las_rotation_trans_quan),
                  "las_read_offset": list(p.las_read_offset)}
        back = INV.img2pc(params, seqs, [len(rows)], Image.open(os.path.join(tiff, stem + ".png")))
        assert np.allclose(back[0, :, 0], 533000.0 + k * 57.6 + rows * 0.05, atol=1e-6)
        assert np.allclose(back[0, :, 1], 3380000.0 + 300 * 0.05, atol=1e-6)
        assert np.all(np.abs(back[0, :, 2] - 21.0) <= p.ele_reso + 0.05)
    # the per-crop point records written for the on-the-fly dataset rasterise to the same crops
    from lanemapping_b200.pcencoder import BatchProjector
    loaded = [np.load(os.path.join(cpts, s + ".npz")) for s in stems]
    crop_pts = [torch.from_numpy(z["points"]).cuda() for z in loaded]
    assert sum(len(c) for c in crop_pts) == n
    proj = BatchProjector()(crop_pts, torch.from_numpy(np.stack([z["geom"] for z in loaded])))
    for k, s in enumerate(stems):
        png = np.array(Image.open(os.path.join(tiff, s + ".png")), dtype=np.uint8)
        assert np.array_equal(proj[k].cpu().numpy(), O.proj_from_image(png))      # bit-identical to the PNG path
    man = json.load(open(os.path.join(param, "181013_road.manifest.json")))       # keyed by the input's own stem
    assert man["stems"] == stems and man["n_points"] == n and man["source"] == las_path
    assert man["las_read_offset"] == [533000.0, 3380000.0, 20.0]                   # floor(min) of the data


def _write_road_las(path, x0, y0, z_of_x, n, seed, scale=0.001, offset=None, outlier=None, gap=None):
    from lanemapping_b200 import las
    rng = np.random.default_rng(seed)
    x = rng.random(n) * 100.0
    if gap is not None:                                       # keep points away from a crop edge (cell-exact tests)
        x = np.where(np.abs(x - gap) < 0.1, x + 1.0, x)
    world = np.stack([x0 + x, y0 + rng.random(n) * 20.0, z_of_x(x) + 0.01 * rng.standard_normal(n)], axis=1)
    if outlier is not None:
        world[0] = (x0 + 5.0, y0 + 5.0, outlier)             # in the first crop
    inten = rng.integers(500, 40000, n)
    las.write_las(path, world, inten, scale=(scale,) * 3, offset=offset)
    return las.read_las(path)


def test_offline_converter_files_sharing_six_digits_do_not_collide(bev, tmp_path):
    """'181013_0130.las' and '181013_0131.las' both suggest sequence id 181013: the driver
    gives them distinct ids, the manifests are per input file, and forcing the same id raises."""
    from lanemapping_b200.convert_data import multiprocessing_las_files, rasterize_single_file
    a, b = str(tmp_path / "181013_0130.las"), str(tmp_path / "181013_0131.las")
    _write_road_las(a, 1000.0, 2000.0, lambda x: 5.0 + 0 * x, 60_000, 1)
    _write_road_las(b, 5000.0, 2000.0, lambda x: 7.0 + 0 * x, 50_000, 2)
    tiff, param = str(tmp_path / "tiff"), str(tmp_path / "param")
    st = {}
    stems = multiprocessing_las_files([b, a], tiff, param, num_process=2, stats=st)
    assert stems == ["181013_0001", "181013_0002", "181014_0001", "181014_0002"]
    assert st["files"] == 2 and st["points"] == 110_000 and st["crops"] == 4 and st["files_per_s"] > 0
    assert 0.0 < st["png_share"] < 1.0 and st["devices"]
    ma = json.load(open(os.path.join(param, "181013_0130.manifest.json")))
    mb = json.load(open(os.path.join(param, "181013_0131.manifest.json")))
    assert ma["stems"] == stems[:2] and mb["stems"] == stems[2:] and ma["source"] == a and mb["source"] == b
    before = {s: os.path.getmtime(os.path.join(tiff, s + ".png")) for s in stems}
    assert multiprocessing_las_files([a, b], tiff, param, num_process=1) == stems          # both skipped via their manifests
    assert before == {s: os.path.getmtime(os.path.join(tiff, s + ".png")) for s in stems}
    with pytest.raises(FileExistsError):                     # the same id for another file: refuse, never overwrite
        rasterize_single_file(b, tiff, param, seq_id=181013)


def test_offline_converter_takes_read_offset_from_the_data(bev, tmp_path):
    """A LAS writer that leaves the header offset at 0 for UTM-scale coordinates: subtracting only the header
    offset would leave 3.38e6 m in float32 (ulp 0.25 m against 0.05 m cells).  The converter takes
    floor(min) of the data instead, writes it into the sidecar, and the raster equals the oracle's on the
    float64-exact local coordinates."""
    from lanemapping_b200.convert_data import multiprocessing_las_files
    path = str(tmp_path / "000042.las")
    xyz, inten, hdr = _write_road_las(path, 533100.0, 3380200.0, lambda x: 31.0 + 0.01 * x, 200_000, 3,
                                      scale=0.01, offset=(0.0, 0.0, 0.0))
    assert hdr.offset == (0.0, 0.0, 0.0)
    for mode, sub in (("gpu", "g"), ("host", "h")):
        tiff, param = str(tmp_path / ("tiff" + sub)), str(tmp_path / ("param" + sub))
        stems = multiprocessing_las_files([path], tiff, param, num_process=1, las_decode=mode, min_ele="file")
        assert stems == ["000042_0001", "000042_0002"]
        for k, stem in enumerate(stems):
            p = sidecar.read_sidecar(os.path.join(param, stem + ".txt"))
            assert tuple(p.las_read_offset) == tuple(np.floor(xyz.min(axis=0)))
            local = xyz - np.asarray(p.las_read_offset)
            pts = np.concatenate([local, inten[:, None]], axis=1).astype(np.float32)
            spec = BevSpec(1152, 1152, bev_img_offset=(0.0, 0.0), local_min_ele=p.local_min_ele, row0=k * 1152)
            img = np.array(Image.open(os.path.join(tiff, stem + ".png")), dtype=np.uint8)
            assert np.array_equal(img, O.rasterize(pts, spec)["image"])
            assert p.bev_img_offset == (k * 57.6, 0.0)


def test_offline_converter_per_crop_min_ele(bev, tmp_path):
    """A run that climbs 30 m over two crops, with one point 40 m below the road: one local_min_ele per FILE
    saturates the upper crop (12.75 m of u8 range at 0.05 m); the robust per-crop value keeps both crops in range,
    and the reference's inverse map recovers the road height from either crop's own sidecar."""
    from lanemapping_b200.convert_data import multiprocessing_las_files
    path = str(tmp_path / "000007.las")
    xyz, inten, _ = _write_road_las(path, 100.0, 200.0, lambda x: 50.0 + 30.0 * (x > 57.6), 300_000, 4, outlier=10.0, gap=57.6)
    res = {}
    for mode in ("file", "min", "robust"):
        tiff, param = str(tmp_path / ("tiff_" + mode)), str(tmp_path / ("param_" + mode))
        stems = multiprocessing_las_files([path], tiff, param, num_process=1, min_ele=mode)
        man = json.load(open(os.path.join(param, "000007.manifest.json")))
        res[mode] = ([sidecar.read_sidecar(os.path.join(param, s + ".txt")).local_min_ele for s in stems],
                     [man["elevation_saturated_fraction"][s] for s in stems])
        if mode == "robust":
            for k, stem in enumerate(stems):
                p = sidecar.read_sidecar(os.path.join(param, stem + ".txt"))
                img = np.array(Image.open(os.path.join(tiff, stem + ".png")), dtype=np.uint8)
                occ = img[..., 2] > 0
                z_back = img[..., 1][occ].astype(np.float64) * p.ele_reso + p.local_min_ele + p.las_read_offset[2]
                assert np.abs(np.median(z_back) - (50.0 + 30.0 * k)) < 0.1
    # local frame: las_read_offset z = floor(10.0) = 10 -> outlier at 0, road at 40 (crop 0) and 70 (crop 1)
    assert res["file"][0] == [0.0, 0.0] and min(res["file"][1]) > 0.99            # the outlier drags the whole file down
    assert res["min"][0][0] == 0.0 and abs(res["min"][0][1] - 69.9) < 0.11        # crop 0 still follows its outlier
    assert res["min"][1][0] > 0.99 and res["min"][1][1] == 0.0
    assert res["robust"][1] == [0.0, 0.0]
    assert abs(res["robust"][0][0] - 39.4) < 0.11 and abs(res["robust"][0][1] - 69.4) < 0.11


def test_on_the_fly_projector_matches_png_loader_path(bev, tmp_path):
    """sample['points'] -> BatchProjector -> sample['proj'] equals to_tensor(PNG).float() of the
    same crop (reference baseline/datasets/laserlane_proposals.py:87-89), for a batch of 3."""
    import torchvision
    from lanemapping_b200.convert_data import _write_png
    from lanemapping_b200.datasets import CropPoints, collate_points
    from lanemapping_b200.pcencoder import OnTheFlyProjector
    root = tmp_path
    for d in ("crop_points", "cropped_tiff_param", "cropped_tiff"):
        os.makedirs(root / d)
    stems = ["000000_0001", "000000_0002", "000000_0003"]
    want = []
    for i, stem in enumerate(stems):
        spec = BevSpec(1152, 1152, bev_img_offset=(10.0 * i, -5.0), local_min_ele=default_min_ele(BevSpec(1152, 1152)))
        cloud = make_cloud(200_000 + 1000 * i, spec, seed=i, order="scan")
        if i == 0:
            np.save(root / "crop_points" / (stem + ".npy"), cloud)           # plain records + sidecar also work
        else:
            np.savez(root / "crop_points" / (stem + ".npz"), points=cloud,
                     geom=np.array([*spec.bev_img_offset, *spec.img_reso, spec.local_min_ele, spec.ele_reso, 0, 0]))
        sidecar.write_sidecar(str(root / "cropped_tiff_param" / (stem + ".txt")),
                              sidecar.PcImgParams("x.las", (0.0, 0.0, 0.0), (0, 0, 0, 1, 0, 0, 0), spec.bev_img_offset,
                                                  spec.img_reso, spec.local_min_ele, spec.ele_reso))
        img = O.rasterize(cloud, spec)["image"]
        _write_png(str(root / "cropped_tiff" / (stem + ".png")), img)
        png = np.array(Image.open(root / "cropped_tiff" / (stem + ".png")), dtype=np.uint8)
        want.append(torchvision.transforms.functional.to_tensor(png).float())
    with open(root / "split.json", "w") as f:
        json.dump({"train": stems, "test": stems, "valid": stems, "single": stems[:1], "pretrain": stems}, f)
    ds = CropPoints(str(root), "split.json", "test")
    assert len(ds) == 3 and ds[1]["image_name"] == "000000_0002"
    batch = collate_points([ds[i] for i in range(3)])
    assert len(batch["points"]) == 3 and batch["bev_geom"].shape == (3, 8)
    batch["points"] = batch["points"].cuda(non_blocking=True)    # Runner.to_cuda's non-list branch (tests/test_plugins.py)

    class Inner(torch.nn.Module):                                 # stands in for PostProjector2.forward
        def forward(self, sample):
            return sample["proj"].mean(dim=(2, 3))

    enc = OnTheFlyProjector(Inner())
    out = enc(batch)
    torch.cuda.synchronize()
    assert batch["proj"].shape == (3, 3, 1152, 1152) and batch["proj"].dtype == torch.float32
    assert torch.equal(batch["proj"].cpu(), torch.stack(want))
    assert out.shape == (3, 3)
