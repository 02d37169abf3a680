"""CPU tests of host-side pieces that need no GPU: LAS reader, sidecar transform, split parsing."""
import json

import numpy as np
import pytest

from lanemapping_b200 import las, sidecar
from lanemapping_b200.datasets import read_split, collate_points


def test_las_roundtrip_and_header(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.random((5000, 3)) * [100, 50, 5] + [533000.0, 3380000.0, 20.0]
    inten = rng.integers(0, 65536, 5000)
    path = str(tmp_path / "a.las")
    las.write_las(path, xyz, inten, scale=(0.001, 0.001, 0.001))
    got, gi, hdr = las.read_las(path)
    assert hdr.n_points == 5000 and hdr.record_length == 20 and hdr.version == (1, 2)
    assert np.abs(got - xyz).max() <= 0.0005 + 1e-9 and np.array_equal(gi, inten)
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        las.read_las(path)


def test_world_local_roundtrip_with_rotation():
    ang = np.deg2rad(40.0)
    p = sidecar.PcImgParams("a.las", (5e5, 3e6, 10.0), (3.0, -2.0, 1.0, np.cos(ang / 2), 0, 0, np.sin(ang / 2)),
                            (0.0, 0.0), (0.05, 0.05), -1.0, 0.05)
    local = np.array([[1.0, 0.0, 0.5], [10.0, -3.0, 0.1]])
    world = sidecar.local_to_world(local, p)
    # a +40 degree rotation about z of (1,0,0), then the two translations
    assert np.allclose(world[0], [np.cos(ang) + 3.0 + 5e5, np.sin(ang) - 2.0 + 3e6, 0.5 + 1.0 + 10.0])
    assert np.allclose(sidecar.world_to_local(world, p), local, atol=1e-8)
    with pytest.raises(ValueError):
        sidecar.format_sidecar(sidecar.PcImgParams("a", (0, 0), (0,) * 7, (0, 0), (1, 1), 0, 1))


def test_split_modes(tmp_path):
    stems = ["%06d_%04d" % (1, i) for i in range(200)]
    with open(tmp_path / "s.json", "w") as f:
        json.dump({"train": stems[:100], "test": stems[100:150], "valid": stems, "single": stems[:1], "pretrain": stems}, f)
    assert len(read_split(str(tmp_path), "s.json", "train")) == 100
    assert len(read_split(str(tmp_path), "s.json", "valid")) == 150        # reference caps valid at 150
    assert read_split(str(tmp_path), "s.json", "all") == stems
    with pytest.raises(AssertionError):
        read_split(str(tmp_path), "s.json", "val")                           # reference asserts on unknown modes


def test_seq_ids_are_unique_per_run_and_read_offset_comes_from_the_data(tmp_path):
    from lanemapping_b200.convert_data import assign_seq_ids, data_read_offset, seq_id_of
    files = ["/d/181013_0131.las", "/d/181013_0130.las", "/d/2021-03-15_run1.las", "/d/2021-03-15_run2.las", "/d/x.las"]
    assert seq_id_of(files[0]) == 181013 == seq_id_of(files[1])               # the suggestion collides ...
    ids = assign_seq_ids(files)
    assert len(set(ids.values())) == len(files)                                # ... the assignment does not
    assert ids["/d/181013_0130.las"] == 181013 and ids["/d/181013_0131.las"] == 181014
    assert ids["/d/2021-03-15_run1.las"] == 202103 and ids["/d/2021-03-15_run2.las"] == 202104 and ids["/d/x.las"] == 0
    assert assign_seq_ids(list(reversed(files))) == ids                        # deterministic in the file list
    assert data_read_offset((533100.37, 3380200.9, -4.2)) == (533100.0, 3380200.0, -5.0)
    # header bounds: written by write_las, parsed by read_header; a header without bounds is flagged
    xyz = np.array([[10.5, 20.25, 3.0], [11.5, 19.0, 4.0], [12.0, 21.0, 2.5]])
    path = str(tmp_path / "b.las")
    las.write_las(path, xyz, np.arange(3))
    raw, hdr = las.read_point_block(path)
    assert hdr.bounds_ok() and hdr.mins == (10.5, 19.0, 2.5) and hdr.maxs == (12.0, 21.0, 4.0)
    assert las.world_min(raw, hdr) == (10.5, 19.0, 2.5)
    blank = bytearray(open(path, "rb").read())
    blank[179:227] = bytes(48)
    open(path, "wb").write(bytes(blank))
    raw2, hdr2 = las.read_point_block(path)
    assert not hdr2.bounds_ok() and las.world_min(raw2, hdr2) == (10.5, 19.0, 2.5)


def test_png_writer_is_lossless_in_pil_channel_order_and_pool_propagates_errors(tmp_path):
    """The converter's PNG writer (zlib level 1 + RLE strategy, crops of one file encoded by a shared thread pool):
    what PIL reads back -- the loader's view, reference baseline/datasets/laserlane_proposals.py:88-89 -- is the
    raster byte for byte, channel c of the raster in channel c of the PIL image; a failing write raises in the caller."""
    from PIL import Image
    from lanemapping_b200 import convert_data as CD
    rng = np.random.default_rng(3)
    imgs = {}
    for c in (3, 4):
        img = np.zeros((96, 160, c), dtype=np.uint8)
        occ = rng.random((96, 160)) < 0.7
        img[occ] = rng.integers(0, 256, (int(occ.sum()), c), dtype=np.uint8)
        imgs[c] = img
    pool = CD._encode_pool()
    assert pool is CD._encode_pool()                       # one pool per process
    futs = [pool.submit(CD._write_png, str(tmp_path / f"c{c}_{i}.png"), imgs[c]) for c in (3, 4) for i in range(4)]
    for f in futs:
        f.result()
    for c in (3, 4):
        for i in range(4):
            back = np.asarray(Image.open(tmp_path / f"c{c}_{i}.png"))
            assert back.shape == imgs[c].shape and np.array_equal(back, imgs[c])
    bad = pool.submit(CD._write_png, str(tmp_path / "no_such_dir" / "x.png"), imgs[3])
    with pytest.raises(Exception):
        bad.result()
    CD.shutdown_encoders()                                 # (later tests fork: leave no threads behind)
    assert CD._encoders is None
