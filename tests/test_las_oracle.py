"""CPU: the LAS decode oracle against a hand-computed record, against the independent reader in
lanemapping_b200/las.py, and against the sidecar's float64 world<->local maps."""
import struct

import numpy as np

from lanemapping_b200 import las, sidecar
from oracle import las_oracle as L


def test_known_record():
    # X=123456, Y=-7, Z=2^31-1, intensity 65535, then 6 bytes of other fields
    rec = struct.pack("<iiiH", 123456, -7, 2**31 - 1, 65535) + bytes([9, 2, 0, 0, 1, 0])
    got = L.decode_records(np.frombuffer(rec, np.uint8), 20, (0.001, 0.01, 0.0001), (500000.0, 3000000.0, -10.0))
    want = [np.float32(123456 * 0.001 + 500000.0), np.float32(-7 * 0.01 + 3000000.0),
            np.float32((2**31 - 1) * 0.0001 - 10.0), np.float32(65535)]
    assert got.shape == (1, 4) and [got[0, k] for k in range(4)] == want


def test_matches_las_reader_and_sidecar_maps(tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.random((4000, 3)) * [120, 60, 6] + [533000.0, 3380000.0, 20.0]
    inten = rng.integers(0, 65536, 4000)
    path = str(tmp_path / "a.las")
    las.write_las(path, xyz, inten)
    world, gi, hdr = las.read_las(path)
    raw, hdr2 = las.read_point_block(path)
    assert hdr2 == hdr and raw.size == 4000 * 20
    # world frame: float32 of the reader's float64 coordinates, bit for bit
    got = L.decode_records(raw, hdr.record_length, hdr.scale, hdr.offset)
    assert np.array_equal(got[:, :3], world.astype(np.float32)) and np.array_equal(got[:, 3], gi.astype(np.float32))
    # local frame of a crop: agrees with sidecar.world_to_local (a BLAS product) to float32 resolution
    ang = np.deg2rad(33.0)
    p = sidecar.PcImgParams(path, (533000.0, 3380000.0, 20.0), (3.0, -2.0, 1.0, np.cos(ang / 2), 0, 0, np.sin(ang / 2)),
                            (0.0, 0.0), (0.05, 0.05), -1.0, 0.05)
    rot = sidecar.quat_to_matrix(p.las_rotation_trans_quan[3:]).T.reshape(9)
    loc = L.decode_records(raw, 20, hdr.scale, hdr.offset, p.las_read_offset, p.las_rotation_trans_quan[:3], rot)
    ref = sidecar.world_to_local(world, p)
    assert np.abs(loc[:, :3] - ref).max() < 2e-5


def test_record_lengths_and_misaligned_fields():
    rng = np.random.default_rng(2)
    for reclen in (14, 15, 20, 26, 28, 34, 37, 100):
        n = 50
        rec = rng.integers(0, 256, (n, reclen), dtype=np.uint8)
        X, Y, Z, I = L.split_records(rec.reshape(-1), reclen)
        for i in (0, 17, n - 1):
            x, y, z, it = struct.unpack_from("<iiiH", rec[i].tobytes(), 0)
            assert (X[i], Y[i], Z[i], I[i]) == (x, y, z, it)
