"""Regenerates tests/golden/* by running the REFERENCE's own code (only possible where
/root/reference is mounted; the fixtures it writes are committed and travel to the GPU box).

    python tests/golden/make_golden.py

What is pinned (SURVEY.md section 8c -- the reference has no rasteriser, so these are the
contracts its in-tree code does fix):
  golden_crop.png / golden_crop.txt   a crop + sidecar written by OUR writer from an oracle raster
  sidecar_parsed.json                 reference io_utils.load_pc_2_img_transform_paras on that sidecar
  inverse_io.json                     reference coor_img2pc.transform_coordinate_from_img_2_pc on
                                      polylines over that crop (inputs and outputs)
  loader_contract.json                the loader's op sequence (laserlane_proposals.py:87-94) on the PNG
  label_frame.json                    reference data/convert_data.py frame facts (1152 hard-coded tile,
                                      NpEncoder) used by the naming tests
  inverse_io2.json                    a harder inverse case: vertices inside / on the rim of empty holes
                                      (in-place hole filling that later vertices see), a non-unit quaternion
  inverse_io3.npz                     random crops / polylines / poses through the same reference function
  labels_seq.json                     the sequence JSON the same call writes (save_seq)
  labels_in.json, labels_*.png        reference data/convert_data.write_instance_orientation_seq run on
                                      synthetic polylines: the four label rasters it writes
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "baseline", "utils"))


def golden_inputs():
    """Deterministic small scene: a crowned road (z depends on y only) with a painted stripe,
    rasterised on a 128 x 128 crop with a rotated + translated sidecar pose."""
    from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z
    from lanemapping_b200.sidecar import PcImgParams
    rng = np.random.default_rng(20211013)
    n = 200000
    spec = BevSpec(128, 128, bev_img_offset=(10.0, -4.0), img_reso=(0.05, 0.05), local_min_ele=-1.0, ele_reso=0.05,
                   channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    x = 10.0 + rng.random(n) * 6.4
    y = -4.0 + rng.random(n) * 6.4
    keep = ~((x > 12.0) & (x < 13.0) & (y > -2.0) & (y < -1.0))      # a 20 x 20 px hole: empty cells
    x, y = x[keep], y[keep]
    z = 0.25 * np.round((y + 4.0) / 1.6)                              # terraces: exact multiples of ele_reso
    inten = np.where(np.abs(y + 0.8) < 0.075, 30000.0, 5000.0) + rng.integers(0, 50, len(x))
    local = np.stack([x, y, z], axis=1)
    ang = np.deg2rad(33.0)
    pose = (120.5, -40.25, 3.0, float(np.cos(ang / 2)), 0.0, 0.0, float(np.sin(ang / 2)))
    params = PcImgParams("/data/LaserLane/las/181013.las", (533000.0, 3380000.0, 20.0), pose,
                         spec.bev_img_offset, spec.img_reso, spec.local_min_ele, spec.ele_reso)
    pts = np.concatenate([local, inten[:, None]], axis=1).astype(np.float32)
    return spec, params, pts


def polylines():
    # rows increase along the line (lanes run low-row -> high-row, reference data/convert_data.py:151-156)
    rows = np.arange(4, 124, 8, dtype=np.float64)
    seqs = np.zeros((3, len(rows), 2))
    lens = [len(rows), len(rows) - 3, len(rows)]
    seqs[0, :, 0], seqs[0, :, 1] = rows, 64.0          # along the painted stripe
    seqs[1, :lens[1], 0], seqs[1, :lens[1], 1] = rows[:lens[1]], 20.0
    seqs[2, :, 0], seqs[2, :, 1] = rows, 50.0          # crosses the empty hole (rows 40-59, cols 40-59)
    return seqs, lens


def label_inputs():
    """Polylines (row, col) of a synthetic label crop: crossing lanes (the later one must win), every
    octant of segment direction incl. right-to-left ones, a degenerate lane, end points inside and
    outside the 20 px clip border."""
    lanes = [
        [(30.0, 200.5), (300.2, 230.7), (700.9, 260.1), (1100.4, 300.8)],           # steep, down the image
        [(1120.0, 640.0), (800.5, 600.2), (400.1, 655.9), (25.7, 610.3)],           # drawn bottom-up
        [(500.0, 100.0), (520.3, 500.6), (480.8, 900.2), (510.0, 1130.5)],          # shallow: crosses the others
        [(10.0, 900.0), (400.0, 905.0), (1145.0, 1140.0)],                           # both end points in the border
        [(600.0, 50.0), (600.0, 50.0)],                                              # degenerate: no lane instance
        [(200.0, 1000.0), (200.0, 700.0), (640.0, 700.0), (640.0, 1000.0), (300.0, 1000.0)],   # axis-aligned loop, leftwards runs
    ]
    n, m = len(lanes), max(len(l) for l in lanes)
    seqs = np.zeros((n, m, 2))
    lens = []
    for i, l in enumerate(lanes):
        seqs[i, :len(l)] = l
        lens.append(len(l))
    semantic = [1, 2, 1, 3, 1, 2]
    instance = [1, 2, 3, 4, 5, 6]
    return seqs, lens, semantic, instance


def make_labels(ref_convert):
    import tempfile
    import cv2
    seqs, lens, semantic, instance = label_inputs()
    orient = ref_convert.cal_seq_orientation(seqs.copy(), lens)          # reference :72-104
    with tempfile.TemporaryDirectory() as d:
        names = [os.path.join(d, k) for k in ("seq.json", "sem.png", "ins.png", "ori.png", "endp.png")]
        ref_convert.write_instance_orientation_seq(seqs.copy(), lens, semantic, instance, orient, *names)   # :319-369
        import shutil
        shutil.copy(names[0], os.path.join(HERE, "labels_seq.json"))                 # the reference's save_seq output
        for k, src in zip(("semantic", "instance", "orient", "endp"), names[1:]):
            img = cv2.imread(src, cv2.IMREAD_UNCHANGED)
            assert img.dtype == np.uint8 and img.shape == (1152, 1152), (img.dtype, img.shape)
            cv2.imwrite(os.path.join(HERE, f"labels_{k}.png"), img, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    with open(os.path.join(HERE, "labels_in.json"), "w") as f:
        json.dump({"seqs": seqs.tolist(), "lens": lens, "semantic": semantic, "instance": instance,
                   "orient": orient.tolist()}, f)


def make_inverse2(io_utils, coor_img2pc, ref_convert):
    """Second inverse case on the golden crop: polylines through the empty hole (rows/cols 40-59) so that
    several vertices are filled in place and later searches see the earlier fills; a vertex at (0,0);
    a non-unit quaternion (the reference divides the conjugate by |q|, not |q|^2)."""
    from PIL import Image
    parsed = io_utils.load_pc_2_img_transform_paras(os.path.join(HERE, "golden_crop.txt"))
    parsed["las_rotation_trans_quan"] = [1.5, -2.25, 0.5, 0.9, 0.1, -0.2, 0.45]
    seqs = np.zeros((4, 12, 2))
    lens = [12, 9, 12, 5]
    seqs[0, :, 0], seqs[0, :, 1] = np.arange(38, 62, 2), np.arange(38, 62, 2) + 0.5      # diagonal through the hole
    seqs[1, :9, 0], seqs[1, :9, 1] = 49.7, np.arange(36, 63, 3)                          # along a row inside it
    seqs[2, :, 0], seqs[2, :, 1] = np.arange(44, 56, 1), 50.2                            # dense: neighbours get filled
    seqs[3, :5, 0], seqs[3, :5, 1] = [0.0, 3.0, 50.0, 120.0, 127.0], [0.0, 3.0, 50.0, 5.0, 127.0]
    img = Image.open(os.path.join(HERE, "golden_crop.png"))
    world = coor_img2pc.transform_coordinate_from_img_2_pc(parsed, seqs.copy(), lens, img)
    filled = coor_img2pc.modify_empty_pixel_elevation(np.array(img), seqs.copy(), lens)
    changed = np.argwhere(filled[:, :, 1] != np.array(img)[:, :, 1])
    with open(os.path.join(HERE, "inverse_io2.json"), "w") as f:
        json.dump({"params": parsed, "img_seqs": seqs.tolist(), "img_seq_lens": lens, "world": world.tolist(),
                   "filled_px": [[int(r), int(c), int(filled[r, c, 1])] for r, c in changed]}, f, cls=ref_convert.NpEncoder)


def make_inverse3(coor_img2pc):
    """Random crops, random polylines, random poses through the reference: pins the summation order of
    LeastSuqare (builtin sum) and the hole filling on arbitrary data.  Stored as a compressed npz."""
    from PIL import Image
    rng = np.random.default_rng(3)
    B, H, W, L, V = 3, 96, 80, 6, 17
    images = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    images[rng.random((B, H, W)) < 0.6] = 0
    images[:, 10:40, 20:50] = 0
    seqs = np.zeros((B, L, V, 2))
    lens = rng.integers(1, V + 1, (B, L))
    lens[:, 0] = V
    poses = np.zeros((B, 13))
    world = np.zeros((B, L, V, 3))
    for b in range(B):
        for l in range(L):
            seqs[b, l, :lens[b, l], 0] = rng.uniform(0, H - 1e-6, lens[b, l])
            seqs[b, l, :lens[b, l], 1] = rng.uniform(0, W - 1e-6, lens[b, l])
        q = rng.normal(size=4)
        par = {"img_reso": [0.05, 0.04], "bev_img_offset": [float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))],
               "ele_reso": 0.05, "local_min_ele": float(rng.uniform(-3, 3)),
               "las_rotation_trans_quan": [*rng.uniform(-10, 10, 3).tolist(), *(q / np.linalg.norm(q)).tolist()],
               "las_read_offset": [533000.0, 3380000.0, 20.0]}
        poses[b] = [*par["bev_img_offset"], par["local_min_ele"], *par["las_rotation_trans_quan"], *par["las_read_offset"]]
        world[b] = coor_img2pc.transform_coordinate_from_img_2_pc(par, seqs[b].copy(), [int(v) for v in lens[b]],
                                                                  Image.fromarray(images[b]))
    np.savez_compressed(os.path.join(HERE, "inverse_io3.npz"), images=images, seqs=seqs, lens=lens.astype(np.int32),
                        poses=poses, world=world)


def main():
    from PIL import Image
    import io_utils                                       # reference baseline/utils/io_utils.py
    import coor_img2pc                                    # reference baseline/utils/coor_img2pc.py
    from data import convert_data as ref_convert          # reference data/convert_data.py
    from lanemapping_b200.convert_data import _write_png
    from lanemapping_b200.sidecar import write_sidecar
    from oracle import bev_oracle as O

    spec, params, pts = golden_inputs()
    img = O.rasterize(pts, spec)["image"]
    png = os.path.join(HERE, "golden_crop.png")
    txt = os.path.join(HERE, "golden_crop.txt")
    _write_png(png, img)
    write_sidecar(txt, params)

    parsed = io_utils.load_pc_2_img_transform_paras(txt)
    with open(os.path.join(HERE, "sidecar_parsed.json"), "w") as f:
        json.dump(parsed, f, indent=1)

    seqs, lens = polylines()
    bev_img = Image.open(png)
    world = coor_img2pc.transform_coordinate_from_img_2_pc(parsed, seqs.copy(), lens, bev_img)
    with open(os.path.join(HERE, "inverse_io.json"), "w") as f:
        json.dump({"img_seqs": seqs.tolist(), "img_seq_lens": lens, "world": world.tolist()}, f, cls=ref_convert.NpEncoder)

    # loader op sequence, reference baseline/datasets/laserlane_proposals.py:87-94
    import torchvision
    t = torchvision.transforms.functional.to_tensor(np.array(Image.open(png), dtype=np.uint8)).float()
    assert t.shape[1] == t.shape[2]
    t = t[0:3]
    with open(os.path.join(HERE, "loader_contract.json"), "w") as f:
        json.dump({"shape": list(t.shape), "dtype": str(t.dtype), "min": float(t.min()), "max": float(t.max()),
                   "sum_per_channel": [float(t[c].double().sum()) for c in range(3)],
                   "pixel_64_64_times_255": [float(v) for v in (t[:, 64, 64] * 255).round()]}, f, indent=1)

    import inspect
    src = inspect.getsource(ref_convert.write_instance_orientation_seq)
    with open(os.path.join(HERE, "label_frame.json"), "w") as f:
        json.dump({"tile_px_hardcoded": 1152 if "(1152, 1152)" in src else None,
                   "pool_processes": 12 if "num_process = 12" in inspect.getsource(ref_convert.multiprocessing_seqs_files) else None},
                  f, indent=1)
    make_inverse2(io_utils, coor_img2pc, ref_convert)
    make_inverse3(coor_img2pc)
    make_labels(ref_convert)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
