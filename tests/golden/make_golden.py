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
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
