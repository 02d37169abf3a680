"""CPU: oracle/label_oracle.py and oracle/inverse_oracle.py against the fixtures the REFERENCE's own
code produced (tests/golden/make_golden.py: write_instance_orientation_seq, transform_coordinate_from_img_2_pc,
modify_empty_pixel_elevation).  These two stages have in-tree reference code, so their parity is pinned."""
import json
import os

import numpy as np
from PIL import Image

from oracle import inverse_oracle as INV
from oracle import label_oracle as LO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_labels():
    d = json.load(open(os.path.join(G, "labels_in.json")))
    want = {k: np.array(Image.open(os.path.join(G, f"labels_{k}.png"))) for k in ("semantic", "instance", "orient", "endp")}
    return np.array(d["seqs"]), d["lens"], d["semantic"], d["instance"], np.array(d["orient"]), want


def test_polyline_labels_match_reference_output():
    seqs, lens, sem, ins, ori, want = load_labels()
    a, b, c = LO.polyline_labels(seqs, lens, sem, ins, ori)
    assert np.array_equal(a, want["semantic"]) and np.array_equal(b, want["instance"]) and np.array_equal(c, want["orient"])
    assert set(np.unique(want["semantic"])) == {0, 128, 255} and want["instance"].max() == 6


def test_endpoint_map_matches_reference_output():
    seqs, lens, *_, want = load_labels()
    starts = np.array([seqs[i, 0] for i in range(len(lens))])
    ends = np.array([seqs[i, lens[i] - 1] for i in range(len(lens))])
    got = LO.endpoint_map(starts, ends)
    assert np.array_equal(got, want["endp"])
    # lanes 0-2 and 5 contribute both end points (lane 3 lies in the clip border, lane 4 is degenerate)
    assert int((want["endp"] == 255).sum()) == 8


def test_line_model_is_cv2_line():
    import cv2
    rng = np.random.default_rng(7)
    for _ in range(2000):
        x1, y1, x2, y2 = (int(v) for v in rng.integers(0, 96, 4))
        img = np.zeros((96, 96), np.uint8)
        cv2.line(img, (x1, y1), (x2, y2), 1)
        mine = np.zeros((96, 96), np.uint8)
        for x, y in LO.line_pixels(x1, y1, x2, y2):
            mine[y, x] = 1
        assert np.array_equal(img, mine), (x1, y1, x2, y2)


def test_inverse_oracle_matches_reference_on_hole_filling_case():
    d = json.load(open(os.path.join(G, "inverse_io2.json")))
    img = np.array(Image.open(os.path.join(G, "golden_crop.png")))
    seqs, lens = np.array(d["img_seqs"]), d["img_seq_lens"]
    world = INV.img2pc(d["params"], seqs, lens, img)
    assert np.array_equal(world, np.array(d["world"]))                    # bit for bit
    filled = INV.fill_empty_elevation(img, seqs, lens)
    changed = np.argwhere(filled[:, :, 1] != img[:, :, 1])
    assert [[int(r), int(c), int(filled[r, c, 1])] for r, c in changed] == d["filled_px"] and len(changed) > 10


def io3_cases():
    z = np.load(os.path.join(G, "inverse_io3.npz"))
    for b in range(len(z["images"])):
        po = z["poses"][b]
        par = {"img_reso": [0.05, 0.04], "bev_img_offset": po[0:2].tolist(), "ele_reso": 0.05, "local_min_ele": float(po[2]),
               "las_rotation_trans_quan": po[3:10].tolist(), "las_read_offset": po[10:13].tolist()}
        yield par, z["images"][b], z["seqs"][b], [int(v) for v in z["lens"][b]], z["world"][b]


def test_inverse_oracle_matches_reference_on_random_crops():
    """Random data: pins LeastSuqare's left-to-right summation (np.sum would differ in the last bits)."""
    for par, img, seqs, lens, world in io3_cases():
        assert np.array_equal(INV.img2pc(par, seqs, lens, img), world)
