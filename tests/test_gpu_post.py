"""GPU parity of include/lm_post.h against fixtures produced by the REFERENCE's own code
(tests/golden/make_golden.py) and against the oracles on random inputs.  Pinned parity: bit-exact."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import inverse_oracle as INV
from oracle import label_oracle as LO

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def post(native_lib):
    from lanemapping_b200 import post as P
    assert torch.cuda.is_available()
    return P


def t64(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def t32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()


# ---- label rasters -------------------------------------------------------------------------
def test_label_rasters_equal_reference_pngs(post):
    d = json.load(open(os.path.join(G, "labels_in.json")))
    got = post.label_rasters(t64(d["seqs"]), t32(d["lens"]), t32(d["semantic"]), t32(d["instance"]), t32(d["orient"]))
    for k in ("semantic", "instance", "orient", "endp"):
        want = np.array(Image.open(os.path.join(G, f"labels_{k}.png")))
        assert np.array_equal(got[k].cpu().numpy(), want), f"{k} label raster differs from the reference's PNG"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_label_rasters_random_polylines_equal_oracle(post, seed):
    rng = np.random.default_rng(seed)
    L, V, H, W = 9, 40, 384, 512
    lens = rng.integers(0, V + 1, L)
    lens[0], lens[1] = V, 1
    seqs = np.zeros((L, V, 2))
    for i in range(L):
        seqs[i, :lens[i], 0] = rng.uniform(0, H - 1e-6, lens[i])
        seqs[i, :lens[i], 1] = rng.uniform(0, W - 1e-6, lens[i])
    semantic, instance = rng.integers(0, 4, L), rng.integers(0, 300, L)
    orient = rng.integers(0, 11, (L, V))
    got = post.label_rasters(t64(seqs), t32(lens), t32(semantic), t32(instance), t32(orient), H, W)
    sem, ins, ori = LO.polyline_labels(seqs, lens, semantic, instance, orient, H, W)
    assert np.array_equal(got["semantic"].cpu().numpy(), sem)
    assert np.array_equal(got["instance"].cpu().numpy(), ins)
    assert np.array_equal(got["orient"].cpu().numpy(), ori)
    starts = seqs[:, 0]
    ends = np.array([seqs[i, max(lens[i], 1) - 1] for i in range(L)])
    assert np.array_equal(got["endp"].cpu().numpy(), LO.endpoint_map(starts, ends, H, W))


def test_label_rasters_no_lanes(post):
    got = post.label_rasters(torch.zeros((0, 1, 2), dtype=torch.float64, device="cuda"), t32([]), t32([]), t32([]),
                             torch.zeros((0, 1), dtype=torch.int32, device="cuda"), 64, 64)
    assert all(int(v.sum()) == 0 for v in got.values())


# ---- pixel -> world ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["inverse_io.json", "inverse_io2.json"])
def test_img2pc_equals_reference_output(post, name):
    d = json.load(open(os.path.join(G, name)))
    params = d.get("params") or json.load(open(os.path.join(G, "sidecar_parsed.json")))
    img = np.array(Image.open(os.path.join(G, "golden_crop.png")))
    images = torch.from_numpy(img[None]).cuda()
    seqs, lens = np.array(d["img_seqs"]), d["img_seq_lens"]
    work = images.clone()
    world = post.img2pc(work, t64(seqs[None]), t32([lens]), [params], fill_in_place=True)
    assert np.array_equal(world[0].cpu().numpy(), np.array(d["world"])), "world coordinates differ from the reference's"
    if "filled_px" in d:
        diff = np.argwhere(work[0, :, :, 1].cpu().numpy() != img[:, :, 1])
        assert [[int(r), int(c), int(work[0, r, c, 1])] for r, c in diff] == d["filled_px"]
    assert torch.equal(post.img2pc(images, t64(seqs[None]), t32([lens]), [params]), world)     # copy mode, same result
    assert np.array_equal(images[0].cpu().numpy(), img)                                         # and the input is untouched


def test_img2pc_batch_equals_reference_output_on_random_crops(post):
    """tests/golden/inverse_io3.npz: random crops / polylines / poses run through the reference itself;
    the three crops go through ONE batched call."""
    from test_label_oracle import io3_cases
    cases = list(io3_cases())
    images = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    seqs = t64(np.stack([c[2] for c in cases]))
    lens = t32(np.stack([c[3] for c in cases]))
    world = post.img2pc(images, seqs, lens, [c[0] for c in cases]).cpu().numpy()
    for b, c in enumerate(cases):
        assert np.array_equal(world[b], c[4]), f"crop {b}: world coordinates differ from the reference's"


def test_img2pc_batch_of_random_crops_equals_oracle(post):
    rng = np.random.default_rng(3)
    B, H, W, L, V = 5, 96, 80, 6, 17
    images = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    images[rng.random((B, H, W)) < 0.6] = 0                      # many empty pixels, holes of every size
    images[:, 10:40, 20:50] = 0
    seqs = np.zeros((B, L, V, 2))
    lens = rng.integers(0, V + 1, (B, L))
    lens[:, 0] = V
    for b in range(B):
        for l in range(L):
            seqs[b, l, :lens[b, l], 0] = rng.uniform(0, H - 1e-6, lens[b, l])
            seqs[b, l, :lens[b, l], 1] = rng.uniform(0, W - 1e-6, lens[b, l])
    params = []
    for b in range(B):
        q = rng.normal(size=4)
        params.append({"img_reso": [0.05, 0.04], "bev_img_offset": [float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))],
                       "ele_reso": 0.05, "local_min_ele": float(rng.uniform(-3, 3)),
                       "las_rotation_trans_quan": [*rng.uniform(-10, 10, 3).tolist(), *(q / np.linalg.norm(q)).tolist()],
                       "las_read_offset": [533000.0, 3380000.0, 20.0]})
    got = post.img2pc(torch.from_numpy(images).cuda(), t64(seqs), t32(lens), params).cpu().numpy()
    for b in range(B):
        keep = [l for l in range(L) if lens[b, l] > 0]          # the reference divides by zero on an empty line
        want = INV.img2pc(params[b], seqs[b, keep], [int(lens[b, l]) for l in keep], images[b])
        assert np.array_equal(got[b, keep], want), f"crop {b}"


# ---- loader colour augmentation ------------------------------------------------------------
def test_color_jitter_matches_torchvision(post):
    import torchvision
    import torchvision.transforms.functional as TF
    torch.manual_seed(2021)
    B, H, W = 7, 96, 160
    proj = (torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8).float() / 255.0)
    proj[:, :, :20] = 0.0                                                # empty cells
    aug = post.GpuColorJitter()
    params = aug.draw(B)
    params[1] = ([3, 2, 1, 0], None, 1.3, 0.6)                           # brightness skipped
    params[2] = ([1, 0, 3, 2], 1.5, None, None)                          # brightness only
    want = []
    for b in range(B):
        fn_idx, bf, cf, sf = params[b]
        img = proj[b].clone()
        for fn in fn_idx:                                                # torchvision ColorJitter.forward
            if fn == 0 and bf is not None:
                img = TF.adjust_brightness(img, bf)
            elif fn == 1 and cf is not None:
                img = TF.adjust_contrast(img, cf)
            elif fn == 2 and sf is not None:
                img = TF.adjust_saturation(img, sf)
        want.append(torchvision.transforms.Normalize(mean=[0.5], std=[0.5])(img))
    want = torch.stack(want)
    got = aug(proj.cuda().contiguous(), params).cpu()
    assert float((got - want).abs().max()) <= 1e-6               # tolerance: float32, summation order of mean(gray)
    # same RNG consumption as the reference's transform: a seeded draw equals torchvision's own
    torch.manual_seed(7)
    mine = post.GpuColorJitter().draw(2)
    torch.manual_seed(7)
    cj = torchvision.transforms.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5)
    for k in range(2):
        fn_idx, b, c, s, _ = cj.get_params(cj.brightness, cj.contrast, cj.saturation, cj.hue)
        assert mine[k] == (fn_idx.tolist(), b, c, s)


# ---- the reference-signature entry points (lanemapping_b200/ref_api.py) ----------------------
def test_ref_api_write_instance_orientation_seq_writes_the_reference_files(post, tmp_path):
    import cv2
    from lanemapping_b200 import ref_api
    d = json.load(open(os.path.join(G, "labels_in.json")))
    names = [str(tmp_path / k) for k in ("seq.json", "sem.png", "ins.png", "ori.png", "endp.png")]
    ref_api.write_instance_orientation_seq(np.array(d["seqs"]), d["lens"], d["semantic"], d["instance"],
                                           np.array(d["orient"]), *names)
    for k, path in zip(("semantic", "instance", "orient", "endp"), names[1:]):
        assert np.array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED), np.array(Image.open(os.path.join(G, f"labels_{k}.png"))))
    assert json.load(open(names[0])) == json.load(open(os.path.join(G, "labels_seq.json")))


def test_ref_api_transform_coordinate_from_img_2_pc(post):
    from lanemapping_b200 import ref_api
    d = json.load(open(os.path.join(G, "inverse_io2.json")))
    world = ref_api.transform_coordinate_from_img_2_pc(d["params"], np.array(d["img_seqs"]), d["img_seq_lens"],
                                                       Image.open(os.path.join(G, "golden_crop.png")))
    assert np.array_equal(world, np.array(d["world"]))
