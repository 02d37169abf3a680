"""GPU parity of the batched call (lm_bev_rasterize_batch, BASELINE.json configs[4]): B clouds ->
B equally-shaped rasters in one set of launches, bit-identical to the oracle and to B separate calls."""
from dataclasses import replace

import numpy as np
import pytest
import torch

from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import bev_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bev(native_lib):
    from lanemapping_b200 import bev as B
    assert torch.cuda.is_available()
    return B


def sample_specs(common, B, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for b in range(B):
        out.append(replace(common, bev_img_offset=(float(np.float32(rng.uniform(-50, 50))), float(np.float32(rng.uniform(-50, 50)))),
                           local_min_ele=float(np.float32(default_min_ele(common) + rng.uniform(-0.5, 0.5))),
                           row0=int(rng.integers(-40, 40)), col0=int(rng.integers(-40, 40))))
    return out


def shifted_cloud(n, spec, seed, order="scan"):
    """A cloud of the synthetic generator placed where ``spec``'s window looks (make_cloud honours
    bev_img_offset and row0/col0)."""
    return make_cloud(n, spec, seed=seed, order=order)


def check_batch(bev, common, counts, outputs, seed=0):
    B = len(counts)
    specs = sample_specs(common, B, seed)
    clouds = [shifted_cloud(n, sp, seed=10 * seed + b, order="scan" if b % 2 == 0 else "shuffled") if n else
              np.zeros((0, 4), np.float32) for b, (n, sp) in enumerate(zip(counts, specs))]
    r = bev.BatchRasterizer(common, B, max(1, sum(counts)), outputs=outputs)
    out = r([torch.from_numpy(c).cuda() for c in clouds], specs)
    torch.cuda.synchronize()
    st = r.stats()
    assert st["error"] == 0, st
    valid = 0
    for b in range(B):
        acc = O.accumulate(clouds[b], specs[b])
        want = O.finalize(acc, specs[b])
        valid += int(acc[O.ACC_COUNT].sum())
        if "image" in outputs:
            assert np.array_equal(out["image"][b].cpu().numpy(), want["image"]), f"sample {b}: u8 image differs"
        if "proj" in outputs:
            assert np.array_equal(out["proj"][b].cpu().numpy(), O.proj_from_image(want["image"])), f"sample {b}: proj differs"
        if "count16" in outputs:
            assert np.array_equal(out["count16"][b].cpu().numpy(), want["count16"]), f"sample {b}: count16 differs"
    assert st["n_valid"] == valid
    return out


def test_batch8_crops_proj_and_image(bev):
    common = BevSpec(1152, 1152, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    check_batch(bev, common, [300_000 + 1111 * b for b in range(8)], ("proj", "image"))


def test_batch_ragged_and_empty_samples(bev):
    # sample sizes around the 1024-point batch of bin_points, one empty sample, one single point
    common = BevSpec(256, 384, channels=(CH_MAX_I, CH_MEAN_I, CH_MIN_Z, CH_MAX_Z), count16=True)
    check_batch(bev, common, [1024, 0, 1, 1023, 1025, 50_000, 2048, 7], ("image", "count16", "proj"), seed=3)


def test_batch_larger_than_one_launch_set(bev):
    # 40 samples > 32 per launch set: two sets, outputs offset by the first set's samples
    common = BevSpec(128, 256, channels=(CH_MEAN_Z, CH_DENSITY))
    check_batch(bev, common, [3000 + 17 * b for b in range(40)], ("image", "proj"), seed=5)


def test_batch_height_not_a_tile_multiple_falls_back_to_single_calls(bev):
    common = BevSpec(100, 200, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    check_batch(bev, common, [5000, 6000, 0, 7000], ("image", "proj"), seed=7)


def test_batch_equals_separate_calls(bev):
    common = BevSpec(1152, 1152, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY))
    specs = sample_specs(common, 3, seed=11)
    clouds = [torch.from_numpy(shifted_cloud(200_000, sp, seed=b)).cuda() for b, sp in enumerate(specs)]
    got = bev.BatchRasterizer(common, 3, 600_000, outputs=("image",))(clouds, specs)["image"]
    for b in range(3):
        one = bev.rasterize(clouds[b], specs[b], outputs=("image",))["image"]
        assert torch.equal(got[b], one)


def test_batch_argument_errors(bev):
    common = BevSpec(1152, 1152)
    r = bev.BatchRasterizer(common, 2, 1000, outputs=("proj",))
    pts = torch.zeros((10, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        r([pts, pts, pts])                                   # more clouds than the batch
    with pytest.raises(ValueError):
        r([pts], [replace(common, img_reso=(0.02, 0.02))])   # samples must share the resolution
    with pytest.raises(ValueError):
        r([torch.zeros((2000, 4), dtype=torch.float32, device="cuda")])   # more points than sized for
