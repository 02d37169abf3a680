"""Seeded synthetic MLS road clouds.  numpy only.

No dataset ships with the reference and there is no network, so every test and
bench number uses clouds from this generator.  Distributions:

* along-track (row axis, the driving direction: reference data/convert_data.py:151-156)
  uniform over the scene length;
* cross-track 0.8 * N(centre, sigma = 6 m) + 0.2 * uniform  -> hot cells under the
  trajectory, sparse verges with empty cells;
* z = 2 % crossfall plane + N(0, 0.02 m), 2 % outliers U(0.2, 5 m) above it;
* intensity (LAS u16 value carried in a float): asphalt log-normal ~3000-8000, paint
  stripes 0.15 m wide every 3.75 m at 20000-33000, a few raw values outside the
  [800, 33000] clip range (reference baseline/datasets/laserlane_proposals.py:626-628)
  so that the clip is exercised;
* ~0.5 % of the points fall outside the grid and must be dropped, never clamped.

Two orderings: ``scan`` (acquisition order: road by road, each along-track monotone with
+-1 m jitter -- what MLS trajectories produce, spatially coherent) and ``shuffled`` (worst case).
"""
from __future__ import annotations

import numpy as np

from .spec import BevSpec

SEED = 2021  # reference configs/*:8


def make_cloud(n_points: int, spec: BevSpec, seed: int = SEED, order: str = "scan",
               chunk: int = 1 << 24, out: np.ndarray | None = None, roads: int | None = None) -> np.ndarray:
    """Return float32 [n_points, 4] (x, y, z, intensity) in the raster's local frame.

    Generated chunk by chunk (deterministic for a given (n_points, seed, order, chunk)) so
    that 1e8-point clouds need no float64 temporaries of full length.
    """
    if order not in ("scan", "shuffled"):
        raise ValueError("order must be 'scan' or 'shuffled'")
    n_points = int(n_points)
    if out is None:
        out = np.empty((n_points, 4), dtype=np.float32)
    assert out.shape == (n_points, 4) and out.dtype == np.float32
    length = spec.height * spec.img_reso[0]      # along-track extent  (rows)
    width = spec.width * spec.img_reso[1]        # cross-track extent  (cols)
    x_lo = spec.bev_img_offset[0] + spec.row0 * spec.img_reso[0]
    y_lo = spec.bev_img_offset[1] + spec.col0 * spec.img_reso[1]
    # one road per 57.6 m (one 1152-px crop at 0.05 m) of cross-track extent unless told otherwise
    if roads is None:
        roads = max(1, int(round(width / 57.6)))
    pitch = width / roads
    ss = np.random.SeedSequence([seed, n_points, 0 if order == "scan" else 1])
    n_chunks = max(1, -(-n_points // chunk))
    for ci, child in enumerate(ss.spawn(n_chunks)):
        lo, hi = ci * chunk, min(n_points, (ci + 1) * chunk)
        m = hi - lo
        if m <= 0:
            break
        rng = np.random.default_rng(child)
        # along-track; which of the `roads` parallel roads a point belongs to
        if order == "scan":
            # acquisition order: the vehicle drives one road after the other, so the cloud is the
            # concatenation of the roads' scans, each along-track monotone with +-1 m jitter
            idx = np.arange(lo, hi, dtype=np.float64)
            road = np.minimum(np.floor(idx * roads / n_points), roads - 1)
            per_road = n_points / roads
            x = (idx - road * per_road + rng.random(m)) * (length / per_road)
            x += rng.uniform(-1.0, 1.0, m)
        else:
            x = rng.random(m) * length
            road = rng.integers(0, roads, m).astype(np.float64)
        x += x_lo
        centre = y_lo + (road + 0.5) * pitch
        y = np.where(rng.random(m) < 0.8, rng.normal(centre, 6.0, m), y_lo + rng.random(m) * width)
        # ~0.5 % thrown well outside (either axis)
        outside = rng.random(m) < 0.005
        y = np.where(outside, y + np.where(rng.random(m) < 0.5, -1.0, 1.0) * (width + rng.random(m) * 10.0), y)
        # elevation
        z = 0.02 * (y - centre) + rng.normal(0.0, 0.02, m)
        z = np.where(rng.random(m) < 0.02, z + rng.uniform(0.2, 5.0, m), z)
        # intensity: asphalt + paint stripes (lane markings run along-track)
        inten = rng.lognormal(np.log(4500.0), 0.35, m)
        stripe = np.mod(y - centre + 0.075, 3.75) < 0.15
        inten = np.where(stripe, rng.uniform(20000.0, 33000.0, m), inten)
        raw = rng.random(m)
        inten = np.where(raw < 0.002, rng.uniform(0.0, 800.0, m), inten)         # below the clip
        inten = np.where(raw > 0.998, rng.uniform(33000.0, 65535.0, m), inten)   # above the clip
        inten = np.floor(np.clip(inten, 0.0, 65535.0))                           # u16 value
        o = out[lo:hi]
        o[:, 0] = x
        o[:, 1] = y
        o[:, 2] = z
        o[:, 3] = inten
    return out


def default_min_ele(spec: BevSpec) -> float:
    """A local_min_ele that keeps the synthetic crossfall plane inside the u8 height range."""
    return -0.02 * 0.5 * spec.width * spec.img_reso[1] - 0.5


# The five BASELINE.json configs
def config_spec(cfg: int) -> tuple[BevSpec, int]:
    from . import spec as S
    if cfg == 1:
        sp = BevSpec(1152, 1152, channels=(S.CH_MAX_I,))
        n = 10_000_000
    elif cfg == 2:
        sp = BevSpec(11520, 1152, channels=(S.CH_MAX_I, S.CH_MEAN_Z, S.CH_DENSITY))
        n = 100_000_000
    elif cfg == 3:
        sp = BevSpec(11520, 11520, channels=(S.CH_MAX_I, S.CH_MEAN_Z, S.CH_DENSITY))
        n = 1_000_000_000
    elif cfg == 4:
        sp = BevSpec(28800, 3456, img_reso=(0.02, 0.02), ele_reso=0.02,
                     channels=(S.CH_MAX_I, S.CH_MIN_Z, S.CH_MEAN_I, S.CH_MAX_Z), count16=True)  # index 1 = elevation
        n = 100_000_000
    elif cfg == 5:
        sp = BevSpec(1152, 1152, channels=(S.CH_MAX_I, S.CH_MEAN_Z, S.CH_DENSITY))
        n = 10_000_000  # per sample, batch 8
    else:
        raise ValueError("cfg must be 1..5")
    from dataclasses import replace
    return replace(sp, local_min_ele=default_min_ele(sp)), n
