"""The DATASETS / PCENCODER plug-ins against the reference's own registry machinery.

CPU part (here): the reference's ``baseline/utils/registry.py`` is loaded by path where /root/reference is
mounted (tests/ref_compat.py restates it for the GPU box and is held equal to it), registries are created the
way ``baseline/models/registry.py:5-36`` and ``baseline/datasets/registry.py:15-25`` do, the plug-ins are
registered and built with ``build_from_cfg``, and the on-the-fly dataset runs inside a forked
``DataLoader(num_workers=2, pin_memory=...)`` with ``collate_points`` exactly as
``baseline/datasets/registry.py:54-59`` builds its loader.
GPU part: the batch goes through ``Runner.to_cuda`` (a stand-in held equal to the reference's method on the same batches) and a
``Detector1stage``-shaped net under ``nn.DataParallel`` (reference baseline/engine/runner.py:103).
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
from torch.utils.data import Dataset

import ref_compat as RC
from lanemapping_b200 import BevSpec, sidecar
from lanemapping_b200 import datasets as lm_datasets
from lanemapping_b200 import pcencoder as lm_pcencoder
from lanemapping_b200.synth import default_min_ele, make_cloud

STEMS = ["000000_0001", "000000_0002", "000000_0003", "000000_0004"]


def make_root(root, n0=40_000):
    """A tiny dataset root in the layout of reference README.md:107-145 + crop_points/."""
    for d in ("crop_points", "cropped_tiff_param"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    specs, clouds = [], []
    for i, stem in enumerate(STEMS):
        spec = BevSpec(1152, 1152, bev_img_offset=(7.0 * i, -3.0), local_min_ele=default_min_ele(BevSpec(1152, 1152)))
        cloud = make_cloud(n0 + 1777 * i, spec, seed=10 + i, order="scan")           # ragged on purpose
        np.savez(os.path.join(root, "crop_points", stem + ".npz"), points=cloud,
                 geom=np.array([*spec.bev_img_offset, *spec.img_reso, spec.local_min_ele, spec.ele_reso, 0, 0]))
        sidecar.write_sidecar(os.path.join(root, "cropped_tiff_param", stem + ".txt"),
                              sidecar.PcImgParams("x.las", (0.0, 0.0, 0.0), (0, 0, 0, 1, 0, 0, 0), spec.bev_img_offset,
                                                  spec.img_reso, spec.local_min_ele, spec.ele_reso))
        specs.append(spec)
        clouds.append(cloud)
    with open(os.path.join(root, "split.json"), "w") as f:
        json.dump({"train": STEMS, "test": STEMS, "valid": STEMS, "single": STEMS[:1], "pretrain": STEMS}, f)
    return specs, clouds


class StubLaserLaneProposal(Dataset):
    """The three things ``make_onthefly_dataset`` uses of the reference's LaserLaneProposal (its real module
    needs skimage / laspy / mmdet3d, absent here): ``image_stem_list`` + ``mode`` set by ``__init__(data_root,
    data_split_file, mode, cfg=None)`` (reference laserlane_proposals.py:37-70,500-520) and
    ``format_gt_column_proposal(idx)`` returning the label dict (:102-252)."""

    def __init__(self, data_root, data_split_file, mode, cfg=None):
        assert mode in {"train", "valid", "test", "single", "all", "infer_only"}     # reference :39
        self.data_root, self.mode, self.cfg = data_root, mode, cfg
        self.image_stem_list = lm_datasets.read_split(data_root, data_split_file, mode)

    def __len__(self):
        return len(self.image_stem_list)

    def format_gt_column_proposal(self, idx):
        return {"label": torch.full((72, 144), idx, dtype=torch.int64), "endp_map": torch.zeros(1, 1152, 1152)}


class Cfg(dict):          # attribute access like the reference's Config (baseline/utils/config.py)
    __getattr__ = dict.__getitem__


def test_restated_registry_behaves_like_the_reference():
    ref = RC.load_reference_registry()
    if ref is None:
        pytest.skip("/root/reference is not mounted")
    for mod in (ref, RC):
        R = mod.Registry("pcencoder")

        @R.register_module
        class A:
            def __init__(self, x, cfg=None):
                self.x, self.cfg = x, cfg
        assert R.get("A") is A and R.get("B") is None and R.name == "pcencoder"
        a = mod.build_from_cfg(dict(type="A", x=3), R, default_args=dict(cfg="c"))
        assert (a.x, a.cfg) == (3, "c")
        assert mod.build_from_cfg(dict(type=A, x=1, cfg="own"), R, default_args=dict(cfg="c")).cfg == "own"
        with pytest.raises(KeyError):
            mod.build_from_cfg(dict(type="B"), R)
        with pytest.raises(KeyError):
            R.register_module(A)
        with pytest.raises(TypeError):
            R.register_module(lambda: 0)


class _Movable:
    """An object with .cuda(): what a list of non-tensors holds in the reference's batches (mmdet3d point containers)."""

    def __init__(self, v):
        self.v, self.moved = v, False

    def cuda(self, *a, **k):
        m = _Movable(self.v)
        m.moved = True
        return m


def test_restated_to_cuda_behaves_like_the_reference_method(monkeypatch):
    """``Runner.to_cuda`` cut out of the reference's runner.py and our stand-in, side by side on the same batches
    (``.cuda`` patched to the identity: there is no GPU here).  Both crash on a ragged list of clouds -- the reason
    ``collate_points`` returns a ``PointBatch`` -- and both move a ``PointBatch`` through its own ``.cuda``."""
    ref_to_cuda = RC.load_reference_to_cuda()
    if ref_to_cuda is None:
        pytest.skip("/root/reference is not mounted")
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(lm_datasets.PointBatch, "cuda", lambda self, *a, **k: ("moved", self))

    def batches():
        g = torch.Generator().manual_seed(0)
        return [
            {"proj": torch.rand(2, 3, 4, 4, generator=g), "image_name": ["a", "b"], "meta": {"k": 1}},
            {"label": [torch.ones(3, 2), torch.zeros(3, 2)], "arr": [np.ones((2, 2), np.float32), np.zeros((2, 2), np.float32)]},
            {"objs": [_Movable(1), _Movable(2)], "x": torch.arange(4)},
            {"points": lm_datasets.PointBatch.from_list([torch.rand(5, 4, generator=g), torch.rand(3, 4, generator=g)]),
             "bev_geom": torch.zeros(2, 8)},
        ]

    def norm(v):
        if isinstance(v, torch.Tensor):
            return ("tensor", tuple(v.shape), v.tolist())
        if isinstance(v, list):
            return [norm(x) for x in v]
        if isinstance(v, _Movable):
            return ("movable", v.v, v.moved)
        if isinstance(v, tuple) and v and v[0] == "moved":
            return ("pointbatch moved", len(v[1]))
        return v

    for b_ref, b_own in zip(batches(), batches()):
        got_ref, got_own = ref_to_cuda(None, b_ref), RC.runner_to_cuda(None, b_own)
        assert {k: norm(v) for k, v in got_ref.items()} == {k: norm(v) for k, v in got_own.items()}
    ragged = lambda: {"points": [torch.zeros(5, 4), torch.zeros(3, 4)]}
    for fn in (ref_to_cuda, RC.runner_to_cuda):
        with pytest.raises(RuntimeError):
            fn(None, ragged())


def build_plugins():
    """Registries as the reference creates them + our two registrations."""
    mod, is_ref = RC.registries()
    PCENCODER = mod.Registry("pcencoder")            # reference baseline/models/registry.py:5
    DATASETS = mod.Registry("datasets")              # reference baseline/datasets/registry.py:12

    @PCENCODER.register_module
    class PostProjector2(nn.Module):
        """Stand-in with the stock encoder's interface (reference postprojector.py:57-82): reads
        ``sample['proj']``, returns (fea, fea_up, bi_seg, endp) of the documented shapes."""

        def __init__(self, resnet="resnet34", cfg=None):
            super().__init__()
            self.cfg = cfg
            self.w = nn.Parameter(torch.ones(1))

        def forward(self, sample):
            proj = sample["proj"] * self.w
            B = proj.shape[0]
            fea = nn.functional.avg_pool2d(proj, 8).mean(1, keepdim=True).expand(B, 64, 144, 144)
            fea_up = nn.functional.avg_pool2d(proj, 4).mean(1, keepdim=True).expand(B, 8, 288, 288)
            return fea, fea_up, proj, proj[:, :1]

    lm_pcencoder.register(PCENCODER, mod.build_from_cfg)
    lm_datasets.register(DATASETS, base=StubLaserLaneProposal)
    return mod, is_ref, PCENCODER, DATASETS


def test_plugins_register_and_build_through_the_registry(tmp_path):
    mod, is_ref, PCENCODER, DATASETS = build_plugins()
    assert PCENCODER.get("OnTheFlyPostProjector") is not None and DATASETS.get("LaserLaneProposalOnTheFly") is not None
    cfg = Cfg(pcencoder=dict(type="OnTheFlyPostProjector", inner=dict(type="PostProjector2", resnet="resnet34")),
              seed=2021, batch_size=2, workers=2, distributed=False)
    # reference baseline/models/registry.py:20-24: build(cfg.pcencoder, PCENCODER, default_args=dict(cfg=cfg))
    enc = mod.build_from_cfg(cfg.pcencoder, PCENCODER, default_args=dict(cfg=cfg))
    assert isinstance(enc, lm_pcencoder.OnTheFlyProjector) and enc.cfg is cfg
    assert type(enc.inner).__name__ == "PostProjector2" and enc.inner.cfg is cfg      # inner built with cfg too
    # reference baseline/datasets/registry.py:24-25: kwargs = the split dict minus 'type' + cfg
    make_root(str(tmp_path))
    split_cfg = dict(type="LaserLaneProposalOnTheFly", data_root=str(tmp_path), data_split_file="split.json", mode="test")
    ds = mod.build_from_cfg(split_cfg, DATASETS, default_args=dict(cfg=cfg))
    assert isinstance(ds, StubLaserLaneProposal) and len(ds) == 4 and ds.cfg is cfg
    s = ds[2]
    assert s["image_name"] == STEMS[2] and s["points"].shape[1] == 4 and s["label"][0, 0] == 2
    # a second registration of the same name fails exactly like any reference module would
    with pytest.raises(KeyError):
        lm_datasets.register(DATASETS, base=StubLaserLaneProposal)


def loader_for(ds, cfg, pin):
    """reference baseline/datasets/registry.py:33-62 with the one documented change (collate_fn)."""
    from functools import partial

    def worker_init_fn(worker_id, seed):
        np.random.seed(worker_id + seed)
    return torch.utils.data.DataLoader(ds, batch_size=cfg.batch_size, sampler=torch.utils.data.SequentialSampler(ds),
                                       num_workers=cfg.workers, pin_memory=pin, drop_last=False,
                                       worker_init_fn=partial(worker_init_fn, seed=cfg.seed),
                                       collate_fn=lm_datasets.collate_points, multiprocessing_context="fork")


def test_onthefly_dataset_in_forked_dataloader_workers(tmp_path):
    mod, _, _, DATASETS = build_plugins()
    cfg = Cfg(seed=2021, batch_size=2, workers=2, distributed=False)
    _, clouds = make_root(str(tmp_path))
    ds = mod.build_from_cfg(dict(type="LaserLaneProposalOnTheFly", data_root=str(tmp_path),
                                 data_split_file="split.json", mode="test"), DATASETS, default_args=dict(cfg=cfg))
    batches = list(loader_for(ds, cfg, pin=False))
    assert len(batches) == 2
    for bi, b in enumerate(batches):
        pb = b["points"]
        assert isinstance(pb, lm_datasets.PointBatch) and len(pb) == 2
        assert b["image_name"] == STEMS[2 * bi:2 * bi + 2] and b["bev_geom"].shape == (2, 8)
        assert b["label"].shape == (2, 72, 144)
        for j in range(2):
            assert np.array_equal(pb[j].numpy(), clouds[2 * bi + j])                 # ragged clouds survive intact
    # the reference's to_cuda would have crashed on the old list form: cat of ragged [1, N_i, 4] tensors
    with pytest.raises(RuntimeError):
        torch.cat([c.unsqueeze(0) for c in batches[0]["points"].clouds()], dim=0)
    # the dense form for multi-device DataParallel: NaN padding + counts
    dense = lm_datasets.collate_points_padded([ds[0], ds[1]])
    assert dense["points"].shape == (2, len(clouds[1]), 4) and dense["points_count"].tolist() == [len(clouds[0]), len(clouds[1])]
    assert torch.isnan(dense["points"][0, len(clouds[0]):]).all()
    got = lm_pcencoder.OnTheFlyProjector.clouds_of(dense)
    assert np.array_equal(got[0].numpy(), clouds[0]) and np.array_equal(got[1].numpy(), clouds[1])


@pytest.mark.gpu
def test_batch_through_to_cuda_and_dataparallel_matches_oracle(native_lib, tmp_path):
    """DataLoader (forked workers, pinned) -> Runner.to_cuda -> nn.DataParallel(net) -> sample['proj']."""
    from oracle import bev_oracle as O
    mod, _, PCENCODER, DATASETS = build_plugins()
    cfg = Cfg(pcencoder=dict(type="OnTheFlyPostProjector", inner=dict(type="PostProjector2")),
              seed=2021, batch_size=4, workers=2, distributed=False, gpus=1)
    specs, clouds = make_root(str(tmp_path), n0=150_000)
    ds = mod.build_from_cfg(dict(type="LaserLaneProposalOnTheFly", data_root=str(tmp_path),
                                 data_split_file="split.json", mode="test"), DATASETS, default_args=dict(cfg=cfg))

    class Detector1stage(nn.Module):          # the consumer: reference baseline/models/net/detector1stage.py:11-28
        def __init__(self, cfg):
            super().__init__()
            self.cfg = cfg
            self.pcencoder = mod.build_from_cfg(cfg.pcencoder, PCENCODER, default_args=dict(cfg=cfg))

        def forward(self, batch):
            fea, fea_up, bi_seg, endp_est = self.pcencoder(batch)
            return {"fea": fea, "fea_up": fea_up, "semantic_seg": bi_seg, "endp_est": endp_est, "proj": batch["proj"]}

    net = torch.nn.parallel.DataParallel(Detector1stage(cfg), device_ids=range(cfg.gpus)).cuda()   # runner.py:103-104
    (batch,) = list(loader_for(ds, cfg, pin=True))
    assert batch["points"].points.is_pinned()
    batch = RC.runner_to_cuda(None, batch)                                           # runner.py:174
    assert batch["points"].points.is_cuda and batch["label"].is_cuda and batch["image_name"] == STEMS
    out = net(batch)
    torch.cuda.synchronize()
    assert out["fea"].shape == (4, 64, 144, 144) and out["fea_up"].shape == (4, 8, 288, 288)
    assert out["semantic_seg"].shape == (4, 3, 1152, 1152) and out["endp_est"].shape == (4, 1, 1152, 1152)
    for i in range(4):
        want = O.proj_from_image(O.rasterize(clouds[i], specs[i])["image"])
        assert np.array_equal(out["proj"][i].cpu().numpy(), want)
    # the dense form through the same net: the rasteriser drops the NaN padding
    dense = RC.runner_to_cuda(None, lm_datasets.collate_points_padded([ds[i] for i in range(4)]))
    out2 = net(dense)
    torch.cuda.synchronize()
    assert torch.equal(out2["proj"], out["proj"])
