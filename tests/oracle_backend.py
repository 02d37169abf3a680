"""A numpy-oracle back end for lanemapping_b200.strips (tests only): lets the strip-sharding
host logic (partition, halo exchange, merge order, gather) run on CPU tensors over gloo."""
import numpy as np
import torch

from oracle import bev_oracle as O


class _OracleRaster:
    def __init__(self, spec, outputs):
        self.spec, self.outputs = spec, outputs
        self.max_points = 1 << 62

    def alloc_outputs(self):
        H, W, C = self.spec.height, self.spec.width, self.spec.n_channels
        out = {"image": torch.zeros((H, W, C), dtype=torch.uint8)}
        if "acc" in self.outputs:
            out["acc"] = torch.zeros((6, H, W), dtype=torch.int32)
        return out

    def __call__(self, points, out=None):
        out = out if out is not None else self.alloc_outputs()
        acc = O.accumulate(points.numpy(), self.spec)
        out["image"].copy_(torch.from_numpy(O.finalize(acc, self.spec)["image"]))
        if "acc" in out:
            out["acc"].copy_(torch.from_numpy(acc.view(np.int32)))
        return out


class OracleBackend:
    def make(self, spec, max_points, outputs, acc_band):
        return _OracleRaster(spec, outputs)

    def merge(self, dst, src):
        a = dst.numpy().view(np.uint32)
        merged = O.merge_acc(a, src.numpy().view(np.uint32))
        dst.copy_(torch.from_numpy(merged.view(np.int32)))

    def finalize(self, spec, acc, r0, r1, out):
        img = O.finalize(acc.numpy().view(np.uint32)[:, r0:r1], spec)["image"]
        out["image"][r0:r1] = torch.from_numpy(img)
