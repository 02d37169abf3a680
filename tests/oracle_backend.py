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

    def planes(self, spec):
        # the same derivation as lanemapping_b200.bev.needed_planes, restated on the oracle's own ids
        need = {O.CH_MAX_I: (O.ACC_MAX_I,), O.CH_MEAN_I: (O.ACC_COUNT, O.ACC_SUM_I), O.CH_MIN_Z: (O.ACC_MIN_Z,),
                O.CH_MAX_Z: (O.ACC_MAX_Z,), O.CH_MEAN_Z: (O.ACC_COUNT, O.ACC_SUM_Z), O.CH_DENSITY: (O.ACC_COUNT,)}
        s = set()
        for c in spec.channels:
            s.update(need[c])
        if spec.count16:
            s.add(O.ACC_COUNT)
        return sorted(s)

    def merge_finalize(self, spec, acc, r0, r1, recv, planes, out):
        a = acc.numpy().view(np.uint32)
        other = a[:, r0:r1].copy()
        other[O.ACC_COUNT] = other[O.ACC_SUM_I] = other[O.ACC_SUM_Z] = other[O.ACC_MAX_I] = other[O.ACC_MAX_Z] = 0
        other[O.ACC_MIN_Z] = O.MIN_Z_EMPTY                       # planes that did not travel merge as empty
        for k, pl in enumerate(planes):
            other[pl] = recv[k].numpy().view(np.uint32)
        merged = O.merge_acc(a[:, r0:r1], other)
        acc[:, r0:r1] = torch.from_numpy(merged.view(np.int32))
        out["image"][r0:r1] = torch.from_numpy(O.finalize(merged, spec)["image"])
