"""Single-GPU replay of one rank's strip step (window + banded accumulators + finalize of bands),
for kernel-level timing under ncu (a multi-rank command must never be wrapped in ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload
from lanemapping_b200.bev import BevRasterizer, acc_merge_, finalize_rows
from lanemapping_b200.strips import strip_bounds
from lanemapping_b200.synth import make_cloud
spec, n, _ = workload(8, 0, int(os.environ.get("N", "125000000")))
r0, r1 = strip_bounds(spec.height, 8, 128)[3]
local = spec.window(r0 - 64, r1 + 64)
cloud = torch.from_numpy(make_cloud(n, spec.window(r0, r1), seed=3)).cuda()
r = BevRasterizer(local, n, outputs=("image", "acc"), acc_band=128)
out = r.alloc_outputs()
for _ in range(3):
    r(cloud, out=out)
    acc_merge_(out["acc"][:, 64:128], out["acc"][:, 64:128].clone())
    finalize_rows(local, out["acc"], 64, 128, {"image": out["image"]})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r(cloud, out=out); e1.record(); e1.synchronize()
print("strip raster step ms:", e0.elapsed_time(e1), r.stats())
from lanemapping_b200 import _cabi
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
best = [1e9] * 3
for _ in range(5):
    for k, stg in enumerate((_cabi.STAGE_BIN, _cabi.STAGE_INDEX, _cabi.STAGE_REDUCE)):
        ev[k].record(); r(cloud, out=out, stages=stg)
    ev[3].record(); ev[3].synchronize()
    best = [min(b, ev[k].elapsed_time(ev[k + 1])) for k, b in enumerate(best)]
print(f"strip stages: bin {best[0]:.3f}  scan+index {best[1]:.3f}  reduce {best[2]:.3f} ms")
# the same points without halo bands / raw accumulators (what an N = 1 run of this geometry would do)
plain = spec.window(r0, r1)
r2 = BevRasterizer(plain, n, outputs=("image",))
o2 = r2.alloc_outputs()
for _ in range(2):
    r2(cloud, out=o2)
best = [1e9] * 3
for _ in range(5):
    for k, stg in enumerate((_cabi.STAGE_BIN, _cabi.STAGE_INDEX, _cabi.STAGE_REDUCE)):
        ev[k].record(); r2(cloud, out=o2, stages=stg)
    ev[3].record(); ev[3].synchronize()
    best = [min(b, ev[k].elapsed_time(ev[k + 1])) for k, b in enumerate(best)]
print(f"no-halo stages: bin {best[0]:.3f}  scan+index {best[1]:.3f}  reduce {best[2]:.3f} ms  {r2.stats()}")
