"""Smallest possible GPU check of bev.PipelinedRasterizer (two-stream overlap mode): parity with the C oracle."""
import os, sys, time
t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lanemapping_b200 import BevSpec
from lanemapping_b200.bev import PipelinedRasterizer
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import c_oracle as C
spec = BevSpec(1152, 1152, local_min_ele=default_min_ele(BevSpec(1152, 1152)))
clouds = [make_cloud(800_000 + 50_000 * i, spec, seed=70 + i, order="scan" if i % 2 else "shuffled") for i in range(4)]
want = [C.rasterize(c, spec)["image"] for c in clouds]
dev = [torch.from_numpy(c).cuda() for c in clouds]
pr = PipelinedRasterizer(spec, max(len(c) for c in clouds))
slots = [pr.submit(d) for d in dev]
ok = [np.array_equal(pr.result(slots[i])["image"].cpu().numpy(), want[i]) for i in (2, 3)]
for i in (0, 1):
    s = pr.submit(dev[i])
    ok.append(np.array_equal(pr.result(s)["image"].cpu().numpy(), want[i]))
pr.flush(); torch.cuda.synchronize(); pr.check_device_errors()
print("pipelined parity:", ok, "slots", slots, f"{time.time() - t0:.1f}s")
