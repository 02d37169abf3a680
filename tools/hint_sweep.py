"""A/B of bin_points' L2 policy hints (LM_BEV_STREAM_HINT) and CTA count on three geometries:
config 2 scan, one strip rank of config 3 (10 roads, 1440+128 rows x 11520), config 2 shuffled.
Each variant is also compared bit for bit with the first one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload
from lanemapping_b200 import _cabi
from lanemapping_b200.bev import BevRasterizer
from lanemapping_b200.strips import strip_bounds
from lanemapping_b200.synth import config_spec, make_cloud

def stages(r, pts, out, reps=7):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    best = [1e9] * 3
    for _ in range(2):
        r(pts, out=out)
    for _ in range(reps):
        for k, stg in enumerate((_cabi.STAGE_BIN, _cabi.STAGE_INDEX, _cabi.STAGE_REDUCE)):
            ev[k].record(); r(pts, out=out, stages=stg)
        ev[3].record(); ev[3].synchronize()
        best = [min(b, ev[k].elapsed_time(ev[k + 1])) for k, b in enumerate(best)]
    return best

def sweep(name, spec, cloud, outputs, band=0):
    pts = torch.from_numpy(cloud).cuda()
    ref = None
    for env in ({}, {"LM_BEV_STREAM_HINT": "1"}, {"LM_BEV_STREAM_HINT": "2"}, {"LM_BEV_STREAM_HINT": "3"},
                {"LM_BEV_BIN_CTAS_PER_SM": "2"}, {"LM_BEV_BIN_CTAS_PER_SM": "3"},
                {"LM_BEV_STREAM_HINT": "1", "LM_BEV_BIN_CTAS_PER_SM": "3"}):
        for k in ("LM_BEV_STREAM_HINT", "LM_BEV_BIN_CTAS_PER_SM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        r = BevRasterizer(spec, len(cloud), outputs=outputs, acc_band=band)
        out = r.alloc_outputs()
        b = stages(r, pts, out)
        r(pts, out=out); torch.cuda.synchronize()
        img = out["image"].clone()
        same = True if ref is None else bool(torch.equal(img, ref))
        ref = img if ref is None else ref
        print(f"{name:14s} {str(env):70s} bin {b[0]:.3f} index {b[1]:.3f} reduce {b[2]:.3f}  sum {sum(b):.3f} ms  same={same} err={r.stats()['error']}", flush=True)
        del r, out
    del pts

spec, n = config_spec(2)
sweep("cfg2 scan", spec, make_cloud(n, spec, order="scan"), ("image",))
s8, n8, _ = workload(8, 0, 125_000_000)
r0, r1 = strip_bounds(s8.height, 8, 128)[3]
sweep("strip rank", s8.window(r0 - 64, r1 + 64), make_cloud(n8, s8.window(r0, r1), seed=3), ("image", "acc"), band=128)
sweep("cfg2 shuffled", spec, make_cloud(n, spec, order="shuffled"), ("image",))
