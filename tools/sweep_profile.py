"""Dev tool (GPU box): time algo='auto' on config 2 and print the sweep's event counters (-DLM_SWEEP_DEBUG build)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lanemapping_b200.bev import BevRasterizer
from lanemapping_b200.synth import config_spec, make_cloud

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
spec, _ = config_spec(2)
if n < 100_000_000:
    spec = spec.window(0, max(1152, int(spec.height * n / 100_000_000) // 128 * 128))
pts = torch.from_numpy(make_cloud(n, spec, order="scan")).cuda()
r = BevRasterizer(spec, n, algo="sweep")
out = r.alloc_outputs()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r(pts, out=out); e1.record(); e1.synchronize()
    print(f"rep {rep}: {e0.elapsed_time(e1):.3f} ms  state {r.sweep_state()}")
d = r.sweep_debug()
if any(d):
    k = 3.0
    names = ["prod batches", "prod extra rounds", "prod markers", "prod wait ns (rounds)", "prod TMA wait ns", "prod total ns",
             "cons polls", "cons hits", "cons gated", "cons emit ns", "cons total ns"]
    for nm, v in zip(names, d):
        print(f"  {nm:24s} {v / k:14.0f} per call")
    print(f"  per producer CTA: total {d[5]/k/296/1e3:.1f} us, round-wait {d[3]/k/296/1e3:.1f} us, TMA wait {d[4]/k/296/1e3:.1f} us, batches {d[0]/k/296:.0f}")
    print(f"  per consumer CTA: total {d[10]/k/148/1e3:.1f} us, emit {d[9]/k/148/1e3:.1f} us, polls {d[6]/k/148:.0f}, hit rate {d[7]/max(d[6],1):.2f}")
