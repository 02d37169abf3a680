"""Dev tool: where does algo='auto' differ from the oracle?  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lanemapping_b200 import BevSpec, CH_DENSITY, CH_MAX_I, CH_MEAN_Z
from lanemapping_b200.bev import BevRasterizer
from lanemapping_b200.synth import default_min_ele, make_cloud
from oracle import c_oracle as CO

h, w, n = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2304, 1152, 6_000_000)))
spec = BevSpec(h, w, channels=(CH_MAX_I, CH_MEAN_Z, CH_DENSITY), local_min_ele=default_min_ele(BevSpec(1152, 1152)))
cloud = make_cloud(n, spec, seed=5, order="scan")
want = CO.rasterize(cloud, spec)["image"]
r = BevRasterizer(spec, n, algo="sweep")
pts = torch.from_numpy(cloud).cuda()
for rep in range(2):
    got = r(pts)["image"].cpu().numpy()
    torch.cuda.synchronize()
    print("rep", rep, "state", r.sweep_state(), "stats", r.stats())
    bad = np.argwhere((got != want).any(axis=2))
    print(" mismatching cells:", len(bad), "of", h * w)
    if len(bad):
        rows, cols = bad[:, 0], bad[:, 1]
        print(" rows", rows.min(), rows.max(), "cols", cols.min(), cols.max())
        print(" by channel:", [(got[..., c] != want[..., c]).sum() for c in range(3)])
        print(" got>want / got<want density:", (got[..., 2] > want[..., 2]).sum(), (got[..., 2] < want[..., 2]).sum())
        print(" rows hist (per 128):", np.bincount(rows // 128))
        print(" col%4 hist:", np.bincount(cols % 4), " (col//4)%148 top:", np.bincount((cols // 4) % 148).argsort()[-5:])
        for k in range(min(8, len(bad))):
            i, j = bad[k]
            print("  cell", i, j, "got", got[i, j], "want", want[i, j])
        tot_g, tot_w = int(got[..., 2].astype(np.int64).sum()), int(want[..., 2].astype(np.int64).sum())
        print(" density sums got/want:", tot_g, tot_w)
