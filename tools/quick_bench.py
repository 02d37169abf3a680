"""Quick device-resident timing of lm_bev_rasterize (dev tool; bench.py is the contract)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lanemapping_b200.synth import make_cloud, config_spec
from lanemapping_b200.bev import BevRasterizer

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--orders", default="scan,shuffled")
ap.add_argument("--algos", default="binned,direct")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tune", default="", help="';'-separated tuning variants of the binned call, each 'field=value,field=value' (e.g. "
                "'bin_compact_table=1,tile_h_log2=5;bin_compact_table=-1'); timed after the default")
ap.add_argument("--las", action="store_true", help="time the LAS front end (decode, fused LAS->BEV) on the cfg's cloud")
a = ap.parse_args()
spec, n = config_spec(a.cfg)
n = a.n or n
if a.cfg == 5:
    # batch-8 on-the-fly proj (BASELINE configs[4]): 8 crops x 1e7 points -> f32 [8,3,1152,1152]
    from lanemapping_b200.pcencoder import BatchProjector
    clouds = [torch.from_numpy(make_cloud(n, spec, seed=100 + b, order="scan")).cuda() for b in range(8)]
    bp = BatchProjector()
    out = torch.empty((8, 3, 1152, 1152), dtype=torch.float32, device="cuda")
    for _ in range(3):
        bp(clouds, None, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bp(clouds, None, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    balg = 16 * 8 * n + out.numel() * 4
    print(f"cfg5 batch-8 proj: best {ms:.3f} ms  {8*n/ms/1e3:.1f} Mpts/s  {balg/ms/1e6:.1f} GB/s alg ({balg/ms/1e6/6538*100:.1f}% of 6538)")
    sys.exit(0)
def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

for order in a.orders.split(","):
    t0 = time.time()
    cloud = make_cloud(n, spec, order=order)
    pts = torch.from_numpy(cloud).cuda()
    print(f"cfg{a.cfg} {order}: generated {n} pts in {time.time()-t0:.1f}s", flush=True)
    if a.las:
        # the same cloud as LAS format-0 records (20 B: X, Y, Z int32 at 0.5 mm, intensity u16, 6 other bytes)
        from lanemapping_b200 import _cabi
        from lanemapping_b200.bev import decode_las
        rec = np.zeros((n, 20), dtype=np.uint8)
        rec[:, :12] = np.rint(cloud[:, :3].astype(np.float64) / 0.0005).astype("<i4").view(np.uint8).reshape(n, 12)
        rec[:, 12:14] = cloud[:, 3].astype("<u2").view(np.uint8).reshape(n, 2)
        recs = torch.from_numpy(rec.reshape(-1)).cuda()
        del rec
        x = _cabi.make_las_xform(20, (0.0005,) * 3, (0.0,) * 3)
        dec = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        ms = timed(lambda: decode_las(recs, n, x, out=dec), a.reps)
        print(f"  las_decode        best {ms:8.3f} ms  {n/ms/1e3:9.1f} Mpts/s  {36*n/ms/1e6:7.1f} GB/s (20 B in + 16 B out)", flush=True)
        r = BevRasterizer(spec, n, outputs=["image"] + (["count16"] if spec.count16 else []))
        out = r.alloc_outputs()
        ms = timed(lambda: r.rasterize_las(recs, n, x, out=out), a.reps)
        balg = 20 * n + spec.algorithmic_bytes(n) - 16 * n
        print(f"  fused LAS->BEV    best {ms:8.3f} ms  {n/ms/1e3:9.1f} Mpts/s  {balg/ms/1e6:7.1f} GB/s alg ({balg/ms/1e6/6538*100:5.1f}% of 6538)", flush=True)
        img_fused = out["image"].clone()
        ms2 = timed(lambda: r(decode_las(recs, n, x, out=dec), out=out), a.reps)
        print(f"  decode + raster   best {ms2:8.3f} ms  (fused is {ms2/ms:.2f}x faster; outputs equal: {bool(torch.equal(img_fused, out['image']))})", flush=True)
        del recs, dec, r, out
    for algo in a.algos.split(","):
        outs = ["image"] + (["count16"] if spec.count16 else [])
        r = BevRasterizer(spec, n, algo=algo, outputs=outs)
        out = r.alloc_outputs()
        for _ in range(2):
            r(pts, out=out)
        torch.cuda.synchronize()
        st = r.stats()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r(pts, out=out); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        if algo == "binned":
            from lanemapping_b200 import _cabi
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            best = [1e9] * 3
            for _ in range(a.reps):
                for k, stg in enumerate((_cabi.STAGE_BIN, _cabi.STAGE_INDEX, _cabi.STAGE_REDUCE)):
                    ev[k].record(); r(pts, out=out, stages=stg)
                ev[3].record(); ev[3].synchronize()
                best = [min(b, ev[k].elapsed_time(ev[k + 1])) for k, b in enumerate(best)]
            print(f"  stages: bin {best[0]:.3f}  scan+index {best[1]:.3f}  reduce {best[2]:.3f} ms", flush=True)
        balg = spec.algorithmic_bytes(n)
        print(f"  {algo:7s} best {ms:8.3f} ms  median {np.median(ts):8.3f}  {n/ms/1e3:9.1f} Mpts/s  "
              f"{balg/ms/1e6:7.1f} GB/s alg ({balg/ms/1e6/6538*100:5.1f}% of 6538)  ws={r.workspace.numel()/1e9:.2f} GB stats={st}", flush=True)
        del r, out
    for variant in [v for v in a.tune.split(";") if v]:
        tuning = {k: int(v) for k, v in (kv.split("=") for kv in variant.split(","))}
        outs = ["image"] + (["count16"] if spec.count16 else [])
        r = BevRasterizer(spec, n, outputs=outs, tuning=tuning)
        out = r.alloc_outputs()
        ms = timed(lambda: r(pts, out=out), a.reps)
        ref = BevRasterizer(spec, n, algo="direct", outputs=outs)(pts)
        same = all(bool(torch.equal(out[k], ref[k])) for k in outs)
        print(f"  tuning {variant}: best {ms:8.3f} ms  {spec.algorithmic_bytes(n)/ms/1e6/6538*100:5.1f}% of 6538  "
              f"ws={r.workspace.numel()/1e9:.2f} GB  == direct: {same}  stats={r.stats()}", flush=True)
        del r, out, ref
    del pts
