"""CPU model of bin_points' store pattern (no GPU): per 1024-point batch the distinct tiles, per warp-wide
store the 32-byte sectors it touches, and the share of lanes whose record is alone in its tile -- for the
config-2 cloud and for one rank's window of config 3.  Output quoted in DESIGN.md section 4."""
import numpy as np, sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from lanemapping_b200.strips import strip_bounds
from lanemapping_b200.synth import config_spec, make_cloud
from oracle import bev_oracle as O

def stats(name, spec, cloud, n_ctas, th_log2=6, sample_ctas=6):
    n=len(cloud); nb=(n+1023)//1024
    rng=np.random.default_rng(0)
    tiles_x=(spec.width+127)//128
    out=[]
    for b in rng.choice(n_ctas, sample_ctas, replace=False):
        b0=nb*b//n_ctas; b1=nb*(b+1)//n_ctas
        pts=cloud[b0*1024:b1*1024]
        row,col,iq,zq,valid=O.quantise_points(pts,spec)
        t=np.where(valid,(row>>th_log2)*tiles_x+(col>>7),-1)
        nbat=len(t)//1024
        t=t[:nbat*1024].reshape(nbat,1024)
        D=[len(np.unique(x[x>=0])) for x in t]
        # warp-store j of a batch = points j*256 + warp*32 .. +32 -> i.e. consecutive 32 points
        w=t.reshape(nbat*32,32)
        sect=[]; lone=0; tot=0
        for x in w[::7]:
            x=x[x>=0]
            u,c=np.unique(x,return_counts=True)
            sect.append(np.sum((c+7)//8))
            lone+=np.sum(c==1); tot+=len(x)
        touched=len(np.unique(t[t>=0]))
        out.append((np.mean(D),np.mean(sect),lone/tot,touched))
    o=np.array(out).mean(0)
    print(f"{name:12s} CTAs {n_ctas}: distinct tiles/batch {o[0]:.1f}, sectors per warp-store >= {o[1]:.1f}, lone-record lanes {o[2]*100:.1f}%, tiles touched per CTA {o[3]:.0f}")

spec,n=config_spec(2)
c=make_cloud(n,spec,order="scan")
stats("cfg2 scan",spec,c,740)
del c
s8,n8,_=workload(8,0,125_000_000)
r0,r1=strip_bounds(s8.height,8,32)[3]
win=s8.window(r0-64,r1+64)
c=make_cloud(n8,s8.window(r0,r1),seed=3)
stats("strip rank",win,c,444)
stats("strip rank",win,c,740)
