"""Round-2 groundwork (CPU only): how many 4x4 / 8x8 / 16x16 blocks of the Sponza shadow map can injectRadiance prove to miss the volume, or to
hit a single voxel, from the 8 corners of the block's (x, y, depth-range) box?  Checks every verdict against the per-texel result.  DESIGN.md section 9."""
import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from tests.oracle_lib import Oracle, lib, ptr
import bench
sc, p, D, W, H, data = bench.build_workload()
S = bench.SHADOW
o = Oracle(sc, D, bench.LEVELS, S, 64, 64)
t=time.time(); o.shadowmap(p); print("shadow map", time.time()-t, "s")
sh = o.shadow.reshape(S,S)
# filtered depth at texel corners: mean of texels (x-1..x, y-1..y), border 1  (fp32 like the kernel: ((t00*.5+t10*.5)*.5 + (t01*.5+t11*.5)*.5)
pad = np.ones((S+1,S+1),np.float32); pad[1:,1:] = sh
top = pad[:-1,:-1]*np.float32(0.5) + pad[:-1,1:]*np.float32(0.5)
bot = pad[1:,:-1]*np.float32(0.5) + pad[1:,1:]*np.float32(0.5)
d = top*np.float32(0.5) + bot*np.float32(0.5)          # d[y,x]
m = np.array(p.ls_inverse[:],np.float64).reshape(4,4).T   # column-major -> m[row,col]
vmin=np.array(p.voxel_min[:],np.float64); vmax=np.array(p.voxel_max[:],np.float64); vc=np.array(p.voxel_center[:],np.float64)
def vox(x,y,dd):   # float64 evaluation of D*voxelLinearPosition(lsInverse*ndc)
    nx = x/S*2-1; ny = y/S*2-1; nz = dd*2-1
    out=[]
    for i in range(3):
        w = m[i,0]*nx + m[i,1]*ny + m[i,2]*nz + m[i,3]
        out.append((w - vc[i] - vmin[i])/(vmax[i]-vmin[i])*D)
    return out
ys,xs = np.mgrid[0:S,0:S]
v = vox(xs.astype(np.float64), ys.astype(np.float64), d.astype(np.float64))
idx = [np.trunc(c).astype(np.int64) for c in v]
inb = np.ones((S,S),bool)
for c in v: inb &= (c > -1.0) & (c < D)
key = np.where(inb, (idx[2]*D+idx[1])*D+idx[0], -1)
print("texels in volume:", inb.mean())
for B in (4,8,16):
    nb = S//B
    dB = d.reshape(nb,B,nb,B)
    dmin = dB.min(axis=(1,3)).astype(np.float64); dmax = dB.max(axis=(1,3)).astype(np.float64)
    x0 = (np.arange(nb)*B)[None,:].astype(np.float64); y0 = (np.arange(nb)*B)[:,None].astype(np.float64)
    lo = [np.full((nb,nb), np.inf) for _ in range(3)]; hi = [np.full((nb,nb), -np.inf) for _ in range(3)]
    for dx in (0,B-1):
        for dy in (0,B-1):
            for dd in (dmin,dmax):
                c = vox(x0+dx, y0+dy, dd)
                for i in range(3): lo[i]=np.minimum(lo[i],c[i]); hi[i]=np.maximum(hi[i],c[i])
    eps = 1e-3
    same = np.ones((nb,nb),bool); outside = np.zeros((nb,nb),bool)
    for i in range(3):
        same &= (np.floor(lo[i]-eps) == np.floor(hi[i]+eps)) & (lo[i]-eps > 0) & (hi[i]+eps < D)
        outside |= (hi[i]+eps < -1.0) | (lo[i]-eps >= D)          # whole block misses the volume on one axis
    kB = key.reshape(nb,B,nb,B)
    uniform = (kB.min(axis=(1,3)) == kB.max(axis=(1,3)))
    # exactness: every block the test calls coherent must be uniform with an in-volume key; every 'outside' block must have no texel in the volume
    bad_same = (same & ~(uniform & (kB.max(axis=(1,3)) >= 0))).sum()
    bad_out = (outside & (kB.max(axis=(1,3)) >= 0)).sum()
    rest = ~(same|outside)
    print(f"B={B}: coherent blocks {same.mean():.3f}, whole-block misses {outside.mean():.3f}, per-texel fallback {rest.mean():.3f}; false coherent {bad_same}, false misses {bad_out}; work vs per-texel: {rest.mean() + (1-rest.mean())*8/(B*B):.3f}")
