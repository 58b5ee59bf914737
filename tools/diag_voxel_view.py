#!/usr/bin/env python3
"""Where does the voxel view (phong.frag:347-404) in the tessellation warp differ from the oracle?  Prints the differing-pixel count,
its distribution over image rows / columns, and the first few pixels with both colours.  usage: diag_voxel_view.py [W H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.oracle_lib import Oracle
from tests.test_glsl_ref import pbr_room
from vct_b200 import params as P
from vct_b200 import scene as S
from vct_b200.pipeline import Pipeline

D, L, SS = 64, 5, 512
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 240)
sc = pbr_room()
g = Pipeline(sc, D, L, SS, W, H)
o = Oracle(sc, D, L, SS, W, H)
for lod in (0.0, 1.3):
    p = S.room_params(W, H); p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1; p.voxelize_tesselation_warp = 1
    p.debug_view = P.VIEW_VOXELS; p.miplevel = lod
    o.frame(p); g.frame(p)
    a = g.read_image().reshape(H, W); b = o.image.reshape(H, W)
    d = a != b
    print(f"lod {lod}: {d.sum()} of {W * H} pixels differ; volumes equal: {np.array_equal(g.read_volume(P.VOL_RADIANCE), o.radiance[0])}")
    rows = np.nonzero(d.sum(1))[0]; cols = np.nonzero(d.sum(0))[0]
    print("  rows with differences:", [(int(r), int(d[r].sum())) for r in rows[:40]])
    print("  columns with differences:", len(cols), "of", W)
    ys, xs = np.nonzero(d)
    for y, x in list(zip(ys, xs))[:8]:
        print(f"  ({x},{y}) gpu {a[y, x]:08x} oracle {b[y, x]:08x}")
g.close()
