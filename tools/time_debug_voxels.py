import sys, time; sys.path.insert(0, '.')
import numpy as np
from vct_b200.workloads import Workload
from vct_b200.pipeline import Pipeline
w = Workload(3)
g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H)
g.frame(w.params); g.sync()
for lod in (0.0, 1.5):
    w.params.miplevel = lod
    g.debug_voxels(w.params); g.sync()
    t = time.time(); g.debug_voxels(w.params); g.sync(); dt = time.time() - t
    img = g.read_image()
    print(f"sponza 256^3 1080p debug voxels lod {lod}: {dt*1e3:.2f} ms, {np.unique(img).size} colours, background pixels {(img == img[-1]).mean():.2f}")
g.counters(); g.close()
