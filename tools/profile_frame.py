#!/usr/bin/env python3
"""Run a few whole frames of the BASELINE configuration (for ncu captures and per-kernel event timing).
usage: profile_frame.py [frames] [--kernels]   (--kernels: print per-kernel CUDA-event times of the last frame)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vct_b200.pipeline import Pipeline

frames = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 3
sc, p, D, W, H, data = bench.build_workload()
g = Pipeline(sc, D, bench.LEVELS, bench.SHADOW, W, H)
if "--kernels" in sys.argv:
    g.set_profiling(2)
for _ in range(frames):
    g.frame(p)
g.sync()
if "--kernels" in sys.argv:
    kt = g.kernel_times()
    for k, (ns, n) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:28s} {ns/1e3:10.1f} us  x{n}")
    print("total", sum(v[0] for v in kt.values()) / 1e3, "us")
info = g.counters()
print("fragments", info.total_fragments, "unique", info.unique_voxels, "max", info.max_fragments_per_voxel, "cone steps", g.cone_steps())
g.close()
