#!/usr/bin/env python3
"""A/B the cone-trace tuning variants (VCT_TRACE_VARIANT, see cone_trace.cu) on the BASELINE configuration:
per-variant kernel time (CUDA events, median of N frames) and image PSNR against variant 1 (all-texture-unit path).
usage: trace_variants.py [frames] [variants...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from vct_b200.pipeline import Pipeline

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variants = [int(v) for v in sys.argv[2:]] or [1, 0, 3, 2, 5, 4]
sc, p, D, W, H, data = bench.build_workload()
ref = None
for v in variants:
    os.environ["VCT_TRACE_VARIANT"] = str(v)
    g = Pipeline(sc, D, bench.LEVELS, bench.SHADOW, W, H)
    g.frame(p); g.frame(p)
    g.set_profiling(2)
    ts, tf, tp = [], [], []
    for _ in range(frames):
        g.cone_trace(p); g.sync()
        ts.append(g.kernel_times()["k_cone_trace"][0] / 1e3)
    for _ in range(frames):                      # inside whole GI steps: the voxel passes evict the trace's inputs from L2
        g.gi_passes(p); g.sync()
        kt = g.kernel_times()
        tf.append(kt["k_cone_trace"][0] / 1e3); tp.append(kt.get("k_l2_prefetch", (0, 0))[0] / 1e3)
    img = g.read_image().view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    if ref is None:
        ref = img
    mse = ((img - ref) ** 2).mean()
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"variant {v}: k_cone_trace standalone median {statistics.median(ts):8.1f} us  min {min(ts):8.1f} us | in GI step {statistics.median(tf):8.1f} us (+ prefetch {statistics.median(tp):5.1f} us)   PSNR vs first {psnr:6.2f} dB  steps {g.cone_steps()}", flush=True)
    g.close()
