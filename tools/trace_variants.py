#!/usr/bin/env python3
"""A/B the cone-trace tuning variants (VCT_TRACE_VARIANT, see cone_trace.cu) on the BASELINE configuration:
per-variant kernel time (CUDA events, median of N frames), whole-step time (host clock around a batch of GI steps, profiling
off — the only figure that shows the side-stream overlap of bit 7) and image PSNR against the first variant listed.
Variant bits: 0 shared-memory last level, 3 L2 prefetch, 4-5 CTA size, 6 split set-up / march kernels, 7 set-up on a side stream
under the voxel passes, 8-9 march kernel at 4 / 5 / 6 CTAs per SM.  Round-2 A/B of the split: trace_variants.py 20 0 64 192 320 448 576 704
usage: trace_variants.py [frames] [variants...]"""
import os, sys, statistics, time
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
        kt = g.kernel_times()
        ts.append(sum(kt.get(k, (0, 0))[0] for k in ("k_cone_trace", "k_trace_setup", "k_trace_march")) / 1e3)
    for _ in range(frames):                      # inside whole GI steps: the voxel passes evict the trace's inputs from L2
        g.gi_passes(p); g.sync()
        kt = g.kernel_times()
        tf.append(sum(kt.get(k, (0, 0))[0] for k in ("k_cone_trace", "k_trace_setup", "k_trace_march")) / 1e3); tp.append(kt.get("k_l2_prefetch", (0, 0))[0] / 1e3)
    g.set_profiling(0)
    for _ in range(5):
        g.gi_passes(p)
    g.sync(); t0 = time.perf_counter()
    for _ in range(50):
        g.gi_passes(p)
    g.sync(); step_us = (time.perf_counter() - t0) / 50 * 1e6
    img = g.read_image().view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    if ref is None:
        ref = img
    mse = ((img - ref) ** 2).mean()
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"variant {v}: k_cone_trace standalone median {statistics.median(ts):8.1f} us  min {min(ts):8.1f} us | in GI step {statistics.median(tf):8.1f} us (+ prefetch {statistics.median(tp):5.1f} us) | whole GI step {step_us:7.1f} us   PSNR vs first {psnr:6.2f} dB  steps {g.cone_steps()}", flush=True)
    g.close()
