#!/usr/bin/env python3
"""Config 5 capacity sweep (BASELINE.json: "synthetic triangle-soup scaling sweep 1M-64M triangles into 512^3 grid"): for each
triangle count build the soup, run whole frames through the C ABI, check that no fixed-capacity buffer overflowed
(vct_get_counters fails if one did) and print fragments, voxels, per-pass times and memory.  The host needs ~0.6 GB of
RAM per Mi triangles (56-byte vertices x 3, twice: numpy + the library's copy); counts that do not fit are skipped with a note.
usage: soup_capacity.py [Mi triangles ...]   (default 1 4 16 64)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vct_b200.pipeline import Pipeline
from vct_b200.workloads import Workload


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            return int(line.split()[1]) / 1e6
    return 0.0


counts = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64]
for mi in counts:
    need = 0.9 * mi + 4
    if mem_available_gb() < need:
        print(json.dumps({"triangles_mi": mi, "skipped": f"host has {mem_available_gb():.0f} GB available, needs ~{need:.0f} GB"}), flush=True)
        continue
    t0 = time.time()
    w = Workload(5, width=3840, height=2160, triangles=mi << 20)
    t_build = time.time() - t0
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H, max_fragments=min(3 * w.scene.n_tris, (1 << 31) - 1))
    w.scene.meshes[0].vertices = None                       # the library holds its own copy now
    try:
        g.frame(w.params); g.sync()
        g.set_profiling(1)
        g.frame(w.params); g.sync()
        t = {k.replace("_ns", "_ms"): round(v / 1e6, 3) for k, v in g.timings().items() if v > 0}
        info = g.counters()                                 # raises VctError on overflow
        import torch
        free, total = torch.cuda.mem_get_info()
        print(json.dumps({"triangles_mi": mi, "triangles": w.scene.n_tris, "fragments": info.total_fragments, "unique_voxels": info.unique_voxels,
                          "max_fragments_per_voxel": info.max_fragments_per_voxel, "overflow": False, "passes": t,
                          "device_gb_used": round((total - free) / 1e9, 1), "host_build_s": round(t_build, 1)}), flush=True)
    except Exception as e:
        print(json.dumps({"triangles_mi": mi, "error": f"{type(e).__name__}: {e}"}), flush=True)
    finally:
        g.close()
    del w, g
