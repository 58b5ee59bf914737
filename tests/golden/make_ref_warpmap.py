#!/usr/bin/env python3
"""Regenerates tests/golden/warpmap_cpu_ref.npz from the REFERENCE'S OWN code: oracle/_ref/warpmap_cpu is
/root/reference/src/Application.cpp:311-370 (partial-sum tables + low/high weight table of the per-frame warp-map
generation) compiled in place by oracle/Makefile.  Inputs: seeded 32^3 occupancy grids (PCG64 seed 0xA9) at three
densities plus the all-empty / all-full edge cases; (high, low) = the reference defaults (2.0, 0.5) and (3.0, 0.25)."""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RIG = os.path.join(HERE, "..", "..", "oracle", "_ref", "warpmap_cpu")
N = 32


def cases():
    rng = np.random.default_rng(0xA9)
    occ = [(rng.random(N ** 3) < d).astype(np.uint32) for d in (0.03, 0.15, 0.6)]
    occ += [np.zeros(N ** 3, np.uint32), np.ones(N ** 3, np.uint32)]
    occ[1][::7] *= 5                      # any non-zero count is "occupied" (imageAtomicOr writes 1; be general)
    return occ, [(2.0, 0.5), (3.0, 0.25)]


def run(occ, high, low):
    with tempfile.NamedTemporaryFile(suffix=".u32") as f:
        occ.astype(np.uint32).tofile(f.name)
        raw = subprocess.run([RIG, f.name, repr(high), repr(low)], check=True, capture_output=True).stdout
    part = np.frombuffer(raw[: N ** 3 * 12], np.int32).reshape(N ** 3, 3)
    w = np.frombuffer(raw[N ** 3 * 12:], np.float32).reshape(2, N + 1)
    return part.copy(), w.copy()


if __name__ == "__main__":
    occ, hl = cases()
    out = {}
    for i, o in enumerate(occ):
        part, _ = run(o, *hl[0])
        out[f"occ{i}"] = np.packbits(o > 0)
        out[f"part{i}"] = part.astype(np.uint8)              # counts <= 32
    for j, (h, l) in enumerate(hl):
        out[f"weights{j}"] = run(occ[0], h, l)[1]
    np.savez_compressed(os.path.join(HERE, "warpmap_cpu_ref.npz"), **out)
    print({k: v.shape for k, v in out.items()})
