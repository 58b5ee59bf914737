#!/usr/bin/env python3
"""Writes tests/golden/glsl_ref.npz, glsl_ref_fragment.npz, glsl_ref_warpmap.npz, glsl_ref_tess.npz, glsl_ref_alpha.npz and glsl_ref_debug_voxels.npz: the outputs of the reference's GLSL shaders, compiled as C++ into
oracle/_ref/libvct_glsl_ref.so (oracle/Makefile; needs /root/reference), on the seeded inputs of tests/test_glsl_ref.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import test_glsl_ref as T  # noqa: E402

if __name__ == "__main__":
    out = T.run_cases("glsl")
    np.savez_compressed(T.GOLD, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD), "bytes")
    out = T.run_fragment_cases("glsl")
    np.savez_compressed(T.GOLD_FRAG, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD_FRAG), "bytes")
    out = T.run_warpmap_cases("glsl")
    np.savez_compressed(T.GOLD_WARP, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD_WARP), "bytes")
    out = T.run_tess_cases("glsl")
    np.savez_compressed(T.GOLD_TESS, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD_TESS), "bytes")
    out = T.run_alpha_cases("glsl")
    np.savez_compressed(T.GOLD_ALPHA, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD_ALPHA), "bytes")
    out = T.run_debug_voxel_cases("glsl")
    np.savez_compressed(T.GOLD_DBGVOX, **out)
    print(len(out), "arrays,", os.path.getsize(T.GOLD_DBGVOX), "bytes")
