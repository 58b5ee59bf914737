"""The reference's DEFAULT voxeliser through the C ABI: vct_frame_params::voxelize_tesselation routes
vct_voxelize / vct_frame through testTesselation.tesc/.tese semantics (src/Application.cpp:585-665) instead of the raster path.
The oracle's version is pinned to the reference's two shaders compiled as C++ (tests/test_glsl_ref.py); here the CUDA kernel
is compared with the oracle: atomicMax mode (the reference default) bit for bit, the running average by occupancy."""
import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.test_glsl_ref import pbr_room
from tests.test_gpu_parity import max_byte_delta, psnr
from vct_b200 import params as P
from vct_b200 import scene as S

pytestmark = pytest.mark.gpu

D, L, SS, W, H = 64, 5, 512, 320, 240


def test_tessellation_voxeliser_matches_the_oracle():
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    # ---- the pass alone: atomicMax (bit-exact), running average (occupancy + counts; colour order-dependent like the GLSL)
    p = S.room_params(W, H); p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1
    o = Oracle(sc, D, L, SS, W, H)
    o.shadowmap(p); o.voxelize(p)
    g.shadowmap(p); g.voxelize(p)
    col, nrm = g.read_volume(P.VOL_COLOR), g.read_volume(P.VOL_NORMAL)
    assert ((o.color[0] >> 24) != 0).sum() > 3000
    assert np.array_equal(col, o.color[0]) and np.array_equal(nrm, o.normal), f"{(col != o.color[0]).sum()} colour / {(nrm != o.normal).sum()} normal words differ"
    assert g.counters().total_fragments == 0                             # the evaluation shader counts nothing
    p.voxelize_atomic_max = 0
    o.voxelize(p); g.voxelize(p)
    col = g.read_volume(P.VOL_COLOR)
    assert np.array_equal(col >> 24, o.color[0] >> 24), "points per voxel must agree (order-independent)"
    print("tess avg: max byte delta", max_byte_delta(col, o.color[0]))
    assert max_byte_delta(col & 0xFFFFFF, o.color[0] & 0xFFFFFF) <= 8    # truncating average, arbitrary insertion order
    # ---- the reference's default frame (tessellation + atomicMax), twice: the second frame takes the sparse path
    p = S.room_params(W, H); p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1
    o.frame(p)
    for it in range(2):
        g.frame(p)
        assert g.frame_was_sparse() == (it == 1)
        for l in range(L):
            assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]) and np.array_equal(g.read_volume(P.VOL_COLOR, l), o.color[l]), (it, l)
        q = psnr(g.read_image(), o.image)
        print("tess frame", it, "PSNR", round(q, 2))
        assert q >= 45.0
    # ---- back to the raster path: nothing left over
    p = S.room_params(W, H)
    o2 = Oracle(sc, D, L, SS, W, H); o2.frame(p)
    g.frame(p)
    assert np.array_equal(g.read_volume(P.VOL_COLOR), o2.color[0]) and np.array_equal(g.read_volume(P.VOL_RADIANCE), o2.radiance[0])
    g.close()
