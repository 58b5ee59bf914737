"""Settings::voxelizeTesselationWarp through the C ABI (the third entry of common.glsl's mapping priority):
the voxel grid becomes the camera frustum, position = (pv * P).xyz / w * 0.5 + 0.5 (common.glsl:37-42).  It is read by
testTesselation.tese, injectRadiance.comp and phong.frag (every cone sample goes back to world space and through pv,
phong.frag:158-162; the voxel view, :348); voxelize.frag declares the uniform and never reads it.  The oracle's version is
pinned to those shaders compiled as C++ (tests/test_glsl_ref.py: inject_tess_warp*, tess_max_tess_warp, shade_tess_warp*,
shade_view_voxels_tess_warp); here the CUDA kernels are compared with the oracle.

Hardware status: written after this round's GPU budget was nearly spent.  The last 28 s of it went into tools/hw_smoke.sh — the
same configurations through the C++ host on one B200 (profiles/r01s_hw_smoke_tess_warp.txt): images 61-71 dB against the oracle, the
voxel view in the warped grid identical to the last bit, VoxelizeInfo counters equal.  This file (volumes bit for bit, through the
Python mirror) has not itself run on a GPU yet; it sorts after the other GPU tests so that `pytest -x` reaches them first."""
import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.test_glsl_ref import pbr_room
from tests.test_gpu_parity import psnr
from vct_b200 import params as P
from vct_b200 import scene as S

pytestmark = pytest.mark.gpu

D, L, SS, W, H = 64, 5, 512, 320, 240


def _compare_frame(g, o, p, what):
    o.frame(p)
    g.frame(p)
    for l in range(L):
        assert np.array_equal(g.read_volume(P.VOL_COLOR, l), o.color[l]), (what, "colour", l)
        assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), (what, "radiance", l)
    q = psnr(g.read_image(), o.image)
    print(what, "PSNR", round(q, 2))
    assert q >= 45.0, (what, q)                                      # the final image's gate


def test_tessellation_warp_frames_match_the_oracle():
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    o = Oracle(sc, D, L, SS, W, H)
    # the reference's default voxeliser in the frustum-aligned grid: tese store, inject and every cone sample go through pv
    p = S.room_params(W, H); p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1; p.voxelize_tesselation_warp = 1
    _compare_frame(g, o, p, "tessellation voxeliser + tess warp")
    assert ((o.color[0] >> 24) != 0).sum() > 500 and (o.radiance[0] != 0).sum() > 100
    _compare_frame(g, o, p, "the same frame again (sparse path)")
    # the raster voxeliser ignores the flag (voxelize.frag never reads it); inject and the cone trace do not
    p = S.room_params(W, H); p.voxelize_tesselation_warp = 1
    _compare_frame(g, o, p, "raster voxeliser + tess warp")
    q = S.room_params(W, H)
    o2 = Oracle(sc, D, L, SS, W, H); o2.shadowmap(q); o2.voxelize(q)
    assert np.array_equal(o.color[0] >> 24 != 0, o2.color[0] >> 24 != 0)
    # priority: warpVoxels wins over the tessellation warp everywhere (common.glsl:44-60, phong.frag:150-162)
    p = S.room_params(W, H); p.voxelize_tesselation_warp = 1; p.warp_voxels = 1
    _compare_frame(g, o, p, "warp_voxels + tess warp")
    # both of the other warps on (no tessellation warp): the voxeliser, inject and the voxel view take warpVoxels first, traceCone takes
    # warpTexture first (phong.frag:150-157) — the dispatch of k_cone_trace follows that since this change
    p = S.room_params(W, H); p.warp_texture = 1; p.warp_voxels = 1
    _compare_frame(g, o, p, "warp_texture + warp_voxels")
    # back to the plain frame: nothing left over
    p = S.room_params(W, H)
    _compare_frame(g, o, p, "plain frame afterwards")
    g.close()


def test_voxel_view_in_the_tessellation_warp():
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    o = Oracle(sc, D, L, SS, W, H)
    for lod in (0.0, 1.3):
        p = S.room_params(W, H); p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1; p.voxelize_tesselation_warp = 1
        p.debug_view = P.VIEW_VOXELS; p.miplevel = lod
        o.frame(p); g.frame(p)
        q = psnr(g.read_image(), o.image)
        print("voxel view, lod", lod, "PSNR", round(q, 2))
        assert q >= 45.0, (lod, q)
    g.close()
