"""GPU parity for the modes that round 1 left without a GPU-vs-oracle comparison ("untested CUDA paths"):
normalizeVoxels.comp (RGBA16F volumes), injectRadiance.comp with radianceLighting, the Blinn-Phong branch of phong.frag
(cooktorrance = 0), filterRadiance.comp's BOX3 / CUBE kernel modes — and the ABI behaviours the advisor asked for
(overflow is reported, not latched; a texture id can be uploaded twice)."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_lib import Oracle, lib, ptr
from tests.test_gpu_parity import psnr
from vct_b200 import params as P
from vct_b200 import scene as S

pytestmark = pytest.mark.gpu

D, L, SS, W, H = 64, 5, 512, 320, 240


@pytest.fixture(scope="module")
def room():
    from vct_b200.pipeline import Pipeline
    sc = S.room_scene()
    p = S.room_params(W, H)
    o = Oracle(sc, D, L, SS, W, H)
    g = Pipeline(sc, D, L, SS, W, H)
    o.frame(p); g.frame(p)
    yield sc, p, o, g
    g.close()


def test_normalize_voxels_f16_bit_exact(room):
    """shaders/normalizeVoxels.comp:19-42 on RGBA16F accumulators: colour /= count, normal /= count, radiance <- (0,0,0,a),
    VoxelizeInfo counters.  Inputs: fp16 sums of 0..40 fragments per voxel (exactly representable counts), ~6 % occupied."""
    import torch
    sc, p, o, g = room
    rng = np.random.default_rng(11)
    n = D ** 3
    cnt = np.where(rng.random(n) < 0.06, rng.integers(1, 41, n), 0).astype(np.float32)
    col = (rng.random((n, 4), np.float32) * cnt[:, None]).astype(np.float16); col[:, 3] = cnt.astype(np.float16)
    ncnt = np.where(rng.random(n) < 0.5, cnt, 0).astype(np.float32)               # normal volume: its own alpha test (:37-40)
    nrm = (rng.random((n, 4), np.float32) * ncnt[:, None]).astype(np.float16); nrm[:, 3] = ncnt.astype(np.float16)
    for opacity in (0.5, 0.0):
        oc, on = col.copy(), nrm.copy()
        orad = np.zeros(n, np.uint32); info = P.VoxelizeInfo()
        lib().orc_normalize_voxels_f16(D, opacity, ptr(oc.view(np.uint16)), ptr(on.view(np.uint16)), ptr(orad), C.byref(info))
        dc = torch.from_numpy(col.view(np.int16).copy()).cuda(); dn = torch.from_numpy(nrm.view(np.int16).copy()).cuda()
        g.write_volume(P.VOL_RADIANCE, 0, np.zeros(n, np.uint32))
        g._ck(g.lib.vct_normalize_voxels_f16(g.h, dc.data_ptr(), dn.data_ptr(), opacity))
        g.sync()
        assert np.array_equal(dc.cpu().numpy().view(np.uint16), oc.view(np.uint16).reshape(-1, 4)), opacity
        assert np.array_equal(dn.cpu().numpy().view(np.uint16), on.view(np.uint16).reshape(-1, 4)), opacity
        assert np.array_equal(g.read_volume(P.VOL_RADIANCE), orad), opacity
        gi = g.counters()
        assert (gi.unique_voxels, gi.max_fragments_per_voxel) == (info.unique_voxels, info.max_fragments_per_voxel)
        assert info.unique_voxels > 1000 and info.max_fragments_per_voxel == 40
    g.frame(p)


@pytest.mark.parametrize("shadow", [512, 500])           # power-of-two fast path (k_inject) and the generic kernel
def test_radiance_lighting_bit_exact(shadow):
    """injectRadiance.comp:66-95 with radianceLighting: radiance = colour * max(0, N . L_voxelspace) * lightIntensity."""
    from vct_b200.pipeline import Pipeline
    sc = S.room_scene()
    g = Pipeline(sc, D, L, shadow, W, H)
    o = Oracle(sc, D, L, shadow, W, H)
    try:
        for warp in (0, 1):
            p = S.room_params(W, H); p.radiance_lighting = 1; p.warp_voxels = warp
            o.frame(p); g.frame(p)
            rad = g.read_volume(P.VOL_RADIANCE)
            assert ((o.radiance[0] & 0xFFFFFF) != 0).sum() > 300
            assert not np.array_equal(o.radiance[0], o.color[0]), "lighting must change the injected words"
            assert np.array_equal(rad, o.radiance[0]), (shadow, warp, int((rad != o.radiance[0]).sum()))
            for l in range(1, L):
                assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l])
            assert psnr(g.read_image(), o.image) >= 45.0
    finally:
        g.close()


def test_blinn_phong_branch(room):
    """phong.frag:260-301 (cooktorrance = false): Blinn-Phong direct light, specular multiplied by albedo, light.intensity used."""
    sc, p, o, g = room
    q = type(p).from_buffer_copy(p); q.cooktorrance = 0
    o.frame(q); g.frame(q)
    blinn = g.read_image().copy()
    v = psnr(blinn, o.image)
    print("Blinn-Phong PSNR", round(v, 2))
    assert v >= 45.0
    q.enable_indirect = 0                                  # direct light only: nothing but the branch under test (+ ambient)
    o.frame(q); g.frame(q)
    v = psnr(g.read_image(), o.image)
    print("Blinn-Phong, direct only, PSNR", round(v, 2))
    assert v >= 45.0
    o.frame(p); g.frame(p)
    assert psnr(blinn, g.read_image()) < 60.0, "the two BRDF branches must differ visibly"


def test_box3_and_cube_mip_kernels_bit_exact(room):
    """filterRadiance.comp:36-60 behind its `kernelMode` uniform (the reference never sets it): every level of both pyramids."""
    sc, p, o, g = room
    o.frame(p); g.frame(p)
    for mode in (1, 2, 0):
        o.mip("radiance", mode); o.mip("color", mode)
        g.mip_kernel(P.VOL_RADIANCE, mode); g.mip_kernel(P.VOL_COLOR, mode)
        for l in range(1, L):
            assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), (mode, l)
            assert np.array_equal(g.read_volume(P.VOL_COLOR, l), o.color[l]), (mode, l)
        if mode:                                            # the traced texture follows: the image changes with the kernel
            o.shade(p); g.cone_trace(p)
            assert psnr(g.read_image(), o.image) >= 45.0, mode
    from vct_b200.lib import VctError
    with pytest.raises(VctError):
        g.mip_kernel(P.VOL_RADIANCE, 3)
    with pytest.raises(VctError):
        g.mip_kernel(P.VOL_NORMAL, 0)
    g.frame(p)


def test_fragment_buffer_overflow_is_reported_once_and_the_context_recovers():
    """api.cu: the overflow flag is reset with the frame counters and surfaces at the NEXT entry point (mapped host word), once."""
    from vct_b200.lib import VctError
    from vct_b200.pipeline import Pipeline
    sc = S.room_scene()
    p = S.room_params(W, H)
    o = Oracle(sc, D, L, SS, W, H); o.frame(p)
    g = Pipeline(sc, D, L, SS, W, H, max_fragments=1000)   # the room produces ~20 k fragments
    try:
        g.frame(p); g.sync()                                # overflows on the device
        with pytest.raises(VctError, match="overflow"):
            g.frame(p)                                      # reported here, nothing executed
        with pytest.raises(VctError, match="overflow"):
            g.frame(p); g.sync(); g.counters()              # runs (and overflows again): the counters say so
    finally:
        g.close()
    g = Pipeline(sc, D, L, SS, W, H, max_fragments=1 << 16)
    try:
        g.frame(p)
        assert g.counters().total_fragments == o.info.total_fragments
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])
        g.frame(p); g.frame(p)                              # no latch: later frames are clean
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])
    finally:
        g.close()


def test_texture_reupload_replaces_the_texels(room):
    sc, p, o, g = room
    t = sc.textures[0]
    px = t.packed()
    inv = (255 - px).astype(np.uint8)
    g._ck(g.lib.vct_upload_texture(g.h, 0, t.width, t.height, t.channels, min(16, len(t.levels)), inv.ctypes.data))
    g.frame(p)
    changed = g.read_volume(P.VOL_COLOR)
    assert not np.array_equal(changed, o.color[0])
    for _ in range(3):                                      # repeated uploads of one id do not accumulate allocations (cudaFree of the old block)
        g._ck(g.lib.vct_upload_texture(g.h, 0, t.width, t.height, t.channels, min(16, len(t.levels)), px.ctypes.data))
    o.frame(p); g.frame(p)
    assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])


def test_long_per_voxel_lists_without_a_warp_mode():
    """Deterministic running average with many fragments per voxel (a dense mesh on a coarse grid): 60 and 300 coincident quads.  The
    8-bit count wraps at 256 and the average forgets its history there (voxelize.frag:127-133), so only the last (N-1)%256+1 fragments
    in canonical order matter — which is what the long-list paths of voxelize.cu replay.  First frame: the resolve kernel meets the
    long lists without its helper kernels and resolves them itself; later frames: the warp-per-voxel kernel.  All equal the oracle."""
    from vct_b200.pipeline import Pipeline
    rng = np.random.default_rng(3)
    for stack in (60, 300):
        sc = S.Scene()
        texs = [sc.add_texture(S.checker_texture(16, 4, a=tuple(int(x) for x in rng.integers(30, 255, 3)), b=tuple(int(x) for x in rng.integers(30, 255, 3)), seed=k)) for k in range(6)]
        mats = [sc.add_material(diffuse=t) for t in texs]
        parts = [S.quad_mesh([(-0.6, -0.2, 0.5), (0.7, -0.2, 0.5), (0.7, -0.2, -0.6), (-0.6, -0.2, -0.6)], (0, 1, 0), mats[k % 6], 1.0 + 0.1 * (k % 7)) for k in range(stack)]
        parts.append(S.quad_mesh([(-1.4, -1.0, 1.4), (1.4, -1.0, 1.4), (1.4, -1.0, -1.4), (-1.4, -1.0, -1.4)], (0, 1, 0), mats[0], 4.0))
        sc.add_actor(S.merge_meshes(parts))
        sc.lights = [P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1)]
        d, l, ss, w, h = 32, 5, 256, 96, 64
        p = S.room_params(w, h)
        o = Oracle(sc, d, l, ss, w, h)
        g = Pipeline(sc, d, l, ss, w, h)
        try:
            o.frame(p)
            assert o.info.max_fragments_per_voxel == stack % 256 or o.info.max_fragments_per_voxel >= 200, o.info.max_fragments_per_voxel
            for k in range(3):
                g.frame(p)
                assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0]), (stack, k)
                assert np.array_equal(g.read_volume(P.VOL_NORMAL), o.normal), (stack, k)
                assert np.array_equal(g.read_volume(P.VOL_RADIANCE), o.radiance[0]), (stack, k)
                i = g.counters()
                assert (i.total_fragments, i.unique_voxels, i.max_fragments_per_voxel) == (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel), (stack, k)
        finally:
            g.close()


def test_pile_up_beyond_1024_fragments_outside_the_warp_modes():
    """1300 coincident quads: the first frame runs without the long-list kernels, leaves the > 1024-fragment voxels unresolved and says so at the
    next entry point (once, with its own message); from then on the k_voxel_huge_* kernels run and the frame equals the oracle."""
    from vct_b200.lib import VctError
    from vct_b200.pipeline import Pipeline
    sc = S.Scene()
    mat = sc.add_material(diffuse=sc.add_texture(S.checker_texture(16, 4, a=(200, 60, 40), b=(40, 90, 220), seed=1)))
    parts = [S.quad_mesh([(-0.6, -0.2, 0.5), (0.7, -0.2, 0.5), (0.7, -0.2, -0.6), (-0.6, -0.2, -0.6)], (0, 1, 0), mat, 1.0 + 0.1 * (k % 7)) for k in range(1300)]
    parts.append(S.quad_mesh([(-1.4, -1.0, 1.4), (1.4, -1.0, 1.4), (1.4, -1.0, -1.4), (-1.4, -1.0, -1.4)], (0, 1, 0), mat, 4.0))
    sc.add_actor(S.merge_meshes(parts))
    sc.lights = [P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1)]
    d, l, ss, w, h = 32, 5, 256, 96, 64
    p = S.room_params(w, h)
    o = Oracle(sc, d, l, ss, w, h); o.frame(p)
    assert o.info.total_fragments > 150000 and o.info.max_fragments_per_voxel == 1300 % 256   # the counter is the word's 8-bit running count
    g = Pipeline(sc, d, l, ss, w, h)
    try:
        g.frame(p); g.sync()
        with pytest.raises(VctError, match="1024 fragments"):
            g.frame(p)
        for k in range(2):
            g.frame(p)
            assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0]), k
            assert np.array_equal(g.read_volume(P.VOL_NORMAL), o.normal), k
            assert np.array_equal(g.read_volume(P.VOL_RADIANCE), o.radiance[0]), k
            i = g.counters()
            assert (i.total_fragments, i.unique_voxels, i.max_fragments_per_voxel) == (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel), k
    finally:
        g.close()


def test_msaa_voxelisation_bit_exact(room):
    """Settings::conservativeRasterization == MSAA (the reference's default raster mode, Application.cpp:244-249, 673-678): any-sample coverage
    of the 4x pattern, inputs at the pixel centre.  Occupancy grid, volumes, every pyramid level and the counters equal the oracle's, in the
    deterministic and the atomic-max mode, with the warp map on top, on a dense and on a sparse frame; a pattern with every sample at the
    pixel centre reproduces centre sampling."""
    sc, p, o, g = room
    o.frame(p); g.frame(p)
    off = g.read_volume(P.VOL_COLOR).copy(); off_frags = g.counters().total_fragments
    for mods in ({}, {"voxelize_atomic_max": 1}, {"warp_texture": 1, "temporal_filter_radiance": 1}):
        q = type(p).from_buffer_copy(p); q.conservative_raster = P.RASTER_MSAA
        for k, v in mods.items():
            setattr(q, k, v)
        for frame in range(2):                                  # second frame: sparse
            o.frame(q); g.frame(q)
            if q.warp_texture:
                assert np.array_equal(g.read_volume(P.VOL_OCCUPANCY), o.occ), (mods, frame)
            assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0]), (mods, frame)
            assert np.array_equal(g.read_volume(P.VOL_NORMAL), o.normal), (mods, frame)
            for l in range(L):
                assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), (mods, frame, l)
            i = g.counters()
            assert (i.total_fragments, i.unique_voxels, i.max_fragments_per_voxel) == (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel), (mods, frame)
            assert psnr(g.read_image(), o.image) >= 45.0
        if not mods:
            assert g.counters().total_fragments > off_frags * 1.05
    q = type(p).from_buffer_copy(p); q.conservative_raster = P.RASTER_MSAA
    q.msaa_samples = (C.c_float * 8)(*([0.5] * 8))
    g.frame(q)
    assert g.counters().total_fragments == off_frags and np.array_equal(g.read_volume(P.VOL_COLOR), off)
    from vct_b200.lib import VctError
    q.conservative_raster = 2                                   # GL_CONSERVATIVE_RASTERIZATION_NV: not built, refused
    with pytest.raises(VctError):
        g.frame(q)
    g.frame(p)


def test_voxelize_multiplier_bit_exact(room):
    """Settings::voxelizeMultiplier (Application.cpp:668): the voxelise pass in a viewport of (int)(m * dim) pixels squared — more (or fewer)
    fragments per voxel, canonical order by (triangle, raster rank in THAT viewport), texture LOD from that viewport's derivatives.  Volumes,
    pyramid and counters equal the oracle's for m = 2, 1.5 and 0.5, with and without multisampling, dense and sparse frame."""
    sc, p, o, g = room
    for m, msaa in ((2.0, 0), (1.5, 1), (0.5, 0)):
        q = type(p).from_buffer_copy(p); q.voxelize_multiplier = m; q.conservative_raster = msaa
        for frame in range(2):
            o.frame(q); g.frame(q)
            assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0]), (m, frame)
            assert np.array_equal(g.read_volume(P.VOL_NORMAL), o.normal), (m, frame)
            for l in range(L):
                assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), (m, frame, l)
            i = g.counters()
            assert (i.total_fragments, i.unique_voxels, i.max_fragments_per_voxel) == (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel), (m, frame)
    from vct_b200.lib import VctError
    q = type(p).from_buffer_copy(p); q.voxelize_multiplier = -1.0
    with pytest.raises(VctError):
        g.frame(q)
    g.frame(p)
