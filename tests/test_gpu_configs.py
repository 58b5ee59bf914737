"""BASELINE.json configurations 1, 2, 4 and 5 through the product (C ABI) against the CPU oracle — config 3 is
tests/test_gpu_parity.py::test_sponza_256_1080p_full_size_gates.  Gates (north star): occupancy mask bit-exact; RGBA8 volumes
max |delta| <= 2/255 (this build: bit-exact in the deterministic mode, and the tests say so); final image PSNR >= 45 dB.
Grid sizes are the configurations' own; configs 4 and 5 shade a reduced frame (the oracle's cone trace at 4K takes minutes) —
the full 4K frame is bench.py's (`--config 4|5`)."""
import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.test_gpu_parity import max_byte_delta, psnr
from vct_b200 import params as P
from vct_b200 import scene as S
from vct_b200.workloads import Workload

pytestmark = pytest.mark.gpu


def _gates(g, o, L, what, image=True):
    col, rad, nrm = g.read_volume(P.VOL_COLOR), g.read_volume(P.VOL_RADIANCE), g.read_volume(P.VOL_NORMAL)
    assert np.array_equal((col >> 24) != 0, (o.color[0] >> 24) != 0), (what, "occupancy mask must be bit-exact")
    info = g.counters()
    assert (info.total_fragments, info.unique_voxels, info.max_fragments_per_voxel) == \
        (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel), what
    d = max(max_byte_delta(col, o.color[0]), max_byte_delta(rad, o.radiance[0]), max_byte_delta(nrm, o.normal))
    assert d <= 2, (what, d)
    assert np.array_equal(col, o.color[0]) and np.array_equal(rad, o.radiance[0]) and np.array_equal(nrm, o.normal), (what, "deterministic mode is bit-exact")
    for l in range(1, L):
        assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), (what, "radiance level", l)
        assert np.array_equal(g.read_volume(P.VOL_COLOR, l), o.color[l]), (what, "colour level", l)
    q = None
    if image:
        assert np.array_equal(g.read_visibility(), o.vis), (what, "visibility buffer")
        q = psnr(g.read_image(), o.image)
        assert q >= 45.0, (what, q)
    print(f"{what}: {info.total_fragments} fragments, {info.unique_voxels} voxels, volumes max byte delta {d}, image PSNR {q if q is None else round(q, 2)} dB")
    return info


def test_config1_cube_64():
    """resources/cube.obj, one shadow-mapped light, 64^3, 512x512 (the configuration the reference runs headless on llvmpipe)."""
    from vct_b200.pipeline import Pipeline
    w = Workload(1)
    assert (w.D, w.W, w.H, w.L, w.S) == (64, 512, 512, 6, 4096) and w.scene.n_tris == 12
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H)
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    try:
        for k in range(2):                                   # dense first frame, sparse second
            o.frame(w.params); g.frame(w.params)
            info = _gates(g, o, w.L, f"config 1 frame {k}")
        assert info.unique_voxels > 1500
        assert np.array_equal(g.read_shadowmap().view(np.uint32), o.shadow.view(np.uint32))
    finally:
        g.close()


@pytest.mark.skipif(not S.baked_available("bunny"), reason="assets/_baked/bunny missing (tools/bake_assets.py)")
def test_config2_bunny_128_occupancy_bit_exact():
    """resources/bunny.obj has no UVs: tangents are NaN (Mesh.cpp:178,196), every diffuse cone direction is NaN and breaks at once
    (zero diffuse indirect); enableReflections = false.  The gate that means something here is the occupancy mask."""
    from vct_b200.pipeline import Pipeline
    w = Workload(2)
    assert (w.D, w.W, w.H) == (128, 1280, 720) and w.scene.n_tris == 4968 and not w.params.enable_reflections
    assert np.isnan(w.scene.meshes[0].vertices[:, 8:14]).all(), "the bunny's tangent frames are NaN in the reference"
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H)
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    try:
        for k in range(2):
            o.frame(w.params); g.frame(w.params)
            info = _gates(g, o, w.L, f"config 2 frame {k}")
        assert info.unique_voxels > 10000
        assert g.cone_steps() == o.cone_steps == 0, "NaN cone directions take no step; reflections are off"
        # atomic-max mode and the free-running CAS loop keep the same occupancy
        for mode in ("max", "cas"):
            q = type(w.params).from_buffer_copy(w.params)
            if mode == "max": q.voxelize_atomic_max = 1
            else: q.deterministic = 0
            g.voxelize(q)
            assert np.array_equal((g.read_volume(P.VOL_COLOR) >> 24) != 0, (o.color[0] >> 24) != 0), mode
    finally:
        g.close()


@pytest.mark.skipif(not (S.baked_available("sponza_pbr") and S.baked_available("nanosuit")), reason="baked Sponza / nanosuit missing")
@pytest.mark.timeout(600)
def test_config4_animated_warped_temporal_512():
    """Sponza + two animated nanosuits (Application.cpp:97-116), 512^3 grid, warp map regenerated every frame (:235-577),
    temporal radiance filter (transferVoxels.comp:55-62): four frames at 60 Hz, every pyramid level bit for bit; 480x270 frame."""
    from vct_b200.pipeline import Pipeline
    w = Workload(4, width=480, height=270)
    assert w.D == 512 and w.animated and w.params.warp_texture and w.params.temporal_filter_radiance
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H)
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    try:
        prev = None
        for k in range(4):
            for actor, model in w.models(k * 20):            # 1/3 s apart: the actors visibly move between frames
                o.set_actor_transform(actor, model); g.set_actor_transform(actor, model)
            o.frame(w.params); g.frame(w.params)
            assert np.array_equal(g.read_volume(P.VOL_OCCUPANCY), o.occ), k
            assert np.array_equal(g.read_volume(P.VOL_WARPMAP), o.warpmap), k
            info = _gates(g, o, w.L, f"config 4 frame {k}")
            if prev is not None:
                assert not np.array_equal(prev, o.color[0]), "the animation must change the voxelisation"
            prev = o.color[0].copy()
        assert info.unique_voxels > 1000000
        hist = (o.radiance[0] != 0) & (o.color[0] == 0)
        assert hist.sum() > 100, "the temporal filter must leave decaying radiance where the actors were"
    finally:
        g.close()


@pytest.mark.timeout(400)
def test_config5_soup_1m_triangles_512():
    """Synthetic soup, 1 Mi triangles (sigma = 1.5 voxels at 512^3), NaN tangents, 1x1 white texture: volumes bit for bit."""
    from vct_b200.pipeline import Pipeline
    w = Workload(5, width=480, height=270)
    assert w.D == 512 and w.scene.n_tris == 1 << 20
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H)
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    try:
        for k in range(2):
            o.frame(w.params); g.frame(w.params)
            info = _gates(g, o, w.L, f"config 5 frame {k}")
        assert info.total_fragments > 500000
    finally:
        g.close()


@pytest.mark.timeout(600)
def test_config5_soup_capacity_8m_triangles():
    """Capacity: 8 Mi triangles into 512^3 (queues, setups and the fragment buffer sized from the triangle count), no overflow;
    counters equal to the oracle's voxeliser.  (The 64 Mi-triangle run is tools/soup_capacity.py -> profiles/.)"""
    from vct_b200.pipeline import Pipeline
    w = Workload(5, width=64, height=48, triangles=8 << 20)
    g = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H, max_fragments=3 * w.scene.n_tris)
    o = Oracle(w.scene, w.D, w.L, w.S, w.W, w.H)
    try:
        g.frame(w.params); g.sync()
        info = g.counters()                                  # raises if any fixed-capacity buffer overflowed
        o.shadowmap(w.params); o.voxelize(w.params); o.transfer(w.params)
        assert (info.total_fragments, info.unique_voxels, info.max_fragments_per_voxel) == \
            (o.info.total_fragments, o.info.unique_voxels, o.info.max_fragments_per_voxel)
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])
        print(f"soup 8 Mi triangles: {info.total_fragments} fragments, {info.unique_voxels} voxels, max {info.max_fragments_per_voxel} per voxel")
        # a fragment buffer that is too small is reported, not silently truncated
        g2 = Pipeline(w.scene, w.D, w.L, w.S, w.W, w.H, max_fragments=1 << 20)
        try:
            from vct_b200.lib import VctError
            g2.frame(w.params); g2.sync()
            with pytest.raises(VctError, match="overflow"):
                g2.counters()
        finally:
            g2.close()
    finally:
        g.close()
