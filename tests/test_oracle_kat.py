"""Known-answer tests that pin the CPU oracle (CPU only, no GPU).

The reference ships no tests and no golden vectors, so the pins are the values derived by hand from
the reference's own arithmetic: the `#if 0` warp rig of src/main.cpp:20-127, imageAtomicRGBA8Avg of
shaders/voxelize.frag:111-139, the weight table of src/Application.cpp:346-370, and closed forms of
filterRadiance.comp / traceCone.  tests/golden/kat.json holds them; tests/golden/make_kat.py is an independent
pure-Python restatement of the same reference lines that regenerates the file.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from tests import oracle_lib as ol
from vct_b200 import params as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


@pytest.fixture(scope="module")
def L():
    return ol.lib()


def insert_seq(L, vals, stored=0):
    words = []
    for r, g, b in vals:
        stored = L.orc_rgba8_avg(stored, r, g, b)
        words.append(stored)
    return words


# ------------------------------------------------------------------ imageAtomicRGBA8Avg, voxelize.frag:111-139
def test_running_average_sequence(L):
    words = insert_seq(L, [(1.0, 0.5, 0.25), (0.0, 0.0, 0.0), (0.2, 0.4, 0.6)])
    assert [f"0x{w:08X}" for w in words] == GOLD["rgba8_avg"]["abc_words"]
    assert words[-1] == 0x03474C65               # bytes 101,76,71 count 3: truncation bias vs the true mean 102,76.5,72.25


def test_running_average_is_order_dependent(L):
    for order, expect in GOLD["rgba8_avg"]["order_dependence"]:
        w = insert_seq(L, [(v, 0.0, 0.0) for v in order])[-1]
        assert (w & 255) == expect, order


def test_running_average_count_wraps_at_256(L):
    words = insert_seq(L, [(0.5, 0.5, 0.5)] * 257)
    assert words[0] == 0x017F7F7F and words[254] == 0xFF7F7F7F
    assert words[255] == 0x007F7F7F              # 256 fragments: count wrapped, the voxel reads as EMPTY
    assert words[256] == 0x017F7F7F
    assert [f"0x{words[i]:08X}" for i in (0, 254, 255, 256)] == GOLD["rgba8_avg"]["grey_wrap"]


def test_running_average_matches_independent_restatement(L):
    rng = np.random.default_rng(7)
    from tests.golden.make_kat import rgba8_avg_py
    for _ in range(200):
        vals = rng.random((int(rng.integers(1, 40)), 3)).astype(np.float32)
        w = wp = 0
        for r, g, b in vals:
            w = L.orc_rgba8_avg(w, float(r), float(g), float(b)); wp = rgba8_avg_py(wp, r, g, b)
            assert w == wp


def test_atomic_max_mode_words(L):
    ws = [L.orc_pack_unorm4x8(*c, 1.0) for c in [(1.0, 0.5, 0.25), (0.0, 0.0, 0.0), (0.2, 0.4, 0.6)]]
    assert [f"0x{w:08X}" for w in ws] == GOLD["atomic_max"]["words"]
    assert max(ws) == 0xFF996633                 # lexicographic B,G,R winner: C wins on blue


def test_store_rounding_tie(L):
    assert (L.orc_pack_unorm4x8(0, 0, 0, 0.5) >> 24) == 128      # 127.5 -> 128 (round half to even)
    assert (L.orc_pack_unorm4x8(0, 0, 0, 2.5 / 255.0) >> 24) == 2
    assert L.orc_pack_unorm4x8(-1.0, 2.0, float("nan"), 1.0) == 0xFF00FF00


# ------------------------------------------------------------------------ warp weights, Application.cpp:346-370
def test_production_weight_table(L):
    lo = np.zeros(33, np.float32); hi = np.zeros(33, np.float32)
    L.orc_warp_weight_table(32, 2.0, 0.5, ol.ptr(lo), ol.ptr(hi))
    for occ, (l, h) in GOLD["weight_table_32"].items():
        assert lo[int(occ)] == pytest.approx(l, abs=1e-6) and hi[int(occ)] == pytest.approx(h, abs=1e-6), occ
    # every row keeps its length: l*empty + h*occupied == 32
    for k in range(1, 32):
        assert lo[k] * (32 - k) + hi[k] * k == pytest.approx(32.0, abs=1e-4)


# -------------------------------------------------------------------------- warp rig, src/main.cpp:20-127
def run_rig(L, cells, tcs, fixed_low=0.5):
    n = cells.shape[0]
    tc = np.ascontiguousarray(tcs, np.float32); out = np.zeros_like(tc)
    px = np.zeros((n, n), np.int32); py = np.zeros((n, n), np.int32)
    wl = np.zeros(n + 1, np.float32); wh = np.zeros(n + 1, np.float32)
    L.orc_warp_rig(n, ol.ptr(np.ascontiguousarray(cells, np.float32)), C.c_float(fixed_low), C.c_float(2.0), C.c_float(0.5),
                   ol.ptr(tc), len(tc), ol.ptr(out), ol.ptr(px), ol.ptr(py), ol.ptr(wl), ol.ptr(wh))
    return out, px, py, wl, wh


def test_warp_rig_known_answers(L):
    g = GOLD["warp_rig"]
    cells = np.array(g["cells"], np.float32)
    tcs = np.array([t for t, _ in g["points"]], np.float32)
    out, px, py, wl, wh = run_rig(L, cells, tcs)
    assert px.tolist() == g["partials_x"] and py.tolist() == g["partials_y"]
    assert np.allclose(wl, g["weights_low"], atol=1e-6) and np.allclose(wh, g["weights_high"], atol=1e-6)
    for (t, w), o in zip(g["points"], out):
        assert o == pytest.approx(w, abs=2e-6), t


def _check_rig_against(L, ref):
    n = ref["dim"]
    cells = np.array(ref["cells"], np.float32).reshape(n, n)
    w = np.array(ref["warp"], np.float32)
    out, px, py, wl, wh = run_rig(L, cells, np.ascontiguousarray(w[:, :2]))
    part = np.array(ref["partials"], np.int32).reshape(n, n, 2)
    assert np.array_equal(px, part[:, :, 0]) and np.array_equal(py, part[:, :, 1])
    wt = np.array(ref["weights"], np.float32)
    assert np.array_equal(wl, wt[:, 0]) and np.array_equal(wh, wt[:, 1])
    assert np.array_equal(out.view(np.uint32), np.ascontiguousarray(w[:, 2:]).view(np.uint32)), \
        f"max |delta| {np.abs(out - w[:, 2:]).max()} vs the reference's own warp code"


def test_warp_rig_matches_reference_run_fixture(L):
    """tests/golden/warp_rig_ref.json holds outputs of the reference's OWN C++ (src/main.cpp:21-128 compiled in place,
    see tests/golden/make_ref_rig.py): partial sums, weights and 576 warped texcoords must match BIT FOR BIT."""
    _check_rig_against(L, json.load(open(os.path.join(os.path.dirname(__file__), "golden", "warp_rig_ref.json"))))


def test_warp_rig_matches_reference_binary(L):
    """Same check against a fresh run of oracle/_ref/warp_rig on a denser grid (where the binary was built)."""
    rig = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "warp_rig")
    if not os.path.isfile(rig):
        pytest.skip("oracle/_ref/warp_rig not built (needs /root/reference at build time)")
    from tests.golden import make_ref_rig
    _check_rig_against(L, make_ref_rig.run(61))


def _check_warpmap_tables(L, occ, part_ref, hl, weights_ref):
    n = P.WARP_DIM if hasattr(P, "WARP_DIM") else 32
    L.orc_warp_partials.argtypes = [C.c_void_p, C.c_void_p]
    for o, pr in zip(occ, part_ref):
        got = np.zeros((n ** 3, 3), np.int32)
        L.orc_warp_partials(ol.ptr(np.ascontiguousarray(o, np.uint32)), ol.ptr(got))
        assert np.array_equal(got, pr.astype(np.int32))
    for (h, l), wr in zip(hl, weights_ref):
        lo = np.zeros(n + 1, np.float32); hi = np.zeros(n + 1, np.float32)
        L.orc_warp_weight_table(n, h, l, ol.ptr(lo), ol.ptr(hi))
        assert np.array_equal(lo.view(np.uint32), wr[0].view(np.uint32)) and np.array_equal(hi.view(np.uint32), wr[1].view(np.uint32))


def test_warpmap_cpu_tables_match_reference_run_fixture(L):
    """tests/golden/warpmap_cpu_ref.npz holds outputs of the reference's OWN C++ (src/Application.cpp:311-370 compiled
    in place): per-axis partial sums of five 32^3 occupancy grids and two weight tables, matched bit for bit."""
    from tests.golden import make_ref_warpmap as M
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "warpmap_cpu_ref.npz"))
    occ, hl = M.cases()
    for i, o in enumerate(occ):
        assert np.array_equal(np.packbits(o > 0), z[f"occ{i}"]), "fixture inputs drifted from make_ref_warpmap.cases()"
    _check_warpmap_tables(L, occ, [z[f"part{i}"] for i in range(len(occ))], hl, [z[f"weights{j}"] for j in range(len(hl))])


def test_warpmap_cpu_tables_match_reference_binary(L):
    if not os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "warpmap_cpu")):
        pytest.skip("oracle/_ref/warpmap_cpu not built (needs /root/reference at build time)")
    from tests.golden import make_ref_warpmap as M
    rng = np.random.default_rng(77)
    occ = [(rng.random(32 ** 3) < d).astype(np.uint32) for d in (0.01, 0.3, 0.9)]
    hl = [(2.0, 0.5), (1.5, 0.75), (4.0, 0.1)]
    _check_warpmap_tables(L, occ, [M.run(o, 2.0, 0.5)[0] for o in occ], hl, [M.run(occ[0], h, l)[1] for h, l in hl])


def test_warp_rig_is_a_bijection_per_row(L):
    cells = np.array(GOLD["warp_rig"]["cells"], np.float32)
    for row in range(4):
        y = (row + 0.5) / 4
        xs = np.linspace(0, 0.9999, 201, dtype=np.float32)
        out, *_ = run_rig(L, cells, np.stack([xs, np.full_like(xs, y)], 1))
        assert out[0, 0] == 0.0 and np.all(np.diff(out[:, 0]) > 0) and out[-1, 0] < 1.0 + 1e-6
        assert out[-1, 0] == pytest.approx(1.0, abs=1e-3)


def test_warpmap_32_quad_quirk_and_packing(L):
    """generateWarpmap via the 0.8-scaled quad (quad.vert:12): only texels 3..28 in x,y are written."""
    rng = np.random.default_rng(3)
    occ = (rng.random(32 ** 3) < 0.15).astype(np.uint32)
    p = P.FrameParams(); p.warp_texture = 1; p.use_warpmap_weights_texture = 1
    p.warp_texture_axes = (C.c_int * 3)(1, 1, 1); p.warp_texture_high_resolution = 2.0; p.warp_texture_low_resolution = 0.5
    wm = np.zeros(32 ** 3 * 4, np.uint16); lo = np.zeros_like(wm); hi = np.zeros_like(wm)
    L.orc_warpmap(ol.ptr(occ), C.byref(p), ol.ptr(wm), ol.ptr(lo), ol.ptr(hi))
    wm = wm.reshape(32, 32, 32, 4)            # z, y, x
    written = np.zeros(32, bool); written[3:29] = True
    assert not wm[:, ~written, :, :].any() and not wm[:, :, ~written, :].any()
    assert wm[:, 3:29, 3:29, :3].any()
    # alpha packs (totalX | totalY<<5 | totalZ<<10 | occupied<<15)/65535 -> exact integer after unorm16 rounding
    z, y, x = 7, 10, 12
    tc = [((i + 0.5) / 32 - 0.5) / 0.8 + 0.5 for i in (x, y)] + [(z + 0.5) / 32]
    cx, cy, cz = (int(np.float32(t) * np.float32(32)) for t in tc)
    o3 = occ.reshape(32, 32, 32)
    bits = (int(o3[cz, cy, :].sum()) & 31) | (int(o3[cz, :, cx].sum()) & 31) << 5 | (int(o3[:, cy, cx].sum()) & 31) << 10 | int(o3[cz, cy, cx]) << 15
    assert wm[z, y, x, 3] == bits
    # Texels 3..5 and 26..28 ARE written, but with 0: their tc falls in warp cells 0..2 / 29..31 whose WEIGHTS
    # texels were never written by the 0.8 quad (generateWarpmap.frag:62-63 reads them NEAREST) -> l = h = 0.
    assert not wm[5, 9, 3:6, 0].any() and not wm[5, 9, 26:29, 0].any()
    # inside, the map is monotone along x within a row (piecewise-linear, slopes l/h > 0)
    row = wm[5, 9, 6:26, 0].astype(np.int64)
    assert np.all(np.diff(row) > 0)
    # without the weights texture (useWarpmapWeightsTexture=false) the table is read directly: no zero fringe
    p.use_warpmap_weights_texture = 0
    wm2 = np.zeros(32 ** 3 * 4, np.uint16)
    L.orc_warpmap(ol.ptr(occ), C.byref(p), ol.ptr(wm2), ol.ptr(lo), ol.ptr(hi))
    row = wm2.reshape(32, 32, 32, 4)[5, 9, 3:29, 0].astype(np.int64)
    assert np.all(np.diff(row) > 0) and row[0] > 0


# ------------------------------------------------------------------------------- filterRadiance.comp:25-35
def test_mip_box2_known_brick(L):
    src = np.zeros((4, 4, 4), np.uint32)
    src[0, 0, 0] = 0xFF0000FF                  # one opaque red child in the (0,0,0) brick
    src[2:4, 2:4, 2:4] = 0x80FFFFFF            # a full brick, alpha 128
    dst = np.zeros(8, np.uint32)
    L.orc_mip(4, ol.ptr(src.reshape(-1)), ol.ptr(dst), 0)
    dst = dst.reshape(2, 2, 2)
    assert dst[0, 0, 0] == 0x20000020          # 255/8 = 31.875 -> 32 on R and A
    assert dst[1, 1, 1] == 0x80FFFFFF
    assert dst[0, 1, 1] == 0 and dst[1, 0, 0] == 0


def test_mip_modes_box3_cube(L):
    src = np.full(4 ** 3, 0xFFFFFFFF, np.uint32)
    for mode, k, taps in ((1, 0.037, 27), (2, 0.143, 7)):
        dst = np.zeros(8, np.uint32)
        L.orc_mip(4, ol.ptr(src), ol.ptr(dst), mode)
        # interior texel (1,1,1) reads `taps` in-bounds texels of value 1.0
        v = min(1.0, np.float32(taps) * np.float32(k))
        assert (int(dst[7]) & 255) == int(np.rint(np.float32(v) * np.float32(255)))


# ------------------------------------------------------------------------------------- traceCone closed form
def test_cone_march_constant_fog(L):
    """Through a volume whose every level holds alpha a, alpha_n = 1-(1-a)^n until it crosses 0.95; the step
    schedule is h *= 1 + tan(theta/2) (phong.frag:145-177)."""
    cs = P.ConeSettings(16, math.radians(60.0), 1.0, 1.0, 0.5)
    steps = C.c_int(0)
    a_byte = 51                                  # 0.2
    alpha = L.orc_cone_trace_const(256, 6, a_byte << 24, C.byref(cs), C.byref(steps))
    a = a_byte / 255.0
    # expected step count: marching from the centre along +z leaves the unit cube when (bias + h)/D > 0.5
    h, n, acc = 1.0, 0, 0.0
    while n < 16 and acc < 0.95:
        if 0.5 + (1.0 + h) / 256.0 > 1.0:
            break
        acc += (1 - acc) * a; h += h * math.tan(math.radians(30.0)); n += 1
    assert steps.value == n == GOLD["cone_fog"]["steps_256"]
    assert alpha == pytest.approx(1 - (1 - a) ** n, abs=1e-5)
    # an opaque volume saturates in one step
    alpha = L.orc_cone_trace_const(256, 6, 0xFF000000, C.byref(cs), C.byref(steps))
    assert steps.value == 1 and alpha == pytest.approx(1.0)


def test_diffuse_march_schedule():
    """h = 1, 1.577, 2.488 ... (x1.57735/step); lambda = log2(max(1,2r))+0.5 = 0.71, 1.37, 2.02 ..."""
    h, t = np.float32(1.0), np.float32(math.tan(np.float32(math.radians(60.0)) / np.float32(2)))
    hs, lam = [], []
    for _ in range(8):
        r = h * t
        hs.append(float(h)); lam.append(float(np.log2(max(np.float32(1), np.float32(2) * r)) + np.float32(0.5)))
        h = h + r
    assert np.allclose(hs, GOLD["diffuse_schedule"]["h"], rtol=2e-3)
    assert np.allclose(lam, GOLD["diffuse_schedule"]["lambda"], atol=6e-3)
    assert lam[7] > 5.0                          # clamped to the coarsest level (L-1 = 5) from step 7 on


# ------------------------------------------------------------------ multisample coverage (Settings::conservativeRasterization == MSAA)
def _one_triangle_scene(cx, cy, r, z=0.3):
    """A small +z-facing triangle around (cx, cy) in the z-dominant view of a +-1 volume (window = (ndc * 0.5 + 0.5) * D, y up)."""
    from vct_b200 import scene as S
    sc = S.Scene()
    mat = sc.add_material(diffuse=sc.add_texture(np.full((1, 1, 3), 200, np.uint8)))
    verts = [[cx - r, cy - r, z, 0, 0, 1, 0, 0] + [0] * 6, [cx + r, cy - r, z, 0, 0, 1, 1, 0] + [0] * 6, [cx, cy + r, z, 0, 0, 1, 0, 1] + [0] * 6]
    sc.add_actor(S.Mesh(np.array(verts, np.float32), [0, 1, 2], np.array([mat], np.int32)))
    sc.lights = [P.make_light(position=(0.0, 0.0, 3.0), direction=(0.0, 0.0, -1.0), shadow_caster=False, type_=1)]
    return sc


def test_msaa_fragment_where_only_a_sample_is_covered():
    """OpenGL 4.5 section 14.6.6 with the standard 4x pattern: a triangle 0.3 pixels across around sample 0 of pixel (5, 7) — (5.375, 7.125) —
    covers neither that pixel's centre nor any other sample: no fragment with centre sampling, exactly one with multisampling, and its voxel
    is the one of the pixel CENTRE (inputs are interpolated there, extrapolated beyond the triangle)."""
    D = 16
    to_world = lambda w: w / D * 2.0 - 1.0
    sc = _one_triangle_scene(to_world(5.375), to_world(7.125), 0.02)
    cam = P.Camera(position=(0.0, 0.0, 2.5), front=(0.0, 0.0, -1.0))
    p = P.default_params(64, 64, cam, sc.lights[0], voxel_min=-1.0, voxel_max=1.0)
    p.voxelize_lighting = 0
    o = ol.Oracle(sc, D, 3, 64, 64, 64)
    o.voxelize(p)
    assert o.info.total_fragments == 0 and not o.color[0].any()
    p.conservative_raster = P.RASTER_MSAA
    o.voxelize(p)
    assert o.info.total_fragments == 1
    (idx,) = np.nonzero(o.color[0])
    assert len(idx) == 1
    x, y, zc = int(idx[0]) % D, (int(idx[0]) // D) % D, int(idx[0]) // (D * D)
    assert (x, y, zc) == (5, 7, 10)                            # z = 0.3 in a +-1 volume: (0.3 + 1) / 2 * 16 = 10.4
    assert o.color[0][idx[0]] >> 24 == 1                       # one insertion into the running average
    # every sample at the pixel centre: the centre-sampling result, whatever the scene
    p.msaa_samples = (C.c_float * 8)(*([0.5] * 8))
    o.voxelize(p)
    assert o.info.total_fragments == 0
    # a sample pattern is a parameter: move sample 0 away and the triangle is missed again
    p.msaa_samples = (C.c_float * 8)(0.125, 0.875, 0.875, 0.375, 0.125, 0.625, 0.625, 0.875)
    o.voxelize(p)
    assert o.info.total_fragments == 0


def test_msaa_is_a_superset_of_its_samples_and_centre_pattern_equals_off():
    from vct_b200 import scene as S
    sc = S.room_scene()
    D, Lv, SS, W, H = 32, 4, 128, 64, 48
    p = S.room_params(W, H)
    o = ol.Oracle(sc, D, Lv, SS, W, H)
    o.shadowmap(p); o.voxelize(p)
    off_color, off_frags = o.color[0].copy(), o.info.total_fragments
    q = type(p).from_buffer_copy(p); q.conservative_raster = P.RASTER_MSAA
    o.voxelize(q)
    assert o.info.total_fragments > off_frags * 1.05            # thin and grazing triangles gain fragments
    assert ((off_color != 0) & (o.color[0] == 0)).sum() <= off_frags * 0.02   # nearly every centre-covered voxel is still hit
    q.msaa_samples = (C.c_float * 8)(*([0.5] * 8))
    o.voxelize(q)
    assert o.info.total_fragments == off_frags and np.array_equal(o.color[0], off_color)
    q.conservative_raster = P.RASTER_MSAA; q.warp_texture = 1; q.msaa_samples = (C.c_float * 8)(*([0.0] * 8))
    o.occupancy(q)
    occ_ms = o.occ.copy()
    q.conservative_raster = P.RASTER_CENTER
    o.occupancy(q)
    assert (occ_ms != 0).sum() > (o.occ != 0).sum()             # the occupancy pass is multisampled too (Application.cpp:244-249)


def test_voxelize_multiplier_scales_the_viewport_not_the_grid():
    """Settings::voxelizeMultiplier (Application.cpp:668): m * dim pixels per side, voxel index from the interpolated position: m = 2 gives about
    four times the fragments in (nearly) the same voxels; 0 reads as 1."""
    from vct_b200 import scene as S
    sc = S.room_scene()
    D, Lv, SS, W, H = 32, 4, 128, 64, 48
    p = S.room_params(W, H); p.voxelize_atomic_max = 1
    o = ol.Oracle(sc, D, Lv, SS, W, H)
    o.shadowmap(p); o.voxelize(p)
    base, n1 = o.color[0].copy(), o.info.total_fragments
    p.voxelize_multiplier = 1.0
    o.voxelize(p)
    assert o.info.total_fragments == n1 and np.array_equal(o.color[0], base)
    p.voxelize_multiplier = 2.0
    o.voxelize(p)
    assert 3.6 * n1 < o.info.total_fragments < 4.4 * n1
    assert ((base != 0) & (o.color[0] == 0)).sum() == 0          # on this scene every voxel of the coarse pass is hit again (no theorem: slivers can slip between samples)
    assert (o.color[0] != 0).sum() >= (base != 0).sum()
    p.voxelize_multiplier = 0.5
    o.voxelize(p)
    assert 0.2 * n1 < o.info.total_fragments < 0.3 * n1


# ------------------------------------------------------------------ Application::debugVoxels: one voxel seen head-on
def test_debug_voxels_single_cube_head_on():
    """One occupied voxel of an 8^3 grid over +-1, the camera on its axis looking down -z: only the cube's near face is visible (the far one
    is back-facing and culled, the sides are edge-on), so the image is one axis-aligned rectangle whose pixel set follows from the projection
    matrix by hand — the face spans [0, 0.25]^2 at z = 0.25, 2.875 in front of the eye."""
    D, W, H = 8, 200, 120
    cam = P.Camera(position=(0.125, 0.125, 3.125), front=(0.0, 0.0, -1.0))
    light = P.make_light(position=(0.0, 3.0, 0.0), direction=(-0.2, -1.0, -0.1), shadow_caster=False, type_=1)
    p = P.default_params(W, H, cam, light, voxel_min=-1.0, voxel_max=1.0)
    vol = np.zeros(D ** 3, np.uint32)
    vol[(4 * D + 4) * D + 4] = 0xFF336699                          # voxel (4, 4, 4): texcoords [0.5, 0.625)^3 -> world [0, 0.25)^3
    image = np.zeros(W * H, np.uint32)
    ol.lib().orc_debug_voxels(C.byref(p), W, H, D, 1, ol.ptr(vol), ol.ptr(image))
    img = image.reshape(H, W)
    clear = img[0, 0]
    ys, xs = np.nonzero(img != clear)
    assert len(xs) > 0 and np.unique(img[img != clear]).tolist() == [0xFF336699]
    proj = np.array(p.projection[:], np.float64).reshape(4, 4).T   # column-major float[16]
    fx, fy = proj[0, 0], proj[1, 1]
    half = 0.125 / 2.875                                           # half extent of the near face over its distance
    x_lo, x_hi = (0.5 - 0.5 * fx * half) * W, (0.5 + 0.5 * fx * half) * W      # window-space edges of the face (the eye looks at its centre)
    y_lo, y_hi = (0.5 - 0.5 * fy * half) * H, (0.5 + 0.5 * fy * half) * H
    want_x = [i for i in range(W) if x_lo < i + 0.5 < x_hi]
    want_y = [j for j in range(H) if y_lo < j + 0.5 < y_hi]
    assert sorted(set(xs.tolist())) == want_x and sorted(set(ys.tolist())) == want_y
    assert len(xs) == len(want_x) * len(want_y)                    # a filled rectangle, nothing else
    # an empty volume draws nothing; a fully transparent colour (alpha 0) draws nothing either (debugVoxels.geom:46)
    vol[:] = 0; vol[0] = 0x00FFFFFF
    ol.lib().orc_debug_voxels(C.byref(p), W, H, D, 1, ol.ptr(vol), ol.ptr(image))
    assert (image == clear).all()
