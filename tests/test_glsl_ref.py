"""The CPU oracle against the REFERENCE'S OWN GLSL, compiled as C++.

oracle/Makefile reads each shader from /root/reference/shaders where it lies, maps the GLSL syntax that is not C++ with
oracle/ref_rig/glsl2cpp.py (the shader bodies stay the reference's: every expression, branch and constant), and compiles it
against oracle/ref_rig/glsl_shim.h into oracle/_ref/libvct_glsl_ref.so.  These tests run the oracle's restatement
(oracle/vct_oracle.cpp) and the compiled shader on the same seeded inputs and require identical words / texels.
Where /root/reference is absent (the GPU box) the live comparison is skipped and the committed outputs of the compiled
shaders (tests/golden/glsl_ref.npz, generator tests/golden/make_glsl_ref.py, same seeds) pin the oracle instead.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import oracle_lib as O
from vct_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libvct_glsl_ref.so")
GOLD = os.path.join(ROOT, "tests", "golden", "glsl_ref.npz")
live = pytest.mark.skipif(not os.path.isfile(SO), reason="oracle/_ref/libvct_glsl_ref.so not built (needs /root/reference)")
_g = None


def glsl():
    global _g
    if _g is None:
        _g = C.CDLL(SO)
        _g.glsl_set_voxel_opacity.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _g.glsl_normalize_voxels_f16.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _g.glsl_temporal_radiance_filter.argtypes = [C.c_int, C.c_float, C.c_void_p]
    return _g


ptr = O.ptr


# ------------------------------------------------------------------------------------------------ seeded inputs
def voxel_volume(D, seed, fill=0.12, counts=True):
    """RGBA8 words: `fill` of the voxels occupied; alpha = fragment count (1..40, a few at 255) like the voxeliser leaves it."""
    rng = np.random.default_rng(seed)
    occ = rng.random(D ** 3) < fill
    rgb = rng.integers(0, 256, (D ** 3, 3), dtype=np.uint32)
    a = rng.integers(1, 41, D ** 3, dtype=np.uint32) if counts else rng.integers(1, 256, D ** 3, dtype=np.uint32)
    a[rng.random(D ** 3) < 0.01] = 255
    w = rgb[:, 0] | rgb[:, 1] << 8 | rgb[:, 2] << 16 | a << 24
    return np.where(occ, w, 0).astype(np.uint32)


def frame_params(D, **kw):
    cam = P.Camera(position=(1.5, 1.0, 2.0), front=(-1.5, -0.8, -2.0))
    light = P.reference_lights()[0]
    light.position = P.F3(3.0, 9.0, 2.0); light.direction = P.F3(-0.3, -0.9, -0.25)
    p = P.default_params(64, 48, cam, light, voxel_min=-4.0, voxel_max=4.0, voxel_center=(0.25, 1.0, -0.5))
    for k, v in kw.items():
        setattr(p, k, v)
    return p, light


def shadow_map(S, seed):
    """Depths of a bumpy surface seen from the light + holes at the far plane (1.0)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:S, 0:S].astype(np.float32) / S
    d = 0.08 + 0.03 * np.sin(9 * x) * np.cos(7 * y) + 0.02 * rng.random((S, S), dtype=np.float32)
    d[rng.random((S, S)) < 0.05] = 1.0
    return np.ascontiguousarray(d.astype(np.float32).reshape(-1))


def warp_map(seed):
    """A monotone, non-trivial 32^3 RGBA16 warp map (piecewise-linear per axis, like generateWarpmap produces)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((32, 32, 32, 4), np.uint16)
    for axis in range(3):
        w = rng.choice([0.5, 1.0, 2.0], 32)
        edges = np.concatenate([[0.0], np.cumsum(w) / w.sum()])
        centres = (edges[:-1] + edges[1:]) / 2
        shape = [1, 1, 1]; shape[2 - axis] = 32
        out[..., axis] = np.round(centres.reshape(shape) * 65535).astype(np.uint16)
    out[..., 3] = 65535
    return np.ascontiguousarray(out.reshape(-1))


# ------------------------------------------------------------------------------------------------------- cases
def run_cases(impl):
    """impl: 'oracle' or 'glsl' -> dict name -> array.  Same inputs, same call order for both."""
    lib = O.lib() if impl == "oracle" else glsl()
    f = (lambda n: getattr(lib, "orc_" + n)) if impl == "oracle" else (lambda n: getattr(lib, "glsl_" + n))
    out = {}
    # a3 transferVoxels.comp — plain, temporal (radiance history kept), opacity 0 (alpha kept)
    for name, kw in (("plain", {}), ("temporal", {"temporal_filter_radiance": 1, "temporal_decay": 0.8}),
                     ("opacity0", {"voxel_set_opacity": 0.0}), ("temporal_decay_third", {"temporal_filter_radiance": 1, "temporal_decay": 1.0 / 3.0})):
        for D in (16, 20):                                              # 20: dispatch overhang past the image
            p, _ = frame_params(D, **kw)
            col, rad = voxel_volume(D, 1), voxel_volume(D, 2, fill=0.3, counts=False)
            info = P.VoxelizeInfo(7, 0, 0)
            f("transfer")(C.byref(p), D, ptr(col), ptr(rad), C.byref(info))
            out[f"transfer_{name}_{D}_color"] = col; out[f"transfer_{name}_{D}_radiance"] = rad
            out[f"transfer_{name}_{D}_info"] = np.array([info.total_fragments, info.unique_voxels, info.max_fragments_per_voxel], np.uint32)
    # a6 filterRadiance.comp — BOX2 (live), BOX3, CUBE; a whole chain 32 -> 1
    for mode in (0, 1, 2):
        src = voxel_volume(32, 3 + mode, fill=0.4, counts=False)
        Ds = 32
        while Ds > 1:
            dst = np.zeros((Ds // 2) ** 3, np.uint32)
            f("mip")(Ds, ptr(src), ptr(dst), mode)
            out[f"mip_mode{mode}_{Ds}"] = dst
            src, Ds = dst, Ds // 2
    # a4 voxelFillHoles.comp
    for D in (16, 12):
        rad = voxel_volume(D, 9, fill=0.2, counts=False)
        f("fill_holes")(D, ptr(rad))
        out[f"fill_holes_{D}"] = rad
    # a5 injectRadiance.comp — linear mapping, cubic camera warp, warp texture, radianceLighting.  (radianceLighting together
    # with the temporal filter is left out: injectRadiance.comp:87-94 reads back voxelRadiance that other invocations of the
    # same dispatch write without atomics — order-dependent in the reference, whose own comment says it "doesn't work".)
    S = 96
    for name, kw in (("linear", {}), ("warp_voxels", {"warp_voxels": 1}), ("warp_texture", {"warp_texture": 1}),
                     ("lighting", {"radiance_lighting": 1}),
                     # voxelizeTesselationWarp (common.glsl:37-42, 52-54): the camera frustum as voxel grid; below warpTexture in priority
                     ("tess_warp", {"voxelize_tesselation_warp": 1}), ("tess_warp_below_warp_texture", {"voxelize_tesselation_warp": 1, "warp_texture": 1})):
        D = 24
        p, light = frame_params(D, **kw)
        col, nrm = voxel_volume(D, 11, fill=0.5), voxel_volume(D, 12, fill=1.0, counts=False)
        rad = voxel_volume(D, 13, fill=0.2, counts=False)
        sm, wm = shadow_map(S, 14), warp_map(15)
        lp = (C.c_float * 3)(*light.position); li = (C.c_float * 3)(0.9, 0.8, 0.7)
        f("inject")(C.byref(p), D, ptr(col), ptr(nrm), ptr(sm), S, ptr(wm) if p.warp_texture else None, lp, li, ptr(rad))
        out[f"inject_{name}"] = rad
    # dead variants: setVoxelOpacity.comp, temporalRadianceFilter.comp, normalizeVoxels.comp (RGBA16F)
    for D in (8, 10):
        col, rad = voxel_volume(D, 21), np.zeros(D ** 3, np.uint32)
        info = P.VoxelizeInfo(0, 0, 0)
        f("set_voxel_opacity")(D, 0.5, ptr(col), ptr(rad), C.byref(info))
        out[f"set_opacity_{D}_color"] = col; out[f"set_opacity_{D}_radiance"] = rad
        out[f"set_opacity_{D}_info"] = np.array([info.unique_voxels, info.max_fragments_per_voxel], np.uint32)
        vol = voxel_volume(D, 22, fill=0.6, counts=False)
        f("temporal_radiance_filter")(D, 0.8, ptr(vol))
        out[f"temporal_filter_{D}"] = vol
        src = voxel_volume(2 * D, 24, fill=0.5, counts=False); dst = np.zeros(D ** 3, np.uint32)      # filter3d.comp: 2x2x2 mean through the sampler
        f("filter3d")(2 * D, ptr(src), ptr(dst))
        out[f"filter3d_{2 * D}"] = dst
        rng = np.random.default_rng(23)
        cnt = np.where(rng.random(D ** 3) < 0.3, rng.integers(1, 30, D ** 3), 0).astype(np.float32)
        c16 = (rng.random((D ** 3, 4), dtype=np.float32) * cnt[:, None]).astype(np.float16); c16[:, 3] = cnt.astype(np.float16)
        n16 = (rng.random((D ** 3, 4), dtype=np.float32) * cnt[:, None]).astype(np.float16); n16[:, 3] = cnt.astype(np.float16)
        c16, n16 = np.ascontiguousarray(c16.view(np.uint16).reshape(-1)), np.ascontiguousarray(n16.view(np.uint16).reshape(-1))
        rad = np.zeros(D ** 3, np.uint32); info = P.VoxelizeInfo(0, 0, 0)
        f("normalize_voxels_f16")(D, 0.5, ptr(c16), ptr(n16), ptr(rad), C.byref(info))
        out[f"normalize_{D}_color"] = c16; out[f"normalize_{D}_normal"] = n16; out[f"normalize_{D}_radiance"] = rad
        out[f"normalize_{D}_info"] = np.array([info.unique_voxels, info.max_fragments_per_voxel], np.uint32)
    return out


def compare(a, b, what):
    assert sorted(a) == sorted(b)
    bad = [k for k in sorted(a) if not np.array_equal(a[k], b[k])]
    detail = ""
    if bad:
        k = bad[0]; d = np.flatnonzero(a[k] != b[k])
        detail = f"; first: {k} differs in {d.size} of {a[k].size} words, e.g. [{d[0]}] {a[k][d[0]]:#x} vs {b[k][d[0]]:#x}"
    assert not bad, f"{what}: {len(bad)} of {len(a)} outputs differ: {bad[:6]}{detail}"


@live
def test_oracle_equals_the_reference_glsl_compiled_as_cpp():
    compare(run_cases("oracle"), run_cases("glsl"), "oracle vs compiled reference GLSL")


def test_oracle_equals_the_committed_outputs_of_the_reference_glsl():
    gold = dict(np.load(GOLD))
    compare(run_cases("oracle"), gold, "oracle vs tests/golden/glsl_ref.npz")


def test_cases_are_not_vacuous():
    o = run_cases("oracle")
    assert (o["transfer_plain_16_radiance"] >> 24 != 0).sum() > 100 and o["transfer_plain_16_info"][1] > 100
    assert (o["inject_linear"] != voxel_volume(24, 13, fill=0.2, counts=False)).sum() > 50       # texels landed and wrote
    assert len({o[k].tobytes() for k in o if k.startswith("inject_")}) == 5                        # every mode differs ...
    assert np.array_equal(o["inject_tess_warp_below_warp_texture"], o["inject_warp_texture"])      # ... except where the priority order says so
    assert (o["inject_tess_warp"] != voxel_volume(24, 13, fill=0.2, counts=False)).sum() > 50
    assert (o["fill_holes_16"] != voxel_volume(16, 9, fill=0.2, counts=False)).sum() > 500
    assert all((o[f"mip_mode{m}_2"] != 0).any() for m in (0, 1, 2))


# ============================================================ fragment shaders: voxelize.frag (a2) and phong.frag (a7)
# The rasteriser (fixed function) is the oracle's: orc_voxelize_trace / orc_shade_trace record what it hands the fragment
# stage; the reference's fragment shaders, compiled as C++, are then run on exactly those inputs, fragment by fragment in
# the same order, and must leave the same voxel words / the same RGBA8 pixels.
def pbr_room():
    """tests' room scene with normal, roughness and metallic maps on two materials (every phong.frag branch is reachable)."""
    from vct_b200 import scene as S
    sc = S.room_scene()
    rng = np.random.default_rng(77)
    nm = np.clip(np.array([128, 128, 235]) + rng.integers(-60, 61, (32, 32, 3)), 0, 255).astype(np.uint8)
    t_n = sc.add_texture(nm)
    t_r = sc.add_texture(rng.integers(20, 256, (16, 16), dtype=np.uint8).astype(np.uint8))
    t_m = sc.add_texture(rng.integers(0, 256, (8, 8), dtype=np.uint8).astype(np.uint8))
    sc.materials[0].normal_tex = t_n; sc.materials[0].roughness_tex = t_r; sc.materials[0].metallic_tex = t_m
    sc.materials[1].roughness_tex = t_r
    return sc


FRAG_MODES = {
    "default": {},
    "atomic_max": {"voxelize_atomic_max": 1},
    # Settings::conservativeRasterization == MSAA: fragments wherever any of the 4 samples is covered, inputs extrapolated to the pixel centre
    "msaa": {"conservative_raster": 1},
    # Settings::voxelizeMultiplier: a 1.5 x dim viewport (texture LOD and fragment count follow the viewport, the voxel index does not)
    "multiplier_1p5": {"voxelize_multiplier": 1.5},
    "no_voxel_lighting": {"voxelize_lighting": 0},
    "warp_voxels": {"warp_voxels": 1},
    "warp_texture": {"warp_texture": 1},
    "blinn_no_post": {"cooktorrance": 0, "enable_postprocess": 0},
    "direct_only": {"enable_indirect": 0},
    "no_reflections_no_normal_map": {"enable_reflections": 0, "enable_normal_map": 0, "draw_occlusion": 0},
    "color_volume": {"draw_radiance": 0},
    "fixed_specular_angle": {"specular_cone_angle_from_roughness": 0},
    # both warps on: the voxeliser and the voxel view take warpVoxels first (voxelize.frag:100-105, common.glsl:44-50), traceCone takes
    # warpTexture first (phong.frag:150-157)
    "warp_texture_and_voxels": {"warp_texture": 1, "warp_voxels": 1},
    "view_voxels_warp_texture_and_voxels": {"debug_view": P.VIEW_VOXELS, "miplevel": 0.7, "warp_texture": 1, "warp_voxels": 1},
    "tess_warp": {"voxelize_tesselation_warp": 1},                        # phong.frag:158-162: every cone sample through pv
    "tess_warp_below_warp_voxels": {"voxelize_tesselation_warp": 1, "warp_voxels": 1},
    # debug views of phong.frag
    "view_voxels_point": {"debug_view": P.VIEW_VOXELS, "miplevel": 0.0},
    "view_voxels_lod1_7": {"debug_view": P.VIEW_VOXELS, "miplevel": 1.7},
    "view_voxels_color_warp": {"debug_view": P.VIEW_VOXELS, "miplevel": 0.8, "draw_radiance": 0, "warp_voxels": 1},
    "view_material_diffuse": {"debug_view": P.VIEW_MATERIAL_DIFFUSE},
    "view_material_roughness": {"debug_view": P.VIEW_MATERIAL_ROUGHNESS},
    "view_material_metallic": {"debug_view": P.VIEW_MATERIAL_METALLIC},
    "view_normals": {"debug_view": P.VIEW_NORMALS},
    "view_dominant_axis": {"debug_view": P.VIEW_DOMINANT_AXIS},
    "view_indirect": {"debug_view": P.VIEW_INDIRECT},
    "view_indirect_no_occlusion": {"debug_view": P.VIEW_INDIRECT, "draw_occlusion": 0},
    "view_occlusion": {"debug_view": P.VIEW_OCCLUSION},
    "view_reflections": {"debug_view": P.VIEW_REFLECTIONS},
    "view_reflections_off": {"debug_view": P.VIEW_REFLECTIONS, "enable_reflections": 0},
    "view_voxels_tess_warp": {"debug_view": P.VIEW_VOXELS, "miplevel": 0.6, "voxelize_tesselation_warp": 1},
    # sub-views of `voxelize` (phong.frag:350-357): voxelNormal at the fragment's voxel; the warp map at the linear position; `toggle`
    "view_voxel_normals": {"debug_view": P.VIEW_VOXEL_NORMALS, "miplevel": 0.0},
    "view_voxel_normals_warp_voxels_lod": {"debug_view": P.VIEW_VOXEL_NORMALS, "miplevel": 2.3, "warp_voxels": 1},
    "view_warp_texture": {"debug_view": P.VIEW_WARP_TEXTURE, "warp_texture": 1},
    "view_warp_texture_tc": {"debug_view": P.VIEW_WARP_TEXTURE_TC, "warp_texture": 1},
}


def run_fragment_cases(impl, modes=None):
    from vct_b200 import scene as S
    D, Lv, SS, W, H = 32, 5, 256, 96, 64
    sc = pbr_room()
    out = {}
    for name, kw in FRAG_MODES.items():
        if modes and name not in modes:
            continue
        p = S.room_params(W, H)
        for k, v in kw.items():
            setattr(p, k, v)
        o = O.Oracle(sc, D, Lv, SS, W, H)
        o.shadowmap(p)
        if p.warp_texture:
            o.occupancy(p); o.warpmap_pass(p)
        wm = ptr(o.warpmap) if p.warp_texture else None
        cap = 1 << 18
        rec = np.zeros(cap * 16, np.float32); n = C.c_longlong(0)
        O.lib().orc_voxelize_trace(C.byref(o.s.c), C.byref(p), D, ptr(o.shadow), SS, wm, ptr(o.color[0]), ptr(o.normal), C.byref(o.info),
                                   ptr(rec), C.c_longlong(cap), C.byref(n))
        assert 1000 < n.value <= cap
        if impl == "glsl":
            col, nrm, info = np.zeros(D ** 3, np.uint32), np.zeros(D ** 3, np.uint32), P.VoxelizeInfo()
            glsl().glsl_voxelize_fragments(C.byref(o.s.c), C.byref(p), D, ptr(o.shadow), SS, wm, ptr(rec), C.c_longlong(n.value), ptr(col), ptr(nrm), C.byref(info))
            out[f"vox_{name}_color"], out[f"vox_{name}_normal"], out[f"vox_{name}_fragments"] = col, nrm, np.array([info.total_fragments], np.uint32)
        else:
            out[f"vox_{name}_color"], out[f"vox_{name}_normal"] = o.color[0].copy(), o.normal.copy()
            out[f"vox_{name}_fragments"] = np.array([o.info.total_fragments], np.uint32)
        o.transfer(p); o.inject(p); o.mip("radiance"); o.mip("color"); o.visibility(p)
        rad, colp = np.concatenate(o.radiance), np.concatenate(o.color)
        prec = np.zeros(W * H * 28, np.float32); steps = C.c_ulonglong(0)
        O.lib().orc_set_normal_volume(ptr(o.normal))
        O.lib().orc_shade_trace(C.byref(o.s.c), C.byref(p), W, H, ptr(o.vis), D, Lv, ptr(rad), ptr(colp), ptr(o.shadow), SS, wm, ptr(o.image), C.byref(steps), ptr(prec))
        if impl == "glsl":
            img = np.zeros(W * H, np.uint32); gsteps = C.c_ulonglong(0)
            glsl().glsl_set_normal_volume(ptr(o.normal))
            glsl().glsl_shade_pixels(C.byref(o.s.c), C.byref(p), W, H, ptr(prec), D, Lv, ptr(rad), ptr(colp), ptr(o.shadow), SS, wm, ptr(img), C.byref(gsteps))
            out[f"shade_{name}_image"], out[f"shade_{name}_cone_steps"] = img, np.array([gsteps.value], np.uint64)
        else:
            out[f"shade_{name}_image"], out[f"shade_{name}_cone_steps"] = o.image.copy(), np.array([steps.value], np.uint64)
    return out


GOLD_FRAG = os.path.join(ROOT, "tests", "golden", "glsl_ref_fragment.npz")


@live
def test_oracle_fragment_stages_equal_the_reference_glsl_compiled_as_cpp():
    compare(run_fragment_cases("oracle"), run_fragment_cases("glsl"), "oracle vs compiled voxelize.frag / phong.frag")


def test_oracle_fragment_stages_equal_the_committed_outputs_of_the_reference_glsl():
    gold = dict(np.load(GOLD_FRAG))
    compare(run_fragment_cases("oracle"), gold, "oracle vs tests/golden/glsl_ref_fragment.npz")


# ===================================================== a9: generateWarpmapWeights.frag + generateWarpmap.frag (thesis pass)
def run_warpmap_cases(impl):
    """Occupancy grids -> (CPU prefix sums + weight table, already pinned to Application.cpp:311-370) -> the two layered
    fragment passes.  The oracle does everything inside orc_warpmap; the GLSL side gets the tables from the oracle's
    orc_warp_partials / orc_warp_weight_table and runs the reference's two fragment shaders."""
    out = {}
    rng = np.random.default_rng(99)
    grids = {"random10": (rng.random(32 ** 3) < 0.10), "random60": (rng.random(32 ** 3) < 0.60),
             "empty": np.zeros(32 ** 3, bool), "full": np.ones(32 ** 3, bool)}
    slab = np.zeros((32, 32, 32), bool); slab[4:9, :, 10:30] = True; slab[20, 3:7, :] = True
    grids["slabs"] = slab.reshape(-1)
    variants = {"default": {}, "no_weights_texture": {"use_warpmap_weights_texture": 0}, "linear": {"warp_texture_linear": 1},
                "axes_x_only": {"warp_texture_axes": (1, 0, 0)}, "high3_low025": {"warp_texture_high_resolution": 3.0, "warp_texture_low_resolution": 0.25}}
    for gname, g in grids.items():
        occ = np.ascontiguousarray(g.astype(np.uint32))
        for vname, kw in variants.items():
            if gname not in ("random10", "slabs") and vname != "default":
                continue
            p, _ = frame_params(32, warp_texture=1)
            for k, v in kw.items():
                if k == "warp_texture_axes":
                    for i in range(3):
                        p.warp_texture_axes[i] = v[i]
                else:
                    setattr(p, k, v)
            wm, lo, hi = np.zeros(32 ** 3 * 4, np.uint16), np.zeros(32 ** 3 * 4, np.uint16), np.zeros(32 ** 3 * 4, np.uint16)
            if impl == "oracle":
                O.lib().orc_warpmap(ptr(occ), C.byref(p), ptr(wm), ptr(lo), ptr(hi))
            else:
                parts = np.zeros(32 ** 3 * 3, np.int32)
                O.lib().orc_warp_partials(ptr(occ), ptr(parts))
                wl, wh = np.zeros(33, np.float32), np.zeros(33, np.float32)
                O.lib().orc_warp_weight_table(32, C.c_float(p.warp_texture_high_resolution), C.c_float(p.warp_texture_low_resolution), ptr(wl), ptr(wh))
                table = np.ascontiguousarray(np.concatenate([wl, wh]))
                g_ = glsl()
                g_.glsl_warp_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
                if p.use_warpmap_weights_texture:
                    g_.glsl_warp_weights(ptr(occ), ptr(parts), ptr(table), p.warp_texture_high_resolution, ptr(lo), ptr(hi))
                g_.glsl_warpmap(ptr(occ), ptr(parts), ptr(table), C.byref(p), ptr(lo), ptr(hi), ptr(wm))
            out[f"warp_{gname}_{vname}_map"], out[f"warp_{gname}_{vname}_low"], out[f"warp_{gname}_{vname}_high"] = wm, lo, hi
    return out


GOLD_WARP = os.path.join(ROOT, "tests", "golden", "glsl_ref_warpmap.npz")


@live
def test_oracle_warpmap_equals_the_reference_glsl_compiled_as_cpp():
    compare(run_warpmap_cases("oracle"), run_warpmap_cases("glsl"), "oracle vs compiled generateWarpmap{,Weights}.frag")


def test_oracle_warpmap_equals_the_committed_outputs_of_the_reference_glsl():
    compare(run_warpmap_cases("oracle"), dict(np.load(GOLD_WARP)), "oracle vs tests/golden/glsl_ref_warpmap.npz")


# ========================================== N4: tessellation voxeliser — testTesselation.tesc / .tese (the reference's default)
def _tess_patch(w9, p, D, cap=8192):
    lev = np.zeros(4, np.float32); uvw = np.zeros(cap * 3, np.float32)
    O.lib().orc_tess_patch.restype = C.c_longlong
    n = O.lib().orc_tess_patch(ptr(np.ascontiguousarray(w9, np.float32)), C.byref(p), D, ptr(lev), ptr(uvw), C.c_longlong(cap))
    return lev, uvw[:3 * min(n, cap)].reshape(-1, 3), n


def random_triangles(n, seed, scale):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-3.0, 3.0, (n, 1, 3)); e = rng.normal(0, 1, (n, 3, 3)) * rng.choice(scale, (n, 1, 1))
    t = (c + e).astype(np.float32)
    t[0] = [[0, 0, 0], [0, 0, 0], [0, 0, 0]]; t[1] = [[1, 1, 1], [2, 2, 2], [3, 3, 3]]          # degenerate: point, collinear
    t[2] = [[0.1, 0.1, 0.1], [0.1001, 0.1, 0.1], [0.1, 0.1001, 0.1]]                           # inside one voxel: discarded
    return t.reshape(n, 9)


def test_canonical_tessellator_invariants():
    """OpenGL 4.5 §11.2.2.1 (triangles, equal_spacing, point_mode) as the oracle states it: point counts, no duplicates,
    barycentric sums, symmetry under relabelling, and the spec's special cases."""
    p, _ = frame_params(64)
    D = 64
    counts = {}
    for w9 in random_triangles(200, 5, [0.02, 0.1, 0.4, 1.5]):
        lev, uvw, n = _tess_patch(w9, p, D)
        if not (lev[1:] > 0).all():
            assert n == 0                                                   # discarded patch
            continue
        rnd = lambda x: 1 if not x > 1 else 64 if x >= 64 else int(np.ceil(x))
        ni, o = rnd(lev[0]), [rnd(x) for x in lev[1:]]
        if ni == 1 and o == [1, 1, 1]:
            want = 3
        else:
            ni = max(ni, 2)
            want = 3 + sum(x - 1 for x in o) + sum(3 * (ni - 2 * j) if ni - 2 * j > 0 else 1 for j in range(1, ni // 2 + 1))
        assert n == want == len(uvw), (lev, n, want)
        assert np.abs(uvw.sum(axis=1) - 1).max() < 3e-7 and uvw.min() >= 0 and uvw.max() <= 1
        assert len({tuple(r) for r in uvw.tolist()}) == n                    # distinct vertices, each emitted once
        counts[n] = counts.get(n, 0) + 1
        # inner rings are invariant under cyclic relabelling of the corners
        inner = uvw[3 + sum(x - 1 for x in o):]
        assert {tuple(r) for r in inner.tolist()} == {tuple(np.roll(r, 1)) for r in inner.tolist()}
    assert len(counts) > 20 and max(counts) > 1000
    lev, uvw, n = _tess_patch([0, 0, 0, 0.2, 0, 0, 0, 0.2, 0], p, D)       # ~1.6 voxels per edge -> outer 2.., inner small
    assert n >= 4 and (uvw[:3] == np.eye(3, dtype=np.float32)).all()


@live
def test_tess_control_shader_equals_the_reference_glsl():
    """testTesselation.tesc:32-78 (levels from edge lengths / altitudes in voxels; 0 for a patch inside one voxel), bit for bit,
    incl. degenerate triangles (NaN altitudes go through GLSL max(1.0, NaN) = 1.0)."""
    g = glsl(); g.glsl_tess_control.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    for D in (32, 256):
        p, _ = frame_params(D)
        for w9 in random_triangles(400, 7, [0.01, 0.05, 0.3, 2.0, 6.0]):
            lev, _, _ = _tess_patch(w9, p, D, cap=1)
            ref = np.zeros(4, np.float32)
            g.glsl_tess_control(ptr(np.ascontiguousarray(w9, np.float32)), C.byref(p), D, ptr(ref))
            assert np.array_equal(lev.view(np.uint32), ref.view(np.uint32)), (w9, lev, ref)


def run_tess_cases(impl):
    from vct_b200 import scene as S
    D, SS, W, H = 32, 64, 32, 32
    sc = pbr_room()
    out = {}
    for name, kw in (("max", {"voxelize_atomic_max": 1}), ("avg", {"voxelize_atomic_max": 0}),
                     ("max_tess_warp", {"voxelize_atomic_max": 1, "voxelize_tesselation_warp": 1}),
                     # voxelIndex(..., false) in a program whose warpTexture uniform the host never sets: the frame's other warps do not apply
                     ("max_ignores_other_warps", {"voxelize_atomic_max": 1, "warp_voxels": 1})):
        p = S.room_params(W, H)
        for k, v in kw.items():
            setattr(p, k, v)
        o = O.Oracle(sc, D, 5, SS, W, H)
        cap = 1 << 20
        rec = np.zeros(cap * 8, np.float32); n = C.c_longlong(0)
        O.lib().orc_voxelize_tess_trace(C.byref(o.s.c), C.byref(p), D, ptr(o.color[0]), ptr(o.normal), C.byref(o.info), ptr(rec), C.c_longlong(cap), C.byref(n))
        assert 5000 < n.value <= cap
        if impl == "glsl":
            nv = o.s.verts.shape[0]
            wpos, wnrm = np.zeros(nv * 3, np.float32), np.zeros(nv * 3, np.float32)
            O.lib().orc_world_vertices(C.byref(o.s.c), ptr(wpos), ptr(wnrm))
            col, nrm = np.zeros(D ** 3, np.uint32), np.zeros(D ** 3, np.uint32)
            glsl().glsl_tess_eval(C.byref(o.s.c), C.byref(p), D, ptr(wpos), ptr(wnrm), ptr(rec), C.c_longlong(n.value), ptr(col), ptr(nrm))
            out[f"tess_{name}_color"], out[f"tess_{name}_normal"] = col, nrm
        else:
            out[f"tess_{name}_color"], out[f"tess_{name}_normal"] = o.color[0].copy(), o.normal.copy()
        out[f"tess_{name}_points"] = np.array([n.value], np.uint64)
    return out


GOLD_TESS = os.path.join(ROOT, "tests", "golden", "glsl_ref_tess.npz")


@live
def test_oracle_tess_evaluation_equals_the_reference_glsl_compiled_as_cpp():
    a = run_tess_cases("oracle")
    assert (a["tess_max_color"] >> 24 != 0).sum() > 1500 and not np.array_equal(a["tess_max_color"], a["tess_avg_color"])
    compare(a, run_tess_cases("glsl"), "oracle vs compiled testTesselation.tese")


def test_oracle_tess_evaluation_equals_the_committed_outputs_of_the_reference_glsl():
    compare(run_tess_cases("oracle"), dict(np.load(GOLD_TESS)), "oracle vs tests/golden/glsl_ref_tess.npz")


def test_reference_cpu_frame_equals_the_oracle_frame():
    """bench.py's CPU arm (tests/oracle_lib.GlslReference: the compiled reference shaders around the oracle's fixed function)
    produces the oracle's frame: every pyramid level, counters, and the sampled image rows."""
    if not os.path.isfile(SO):
        pytest.skip("oracle/_ref/libvct_glsl_ref.so not built")
    import bench
    from vct_b200 import scene as S
    D, Lv, SS, W, H = 32, 5, 256, 96, 64
    sc = pbr_room()
    p = S.room_params(W, H)
    a, b = O.Oracle(sc, D, Lv, SS, W, H), O.Oracle(sc, D, Lv, SS, W, H)
    for o in (a, b):
        o.shadowmap(p); o.visibility(p)
    bench.oracle_gi_frame(a, p, 2)
    r = O.GlslReference(b)
    t = bench.reference_gi_frame(r, p, 2)
    assert set(t) == {"voxelize", "transfer", "inject", "mip", "cone_trace"} and all(v > 0 for v in t.values())
    for l in range(Lv):
        assert np.array_equal(a.radiance[l], b.radiance[l]) and np.array_equal(a.color[l], b.color[l]), l
    assert np.array_equal(a.normal, b.normal)
    assert (a.info.total_fragments, a.info.unique_voxels, a.info.max_fragments_per_voxel) == (b.info.total_fragments, b.info.unique_voxels, b.info.max_fragments_per_voxel)
    rows = np.arange(0, H, 2)
    ia, ib = a.image.reshape(H, W)[rows], r.image.reshape(H, W)[rows]
    covered = b.vis.reshape(H, W)[rows] != np.uint64(0xFFFFFFFFFFFFFFFF)
    assert covered.sum() > 1000 and np.array_equal(ia[covered], ib[covered]) and a.cone_steps == r.cone_steps


# ===================================================== vertex / geometry stages: voxelize.vert/.geom, simple.vert, phong.vert
def _ulp_distance(a, b):
    """distance in representable fp32 values between two arrays (finite values)"""
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia); ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def vertex_stage_scene():
    """pbr_room with rotated, non-uniformly scaled and translated actors (the normal matrix and the TBN re-orthogonalisation matter)"""
    sc = pbr_room()
    rng = np.random.default_rng(31)
    for i in range(1, len(sc.models)):
        a = rng.uniform(0, 2 * np.pi)
        rot = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], np.float32)
        scl = np.diag(np.append(rng.uniform(0.5, 1.5, 3), 1.0)).astype(np.float32)
        m = np.asarray(sc.models[i], np.float32).reshape(4, 4).T                 # column-major storage -> maths layout
        sc.models[i] = np.ascontiguousarray((m @ rot @ scl).T.astype(np.float32))
    return sc


@live
def test_vertex_and_geometry_stages_against_the_reference_glsl():
    """voxelize.vert + voxelize.geom (the voxelisation path): world positions, normals, dominant axis and clip positions
    IDENTICAL.  phong.vert: fragPosition, fragNormal and the re-orthogonalised T / B identical.  simple.vert / phong.vert clip
    positions and lightFragPos: the shaders write `projection * view * model * vec4(position, 1)` (left to right: two matrix
    products first), this repository's passes apply the matrices to the vector one after the other — the same real-number
    result, different fp32 rounding; GLSL guarantees neither order without `precise`.  The difference is bounded here."""
    from vct_b200 import scene as S
    sc = vertex_stage_scene()
    W, H = 96, 64
    p = S.room_params(W, H)
    o = O.Oracle(sc, 32, 5, 64, W, H)
    nv, nt = o.s.verts.shape[0], o.s.tmat.shape[0]
    vox, light, cam, ph = np.zeros(nt * 13, np.float32), np.zeros(nv * 4, np.float32), np.zeros(nv * 4, np.float32), np.zeros(nv * 16, np.float32)
    O.lib().orc_vertex_stage(C.byref(o.s.c), C.byref(p), ptr(vox), ptr(light), ptr(cam), ptr(ph))
    g = glsl()
    vs6 = np.zeros(nv * 6, np.float32); g.glsl_voxelize_vert(C.byref(o.s.c), ptr(vs6))
    wpos, wnrm = np.zeros(nv * 3, np.float32), np.zeros(nv * 3, np.float32)
    O.lib().orc_world_vertices(C.byref(o.s.c), ptr(wpos), ptr(wnrm))
    assert np.array_equal(vs6.reshape(nv, 6)[:, :3].view(np.uint32), wpos.reshape(nv, 3).view(np.uint32))       # model * position
    assert np.array_equal(vs6.reshape(nv, 6)[:, 3:].view(np.uint32), wnrm.reshape(nv, 3).view(np.uint32))       # mat3(transpose(inverse(model))) * normal
    gvox = np.zeros(nt * 13, np.float32); g.glsl_voxelize_geom(C.byref(o.s.c), C.byref(p), ptr(vs6), ptr(gvox))
    assert len(set(vox.reshape(nt, 13)[:, 0].tolist())) == 3                                                       # all three axes occur
    assert np.array_equal(gvox.view(np.uint32), vox.view(np.uint32)), "voxelize.geom: axis / clip positions differ"
    for ax in (0, 1, 2):                                                                                           # axis_override (Application.cpp:713)
        p.axis_override = ax
        O.lib().orc_vertex_stage(C.byref(o.s.c), C.byref(p), ptr(vox), None, None, None)
        g.glsl_voxelize_geom(C.byref(o.s.c), C.byref(p), ptr(vs6), ptr(gvox))
        assert np.array_equal(gvox.view(np.uint32), vox.view(np.uint32)) and (vox.reshape(nt, 13)[:, 0] == ax).all()
    p.axis_override = -1
    g16, g4 = np.zeros(nv * 16, np.float32), np.zeros(nv * 4, np.float32)
    g.glsl_phong_vert(C.byref(o.s.c), C.byref(p), ptr(g16), ptr(g4))
    a, b = ph.reshape(nv, 16), g16.reshape(nv, 16)
    exact = [0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 14, 15]                                                             # fragPosition, fragNormal, T, B
    finite = np.isfinite(a[:, exact]).all(axis=1)
    assert finite.sum() > 0.9 * nv
    assert np.array_equal(a[finite][:, exact].view(np.uint32), b[finite][:, exact].view(np.uint32)), "phong.vert world-space outputs differ"
    assert np.array_equal(np.isnan(a[:, exact]), np.isnan(b[:, exact]))
    gl4 = np.zeros(nv * 4, np.float32)
    g.glsl_simple_vert(C.byref(o.s.c), p.lp, p.lv, ptr(gl4))
    gc4 = np.zeros(nv * 4, np.float32)
    g.glsl_simple_vert(C.byref(o.s.c), p.projection, p.view, ptr(gc4))
    assert np.array_equal(gc4.view(np.uint32), g4.view(np.uint32))                                                 # simple.vert and phong.vert agree with each other
    report = {}
    for name, ours, ref in (("shadow clip", light, gl4), ("camera clip", cam, gc4), ("lightFragPos", a[:, 6:10].reshape(-1).copy(), b[:, 6:10].reshape(-1).copy())):
        d = _ulp_distance(ours, ref)
        scale = np.abs(ref).reshape(-1, 4).max(axis=1).repeat(4)                                                    # cancellation: compare against the vector's magnitude
        rel = np.abs(ours.astype(np.float64) - ref.astype(np.float64)) / np.maximum(scale, 1e-30)
        report[name] = (int((d != 0).sum()), int(d.size), float(rel.max()))
        assert rel.max() < 4e-7, (name, report[name])                                                              # a few units of 2^-24 relative to the vector
    print("association-order differences (values differing, of, max relative):", report)
    # with the shaders' own association switched on in the oracle (tests only) nothing differs any more
    O.lib().orc_set_literal_vertex_transforms(1)
    try:
        O.lib().orc_vertex_stage(C.byref(o.s.c), C.byref(p), None, ptr(light), ptr(cam), ptr(ph))
    finally:
        O.lib().orc_set_literal_vertex_transforms(0)
    assert np.array_equal(light.view(np.uint32), gl4.view(np.uint32)) and np.array_equal(cam.view(np.uint32), gc4.view(np.uint32))
    assert np.array_equal(ph.reshape(nv, 16)[:, 6:10].view(np.uint32), b[:, 6:10].view(np.uint32))


def test_association_order_of_the_clip_transforms_changes_no_pixel_class():
    """End-to-end effect of the one known deviation: the same frame with the oracle's stepwise transforms and
    with the shaders' literal ones.  Shadow-map depths move by ulps, coverage does not change, the image stays far above the gate."""
    from vct_b200 import scene as S
    D, Lv, SS, W, H = 32, 5, 256, 96, 64
    sc = vertex_stage_scene()
    p = S.room_params(W, H)
    a, b = O.Oracle(sc, D, Lv, SS, W, H), O.Oracle(sc, D, Lv, SS, W, H)
    a.frame(p)
    O.lib().orc_set_literal_vertex_transforms(1)
    try:
        b.frame(p)
    finally:
        O.lib().orc_set_literal_vertex_transforms(0)
    covered_a, covered_b = a.shadow < 1.0, b.shadow < 1.0
    assert np.array_equal(covered_a, covered_b), "shadow-map coverage changed"
    d = _ulp_distance(a.shadow[covered_a], b.shadow[covered_b])
    assert d.max() <= 16, d.max()                                          # depths agree to a few ulp
    assert np.array_equal(a.color[0], b.color[0])                         # the voxelisation path does not depend on it
    same_tri = (a.vis & np.uint64(0xFFFFFFFF)) == (b.vis & np.uint64(0xFFFFFFFF))
    assert same_tri.mean() > 0.999, same_tri.mean()                       # the visible triangle of (almost) every pixel is the same
    x = a.image.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64); y = b.image.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    mse = ((x - y) ** 2).mean()
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"literal vs stepwise transforms: shadow texels differing {(d != 0).mean():.3f} (max {d.max()} ulp), visibility pixels differing {(~same_tri).sum()}, "
          f"radiance words differing {(a.radiance[0] != b.radiance[0]).sum()}, image PSNR {psnr:.1f} dB")
    assert psnr >= 60.0


# ===================================================== a0 / a0': the alpha tests of reflectiveShadowMap.frag and dither.frag
def alpha_scenes():
    """(name, scene, params, S, W, H): the room's alpha-masked quad with its binary hole mask, and the same quad with a smooth noise
    alpha map whose values cross the 0.1 threshold everywhere (many fragments land within a few 1/255 of it); near and far cameras
    so that magnified (bilinear) and minified (mip-mapped) lookups both occur."""
    from vct_b200 import scene as S
    out = []
    for name, noise in (("holes", False), ("noise", True)):
        sc = S.room_scene()
        if noise:
            rng = np.random.default_rng(5)
            a = rng.random((8, 8)).astype(np.float32)
            a = np.kron(a, np.ones((8, 8), np.float32))
            for _ in range(6):                                   # smooth: a field that hovers around the threshold
                a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5
            a = np.clip((a - a.mean()) * 120 + 26, 0, 255).astype(np.uint8)
            t = sc.add_texture(a)
            for m in sc.materials:
                if m.alpha_tex >= 0:
                    m.alpha_tex = t
        for cam_name, pos, W, H, SS in (("near", (0.9, 0.0, 1.6), 96, 64, 256), ("far", (1.4, 1.2, 1.45), 48, 32, 64)):
            p = S.room_params(W, H)
            cam = P.Camera(position=pos, front=(0.45 - pos[0], -0.15 - pos[1], 0.7 - pos[2]))
            q = P.default_params(W, H, cam, P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1),
                                 voxel_min=-1.5, voxel_max=1.5)
            p.projection, p.view, p.pv, p.eye = q.projection, q.view, q.pv, q.eye
            out.append((f"{name}_{cam_name}", sc, p, SS, W, H))
    return out


def run_alpha_cases(impl):
    out = {}
    for name, sc, p, SS, W, H in alpha_scenes():
        o = O.Oracle(sc, 32, 5, SS, W, H)
        cap = 1 << 17
        rec = np.zeros(cap * 6, np.float32)
        O.lib().orc_alpha_trace(ptr(rec), C.c_longlong(cap))
        try:
            o.shadowmap(p); o.visibility(p)
            n = O.lib().orc_alpha_trace_count()
        finally:
            O.lib().orc_alpha_trace(None, C.c_longlong(0))
        assert 0 < n <= cap, n
        r = rec[:6 * n].reshape(n, 6)
        r = np.ascontiguousarray(r[np.lexsort(r.view(np.uint32).T[::-1])])       # raster bands append in any order: canonical order
        for ps, fn in ((0, "glsl_rsm_fragments"), (1, "glsl_dither_fragments")):
            q = np.ascontiguousarray(r[r[:, 0] == ps])
            assert len(q) > 50, (name, ps, len(q))
            if impl == "glsl":
                kept = np.zeros(len(q), np.uint8)
                getattr(glsl(), fn)(C.byref(o.s.c), ptr(q), C.c_longlong(len(q)), ptr(kept))
            else:
                kept = q[:, 5].astype(np.uint8)
            out[f"alpha_{name}_pass{ps}_kept"] = kept
            out[f"alpha_{name}_pass{ps}_uv_rho2"] = np.ascontiguousarray(q[:, 2:5]).view(np.uint32).reshape(-1)
    return out


GOLD_ALPHA = os.path.join(ROOT, "tests", "golden", "glsl_ref_alpha.npz")


@live
def test_oracle_alpha_tests_equal_the_reference_glsl_compiled_as_cpp():
    a, b = run_alpha_cases("oracle"), run_alpha_cases("glsl")
    compare(a, b, "oracle vs compiled reflectiveShadowMap.frag / dither.frag")
    kept = np.concatenate([v for k, v in a.items() if k.endswith("_kept")])
    assert 0.05 < kept.mean() < 0.95, kept.mean()                      # both outcomes occur
    print(f"{kept.size} alpha-tested fragments, {kept.mean():.3f} kept")


def test_oracle_alpha_tests_equal_the_committed_outputs_of_the_reference_glsl():
    compare(run_alpha_cases("oracle"), dict(np.load(GOLD_ALPHA)), "oracle vs tests/golden/glsl_ref_alpha.npz")


# ============================================================ debugVoxels.vert / .geom (N3: Application::debugVoxels, Application.cpp:1222-1275)
# Vertex + geometry stage of the instanced cube renderer: per instance id the voxel's texture coordinate, colour (textureLod at Settings::miplevel),
# world position and the 21 clip-space vertices of its triangle strip.  Rasterisation, depth test and culling are fixed function: the oracle's.
def run_debug_voxel_cases(impl):
    from vct_b200 import scene as S
    out = {}
    rng = np.random.default_rng(321)
    for name, D, Lv, lod, ids in (("d64_point", 64, 5, 0.0, rng.integers(0, 64 ** 3, 3000)), ("d64_lod1_6", 64, 5, 1.6, rng.integers(0, 64 ** 3, 3000)),
                                  # 512^3 instances exceed 2^24: float(gl_InstanceID) rounds and ids collapse onto even multiples (reference quirk)
                                  ("d512_float_ids", 512, 1, 0.0, np.concatenate([rng.integers(0, 512 ** 3, 2000), np.arange(2 ** 24 - 4, 2 ** 24 + 12), np.arange(2 ** 26 + 1, 2 ** 26 + 19)]))):
        ids = np.ascontiguousarray(ids, np.uint32)
        p = S.room_params(320, 240); p.miplevel = lod
        if D == 512:
            pyr = np.zeros(D ** 3, np.uint32); pyr[ids[::3]] = 0xFF336699        # virtual zeros; a few texels set
        else:
            lv = [voxel_volume(D, 17, fill=0.3, counts=False)]
            for l in range(Lv - 1):
                d = D >> l
                dst = np.zeros((d // 2) ** 3, np.uint32); O.lib().orc_mip(d, ptr(lv[-1]), ptr(dst), 0); lv.append(dst)
            pyr = np.concatenate(lv)
        n = len(ids)
        world, color, clip, emitted = np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 84), np.float32), np.zeros(n, np.int32)
        if impl == "glsl":
            v7 = np.zeros((n, 7), np.float32)
            glsl().glsl_debug_voxels_vert(C.byref(p), D, Lv, ptr(pyr), ptr(ids), n, ptr(v7))
            g85 = np.zeros((n, 85), np.float32)
            glsl().glsl_debug_voxels_geom(C.byref(p), D, ptr(v7), n, ptr(g85))
            world, color, emitted, clip = v7[:, :3].copy(), v7[:, 3:].copy(), g85[:, 0].astype(np.int32), g85[:, 1:].copy()
        else:
            for i, vid in enumerate(ids):
                O.lib().orc_debug_voxel_vertices(C.byref(p), D, C.c_uint(int(vid)), ptr(world[i]), ptr(clip[i]))
                O.lib().orc_debug_voxel_color(C.byref(p), D, Lv, ptr(pyr), C.c_uint(int(vid)), ptr(color[i]))
            emitted = np.where(color[:, 3] > 0, 21, 0).astype(np.int32)              # debugVoxels.geom:46: nothing for an empty voxel
            clip[emitted == 0] = 0
        out[f"dbgvox_{name}_world"] = world.view(np.uint32).reshape(-1)
        out[f"dbgvox_{name}_color"] = color.view(np.uint32).reshape(-1)
        out[f"dbgvox_{name}_emitted"] = emitted
        out[f"dbgvox_{name}_clip"] = np.ascontiguousarray(clip).view(np.uint32).reshape(-1)
    return out


GOLD_DBGVOX = os.path.join(ROOT, "tests", "golden", "glsl_ref_debug_voxels.npz")


@live
def test_oracle_debug_voxel_stages_equal_the_reference_glsl_compiled_as_cpp():
    a, b = run_debug_voxel_cases("oracle"), run_debug_voxel_cases("glsl")
    compare(a, b, "oracle vs compiled debugVoxels.vert / debugVoxels.geom")
    assert 0 < (a["dbgvox_d64_point_emitted"] > 0).mean() < 1 and (a["dbgvox_d64_lod1_6_emitted"] > 0).mean() > 0.9
    w = a["dbgvox_d512_float_ids_world"].reshape(-1, 3)
    assert len(np.unique(w[2000:2016], axis=0)) < 16, "ids around 2^24 must collapse (float(gl_InstanceID))"


def test_oracle_debug_voxel_stages_equal_the_committed_outputs_of_the_reference_glsl():
    compare(run_debug_voxel_cases("oracle"), dict(np.load(GOLD_DBGVOX)), "oracle vs tests/golden/glsl_ref_debug_voxels.npz")
