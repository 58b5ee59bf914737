"""ctypes binding of the CPU oracle (oracle/).  TEST INFRASTRUCTURE: imported only by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

from vct_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libvct_oracle.so")


class OrcTexture(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("levels", C.c_int),
                ("offset", C.c_longlong * 16)]


class OrcScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_actor", C.c_void_p), ("n_vertices", C.c_int),
                ("indices", C.c_void_p), ("tri_material", C.c_void_p), ("n_tris", C.c_int),
                ("actor_model", C.c_void_p), ("n_actors", C.c_int),
                ("materials", C.c_void_p), ("n_materials", C.c_int),
                ("textures", C.c_void_p), ("n_textures", C.c_int), ("texels", C.c_void_p),
                ("lights", C.c_void_p), ("n_lights", C.c_int)]


def build(force=False):
    src = os.path.join(ROOT, "oracle", "vct_oracle.cpp")
    if force or not os.path.isfile(SO) or os.path.getmtime(SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "oracle", "vct_oracle.h")),
            os.path.getmtime(os.path.join(ROOT, "include", "vct_b200.h"))):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_rgba8_avg.restype = C.c_uint
        _lib.orc_rgba8_avg.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float]
        _lib.orc_pack_unorm4x8.restype = C.c_uint
        _lib.orc_pack_unorm4x8.argtypes = [C.c_float] * 4
        _lib.orc_cone_trace_const.restype = C.c_float
        _lib.orc_warp_weight_table.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_temporal_radiance_filter.argtypes = [C.c_int, C.c_float, C.c_void_p]
        _lib.orc_set_voxel_opacity.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_normalize_voxels_f16.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleScene:
    """Keeps the flattened arrays of a vct_b200.scene.Scene alive and exposes the orc_scene struct."""

    def __init__(self, scene):
        self.verts, self.vact, self.idx, self.tmat, self.models = scene.flat()
        self.mats = (P.Material * max(1, len(scene.materials)))(*scene.materials)
        self.lights = (P.Light * max(1, len(scene.lights)))(*scene.lights)
        self.tex = (OrcTexture * max(1, len(scene.textures)))()
        blobs, off = [], 0
        for i, t in enumerate(scene.textures):
            self.tex[i].width, self.tex[i].height, self.tex[i].channels, self.tex[i].levels = t.width, t.height, t.channels, min(16, len(t.levels))
            for l, lv in enumerate(t.levels[:16]):
                self.tex[i].offset[l] = off; blobs.append(lv.reshape(-1)); off += lv.size
        self.texels = np.concatenate(blobs) if blobs else np.zeros(1, np.uint8)
        s = OrcScene()
        s.vertices, s.vertex_actor, s.n_vertices = ptr(self.verts), ptr(self.vact), len(self.verts)
        s.indices, s.tri_material, s.n_tris = ptr(self.idx), ptr(self.tmat), len(self.tmat)
        s.actor_model, s.n_actors = ptr(self.models), len(self.models)
        s.materials, s.n_materials = C.cast(self.mats, C.c_void_p), len(scene.materials)
        s.textures, s.n_textures, s.texels = C.cast(self.tex, C.c_void_p), len(scene.textures), ptr(self.texels)
        s.lights, s.n_lights = C.cast(self.lights, C.c_void_p), len(scene.lights)
        self.c = s
        self.light0 = scene.lights[0] if scene.lights else None


def level_dims(D, L):
    return [max(1, D >> l) for l in range(L)]


class Oracle:
    """Pass-by-pass driver mirroring the order of Application::render (reference src/Application.cpp:196-1085)."""

    def __init__(self, scene, dim, levels, shadow_size, width, height):
        self.s = OracleScene(scene)
        self.D, self.L, self.S, self.W, self.H = dim, levels, shadow_size, width, height
        n = dim ** 3
        self.shadow = np.ones(shadow_size * shadow_size, np.float32)
        self.color = [np.zeros(d ** 3, np.uint32) for d in level_dims(dim, levels)]
        self.radiance = [np.zeros(d ** 3, np.uint32) for d in level_dims(dim, levels)]
        self.normal = np.zeros(n, np.uint32)
        self.occ = np.zeros(32 ** 3, np.uint32)
        self.warpmap = np.zeros(32 ** 3 * 4, np.uint16)
        self.wlo = np.zeros(32 ** 3 * 4, np.uint16); self.whi = np.zeros(32 ** 3 * 4, np.uint16)
        self.vis = np.zeros(width * height, np.uint64)
        self.image = np.zeros(width * height, np.uint32)
        self.info = P.VoxelizeInfo()
        self.cone_steps = 0

    def set_actor_transform(self, actor, model):
        """Per-frame actor animation (Scene::update): overwrite the model matrix in place (the orc_scene keeps its pointer)."""
        self.s.models[actor] = np.asarray(model, np.float32).reshape(16)

    def _wm(self, p):
        return ptr(self.warpmap) if p.warp_texture else None

    def shadowmap(self, p):
        lib().orc_shadowmap(C.byref(self.s.c), C.byref(p), self.S, ptr(self.shadow))

    def occupancy(self, p):
        lib().orc_occupancy(C.byref(self.s.c), C.byref(p), ptr(self.occ))

    def warpmap_pass(self, p):
        lib().orc_warpmap(ptr(self.occ), C.byref(p), ptr(self.warpmap), ptr(self.wlo), ptr(self.whi))

    def voxelize(self, p):
        if p.voxelize_tesselation:            # the reference's default voxeliser (Application.cpp:585-665)
            lib().orc_voxelize_tess(C.byref(self.s.c), C.byref(p), self.D, ptr(self.color[0]), ptr(self.normal), C.byref(self.info))
            return
        lib().orc_voxelize(C.byref(self.s.c), C.byref(p), self.D, ptr(self.shadow), self.S, self._wm(p),
                           ptr(self.color[0]), ptr(self.normal), C.byref(self.info))

    def transfer(self, p):
        lib().orc_transfer(C.byref(p), self.D, ptr(self.color[0]), ptr(self.radiance[0]), C.byref(self.info))

    def inject(self, p):
        l0 = self.s.light0
        lp = (C.c_float * 3)(*l0.position); li = (C.c_float * 3)(*l0.color)
        lib().orc_inject(C.byref(p), self.D, ptr(self.color[0]), ptr(self.normal), ptr(self.shadow), self.S, self._wm(p),
                         lp, li, ptr(self.radiance[0]))

    def fill_holes(self):
        lib().orc_fill_holes(self.D, ptr(self.radiance[0]))

    def mip(self, which="radiance", mode=0):
        vol = self.radiance if which == "radiance" else self.color
        for l in range(self.L - 1):                     # levels 1..L-1 (Application.cpp:889-902, last write is invalid)
            lib().orc_mip(max(1, self.D >> l), ptr(vol[l]), ptr(vol[l + 1]), mode)

    def visibility(self, p):
        lib().orc_visibility(C.byref(self.s.c), C.byref(p), self.W, self.H, ptr(self.vis))

    def shade(self, p, y_lo=0, y_hi=None, y_stride=1):
        steps = C.c_ulonglong(0)
        rad = np.concatenate(self.radiance); col = np.concatenate(self.color)
        lib().orc_set_normal_volume(ptr(self.normal))            # VIEW_VOXEL_NORMALS samples voxelNormal (texture unit 1)
        lib().orc_shade_rows(C.byref(self.s.c), C.byref(p), self.W, self.H, y_lo, self.H if y_hi is None else y_hi, y_stride,
                             ptr(self.vis), self.D, self.L, ptr(rad), ptr(col),
                             ptr(self.shadow), self.S, self._wm(p), ptr(self.image), C.byref(steps))
        self.cone_steps = steps.value

    def debug_voxels(self, p):
        """Application::debugVoxels: cubes of the traced pyramid's non-empty voxels into self.image."""
        vol = self.radiance if p.draw_radiance else self.color
        lib().orc_debug_voxels(C.byref(p), self.W, self.H, self.D, self.L, ptr(np.concatenate(vol)), ptr(self.image))

    def frame(self, p):
        self.shadowmap(p)
        if p.warp_texture:
            self.occupancy(p); self.warpmap_pass(p)
        self.voxelize(p); self.transfer(p); self.inject(p)
        if p.voxel_fill_holes:
            self.fill_holes()
        self.mip("radiance")
        if p.mip_color_chain:
            self.mip("color")
        self.visibility(p); self.shade(p)

    def image_rgba(self):
        return self.image.view(np.uint8).reshape(self.H, self.W, 4)[::-1]      # flip: row 0 of the buffer is the bottom


# ------------------------------------------------------------------------------------------------------------------
GLSL_SO = os.path.join(ROOT, "oracle", "_ref", "libvct_glsl_ref.so")


class GlslReference:
    """The GI frame on the CPU with the REFERENCE'S OWN SHADERS: oracle/_ref/libvct_glsl_ref.so holds the reference's GLSL
    (transferVoxels.comp, injectRadiance.comp, filterRadiance.comp, voxelize.frag, phong.frag, ...) compiled as C++ from
    where it lies (oracle/Makefile, oracle/ref_rig/).  What OpenGL's fixed function would do — rasterise, interpolate, sample
    — comes from the oracle's recorders (orc_voxelize_trace / orc_shade_trace_rows with NULL outputs: no shading there).
    Compute dispatches and the per-pixel pass run on all host threads (OpenMP); voxelize.frag's fragments run on one thread in
    canonical order (its running-average atomic is order-dependent).  bench.py times this as cpu_baseline kind "reference"."""

    FRAG_CAP = 1 << 21

    def __init__(self, oracle):
        if not os.path.isfile(GLSL_SO):
            raise FileNotFoundError(GLSL_SO)
        self.o = oracle
        self.g = C.CDLL(GLSL_SO)
        o = oracle
        self.frag = np.zeros(self.FRAG_CAP * 16, np.float32)
        self.prec = np.zeros(o.W * o.H * 28, np.float32)
        self.image = np.zeros(o.W * o.H, np.uint32)
        self.cone_steps = 0

    def voxelize(self, p):
        o, n = self.o, C.c_longlong(0)
        lib().orc_voxelize_trace(C.byref(o.s.c), C.byref(p), o.D, ptr(o.shadow), o.S, o._wm(p), None, None, C.byref(o.info),
                                 ptr(self.frag), C.c_longlong(self.FRAG_CAP), C.byref(n))
        if n.value > self.FRAG_CAP:
            raise RuntimeError("fragment record buffer too small")
        self.g.glsl_voxelize_fragments(C.byref(o.s.c), C.byref(p), o.D, ptr(o.shadow), o.S, o._wm(p), ptr(self.frag), C.c_longlong(n.value),
                                       ptr(o.color[0]), ptr(o.normal), C.byref(o.info))

    def transfer(self, p):
        o = self.o
        self.g.glsl_transfer(C.byref(p), o.D, ptr(o.color[0]), ptr(o.radiance[0]), C.byref(o.info))

    def inject(self, p):
        o, l0 = self.o, self.o.s.light0
        lp = (C.c_float * 3)(*l0.position); li = (C.c_float * 3)(*l0.color)
        self.g.glsl_inject(C.byref(p), o.D, ptr(o.color[0]), ptr(o.normal), ptr(o.shadow), o.S, o._wm(p), lp, li, ptr(o.radiance[0]))

    def mip(self, which="radiance"):
        o = self.o
        vol = o.radiance if which == "radiance" else o.color
        for l in range(o.L - 1):
            self.g.glsl_mip(max(1, o.D >> l), ptr(vol[l]), ptr(vol[l + 1]), 0)

    def shade(self, p, y_lo=0, y_hi=None, y_stride=1):
        o = self.o
        self.prec[23::28] = 2.0                                         # "row not sampled"; the recorder flags covered pixels of sampled rows with 1
        steps = C.c_ulonglong(0)
        rad = np.concatenate(o.radiance); col = np.concatenate(o.color)
        lib().orc_shade_trace_rows(C.byref(o.s.c), C.byref(p), o.W, o.H, y_lo, o.H if y_hi is None else y_hi, y_stride, ptr(o.vis), o.D, o.L,
                                   ptr(rad), ptr(col), ptr(o.shadow), o.S, o._wm(p), None, None, ptr(self.prec))
        self.g.glsl_shade_pixels(C.byref(o.s.c), C.byref(p), o.W, o.H, ptr(self.prec), o.D, o.L, ptr(rad), ptr(col), ptr(o.shadow), o.S, o._wm(p),
                                 ptr(self.image), C.byref(steps))
        self.cone_steps = steps.value
