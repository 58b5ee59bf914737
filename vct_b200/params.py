"""ctypes mirror of include/vct_b200.h plus the reference's `Settings` / `VCT` / `Camera` defaults.

Host-side mirror of the uniforms Application::render uploads (reference src/Application.h:36-156,
src/Application.cpp:196-210, 689-692, 804).  Matrix helpers follow GLM 0.9.9 (right-handed, depth -1..1,
column-major) because the reference computes every matrix on the CPU with GLM and uploads it as a uniform.
"""
import ctypes as C
import math

import numpy as np

WARP_DIM = 32
F16 = C.c_float * 16
F3 = C.c_float * 3

VIEW_SHADED, VIEW_VOXELS, VIEW_MATERIAL_DIFFUSE, VIEW_MATERIAL_ROUGHNESS, VIEW_MATERIAL_METALLIC = 0, 1, 2, 3, 4
VIEW_NORMALS, VIEW_DOMINANT_AXIS, VIEW_INDIRECT, VIEW_OCCLUSION, VIEW_REFLECTIONS = 5, 6, 7, 8, 9
VIEW_VOXEL_NORMALS, VIEW_WARP_TEXTURE, VIEW_WARP_TEXTURE_TC = 10, 11, 12
VOL_COLOR, VOL_NORMAL, VOL_RADIANCE, VOL_OCCUPANCY, VOL_WARPMAP, VOL_WARP_WEIGHTS_LOW, VOL_WARP_WEIGHTS_HIGH, BUF_IMAGE = range(8)


class Config(C.Structure):
    _fields_ = [("dim", C.c_int), ("levels", C.c_int), ("shadow_size", C.c_int), ("width", C.c_int),
                ("height", C.c_int), ("device", C.c_int), ("rank", C.c_int), ("world_size", C.c_int),
                ("max_fragments", C.c_int), ("n_devices", C.c_int), ("devices", C.POINTER(C.c_int)), ("slab_stripe", C.c_int)]


EXCHANGE_HANDLE_BYTES = 384


class Peer(C.Structure):          # vct_peer
    _fields_ = [("staging", C.c_void_p), ("radiance", C.c_void_p), ("color", C.c_void_p), ("image", C.c_void_p), ("shadow", C.c_void_p)]


class Light(C.Structure):
    _fields_ = [("position", F3), ("_pad0", C.c_float), ("direction", F3), ("_pad1", C.c_float),
                ("color", F3), ("range", C.c_float), ("intensity", C.c_float), ("enabled", C.c_int),
                ("selected", C.c_int), ("shadow_caster", C.c_int), ("type", C.c_uint), ("_pad2", F3)]


assert C.sizeof(Light) == 80   # reference src/Scene.h:33 glslSize


class Material(C.Structure):
    _fields_ = [("diffuse_tex", C.c_int), ("specular_tex", C.c_int), ("normal_tex", C.c_int),
                ("roughness_tex", C.c_int), ("metallic_tex", C.c_int), ("alpha_tex", C.c_int),
                ("shininess", C.c_float), ("diffuse", F3)]


class IngestMesh(C.Structure):        # vct_ingest_mesh
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("n_vertices", C.c_size_t), ("indices", C.POINTER(C.c_uint32)),
                ("n_indices", C.c_size_t), ("material_of_triangle", C.POINTER(C.c_int32)), ("n_materials", C.c_int),
                ("n_textures", C.c_int), ("bounds_min", F3), ("bounds_max", F3), ("radius", C.c_float)]


class IngestTexture(C.Structure):     # vct_ingest_texture
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("levels", C.c_int),
                ("pixels", C.c_void_p), ("bytes", C.c_size_t), ("name", C.c_char_p)]


class VoxelizeInfo(C.Structure):
    _fields_ = [("total_fragments", C.c_uint), ("unique_voxels", C.c_uint), ("max_fragments_per_voxel", C.c_uint)]


class ConeSettings(C.Structure):
    _fields_ = [("steps", C.c_int), ("cone_angle", C.c_float), ("bias", C.c_float),
                ("cone_initial_height", C.c_float), ("lod_offset", C.c_float)]


class FrameParams(C.Structure):
    _fields_ = [
        ("projection", F16), ("view", F16), ("pv", F16), ("lp", F16), ("lv", F16), ("ls", F16),
        ("ls_inverse", F16), ("mvp_x", F16), ("mvp_y", F16), ("mvp_z", F16),
        ("eye", F3), ("voxel_min", F3), ("voxel_max", F3), ("voxel_center", F3), ("clear_color", F3),
        ("voxelize_lighting", C.c_int), ("voxelize_atomic_max", C.c_int), ("axis_override", C.c_int),
        ("deterministic", C.c_int), ("voxel_set_opacity", C.c_float),
        ("temporal_filter_radiance", C.c_int), ("temporal_decay", C.c_float),
        ("radiance_lighting", C.c_int), ("radiance_dilate", C.c_int), ("voxel_fill_holes", C.c_int),
        ("mip_color_chain", C.c_int),
        ("warp_voxels", C.c_int), ("warp_texture", C.c_int), ("warp_texture_linear", C.c_int),
        ("warp_texture_axes", C.c_int * 3), ("use_warpmap_weights_texture", C.c_int),
        ("warp_texture_high_resolution", C.c_float), ("warp_texture_low_resolution", C.c_float),
        ("draw_radiance", C.c_int), ("draw_occlusion", C.c_int), ("cooktorrance", C.c_int),
        ("enable_postprocess", C.c_int), ("enable_normal_map", C.c_int),
        ("enable_indirect", C.c_int), ("enable_diffuse", C.c_int), ("enable_specular", C.c_int),
        ("enable_reflections", C.c_int), ("ambient_scale", C.c_float), ("reflect_scale", C.c_float),
        ("diffuse_cone", ConeSettings), ("specular_cone", ConeSettings),
        ("specular_cone_angle_from_roughness", C.c_int),
        ("debug_view", C.c_int), ("miplevel", C.c_float), ("voxelize_tesselation", C.c_int), ("voxelize_tesselation_warp", C.c_int),
        ("conservative_raster", C.c_int), ("msaa_samples", C.c_float * 8), ("voxelize_multiplier", C.c_float),
    ]


RASTER_CENTER, RASTER_MSAA = 0, 1
MSAA_STANDARD_4X = (0.375, 0.125, 0.875, 0.375, 0.125, 0.625, 0.625, 0.875)   # glGetMultisamplefv on NVIDIA GL / the D3D standard pattern


class Timings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("voxelize_ns", "shadowmap_ns", "radiance_ns", "mipmap_ns", "render_ns",
                                          "total_ns", "transfer_ns", "gbuffer_ns", "warpmap_ns", "clear_ns",
                                          "exchange_ns")]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ns", C.c_double), ("launches", C.c_uint), ("_pad", C.c_uint)]


# --------------------------------------------------------------------------------- GLM-equivalent helpers
f32 = np.float32


def _v(x):
    return np.asarray(x, dtype=f32)


def _normalize(v):
    return (v / np.sqrt(np.dot(v, v), dtype=f32)).astype(f32)


def perspective(fovy, aspect, near, far):
    """glm::perspective (RH, NO).  The reference passes fov=45.0f un-converted, i.e. 45 RADIANS
    (src/Camera.h:18, src/Application.cpp:200) — replicate, do not fix."""
    t = f32(math.tan(f32(fovy) / f32(2)))
    m = np.zeros((4, 4), f32)          # m[col][row]
    m[0][0] = f32(1) / (f32(aspect) * t)
    m[1][1] = f32(1) / t
    m[2][2] = -(f32(far) + f32(near)) / (f32(far) - f32(near))
    m[2][3] = -1
    m[3][2] = -(f32(2) * f32(far) * f32(near)) / (f32(far) - f32(near))
    return m


def ortho(l, r, b, t, n, f):
    l, r, b, t, n, f = map(f32, (l, r, b, t, n, f))
    m = np.eye(4, dtype=f32)
    m[0][0] = f32(2) / (r - l)
    m[1][1] = f32(2) / (t - b)
    m[2][2] = -f32(2) / (f - n)
    m[3][0] = -(r + l) / (r - l)
    m[3][1] = -(t + b) / (t - b)
    m[3][2] = -(f + n) / (f - n)
    return m


def look_at(eye, center, up):
    eye, center, up = _v(eye), _v(center), _v(up)
    f = _normalize(center - eye)
    s = _normalize(np.cross(f, up).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.eye(4, dtype=f32)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -np.dot(s, eye)
    m[3][1] = -np.dot(u, eye)
    m[3][2] = np.dot(f, eye)
    return m


def matmul(a, b):
    """GLM a*b for m[col][row] arrays: (a*b)[c][r] = sum_k a[k][r] * b[c][k]."""
    return (b.astype(f32) @ a.astype(f32)).astype(f32)


def inverse(m):
    return np.linalg.inv(m.astype(np.float64).T).T.astype(f32)


def scale_matrix(s):
    m = np.eye(4, dtype=f32)
    s = _v(s) if np.ndim(s) else _v([s, s, s])
    m[0][0], m[1][1], m[2][2] = s
    return m


def translate_matrix(t):
    m = np.eye(4, dtype=f32)
    m[3][0], m[3][1], m[3][2] = _v(t)
    return m


def mat_to_c(m):
    return F16(*[float(x) for x in np.asarray(m, f32).reshape(16)])


class Camera:
    """reference src/Camera.h / Camera.cpp (position, yaw/pitch in degrees, fov 45.0 used as radians)."""

    def __init__(self, position=(0, 0, 0), yaw=-90.0, pitch=0.0, fov=45.0, front=None):
        self.position = _v(position)
        self.yaw, self.pitch, self.fov = yaw, pitch, fov
        self.up = _v([0, 1, 0])
        self._front = None if front is None else _normalize(_v(front))

    @property
    def front(self):
        if self._front is not None:
            return self._front
        cy, sy = math.cos(math.radians(self.yaw)), math.sin(math.radians(self.yaw))
        cp, sp = math.cos(math.radians(self.pitch)), math.sin(math.radians(self.pitch))
        return _v([cp * cy, sp, cp * sy])

    def look_at(self):
        return look_at(self.position, self.position + self.front, self.up)


def make_light(position=(0, 0, 0), direction=(0, 0, -1), color=(1, 1, 1), range_=5.0, intensity=1.0, enabled=True,
               shadow_caster=False, type_=0):
    """reference struct Light defaults, src/Scene.h:13-29."""
    L = Light()
    L.position = F3(*position); L.direction = F3(*direction); L.color = F3(*color)
    L.range = range_; L.intensity = intensity; L.enabled = int(enabled); L.selected = 0
    L.shadow_caster = int(shadow_caster); L.type = type_
    return L


def reference_lights():
    """src/Application.cpp:125-136."""
    return [make_light(position=(12.0, 40.0, -7.0), direction=(-0.38, -0.88, 0.2), shadow_caster=True, type_=1),
            make_light(position=(0.0, 10.0, 0.0), color=(1.0, 0.0, 1.0), type_=0)]


def default_params(width, height, camera, light, voxel_min=-20.0, voxel_max=20.0, voxel_center=(0, 0, 0), parity=True):
    """Fill a FrameParams the way Application::render does for one frame.

    parity=True applies SURVEY.md §8 "parity settings" (raster path, running-average atomics, pixel-centre
    coverage); everything else is the reference default (src/Application.h:36-103)."""
    p = FrameParams()
    near, far = 0.1, 100.0                                   # Application.h:172
    aspect = f32(width) / f32(height)
    proj = perspective(camera.fov, aspect, near, far)
    view = camera.look_at()
    pv = matmul(perspective(camera.fov, aspect, 1.0, 20.0), view)
    lp = ortho(-25.0, 25.0, -25.0, 25.0, 0.0, 100.0)         # Application.cpp:207-208
    lpos, ldir = _v(list(light.position)), _v(list(light.direction))
    lv = look_at(lpos, lpos + ldir, [0, 1, 0])
    ls = matmul(lp, lv)
    vmin = _v(voxel_min) if np.ndim(voxel_min) else _v([voxel_min] * 3)
    vmax = _v(voxel_max) if np.ndim(voxel_max) else _v([voxel_max] * 3)
    c = _v(voxel_center)
    vproj = ortho(vmin[0], vmax[0], vmin[1], vmax[1], 0.0, vmax[2] - vmin[2])      # Application.cpp:689
    mvp_x = matmul(vproj, look_at(_v([vmax[0], 0, 0]) + c, c, [0, 1, 0]))
    mvp_y = matmul(vproj, look_at(_v([0, vmax[1], 0]) + c, c, [0, 0, -1]))
    mvp_z = matmul(vproj, look_at(_v([0, 0, vmax[2]]) + c, c, [0, 1, 0]))
    for name, m in (("projection", proj), ("view", view), ("pv", pv), ("lp", lp), ("lv", lv), ("ls", ls),
                    ("ls_inverse", inverse(ls)), ("mvp_x", mvp_x), ("mvp_y", mvp_y), ("mvp_z", mvp_z)):
        setattr(p, name, mat_to_c(m))
    p.eye = F3(*camera.position); p.voxel_min = F3(*vmin); p.voxel_max = F3(*vmax); p.voxel_center = F3(*c)
    p.clear_color = F3(0.5294, 0.8078, 0.9216)               # Application.cpp:41
    p.voxelize_lighting = 1
    p.voxelize_atomic_max = 0 if parity else 1
    p.axis_override = -1
    p.deterministic = 1
    p.voxel_set_opacity = 0.5
    p.temporal_filter_radiance = 0; p.temporal_decay = 0.8
    p.radiance_lighting = 0; p.radiance_dilate = 0; p.voxel_fill_holes = 0
    p.mip_color_chain = 1
    p.warp_voxels = 0; p.warp_texture = 0; p.warp_texture_linear = 0
    p.warp_texture_axes = (C.c_int * 3)(1, 1, 1)
    p.use_warpmap_weights_texture = 1
    p.warp_texture_high_resolution = 2.0; p.warp_texture_low_resolution = 0.5
    p.draw_radiance = 1; p.draw_occlusion = 1; p.cooktorrance = 1; p.enable_postprocess = 1; p.enable_normal_map = 1
    p.enable_indirect = 1; p.enable_diffuse = 1; p.enable_specular = 1; p.enable_reflections = 1
    p.ambient_scale = 1.0; p.reflect_scale = 1.0
    p.diffuse_cone = ConeSettings(16, math.radians(60.0), 1.0, 1.0, 0.5)       # Application.h:96
    p.specular_cone = ConeSettings(32, math.radians(30.0), 1.7, 0.5, 0.1)     # Application.h:97
    p.specular_cone_angle_from_roughness = 1
    p.voxelize_multiplier = 1.0                                               # Application.h:87
    p.conservative_raster = RASTER_CENTER if parity else RASTER_MSAA          # Application.h:62 (reference default MSAA)
    return p
