"""Scene containers for the harness: flat arrays in the reference's `Vertex`/index/material layout.

Inputs of the hot path (reference src/Graphics/Mesh.h:72-76, Mesh.cpp:340-371, Scene.cpp:31-36).  Scenes come
either from assets/_baked/ (tools/bake_assets.py, run where /root/reference exists) or from the procedural
generators below (used when the baked assets are absent and for unit tests).
"""
import os

import numpy as np

from . import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAKED = os.path.join(ROOT, "assets", "_baked")


def build_mips(img):
    """2x2 box filter, round half up — the host supplies the mip chain (vct_upload_texture contract)."""
    levels = [np.ascontiguousarray(img)]
    while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:
        a = levels[-1].astype(np.uint16)
        h, w = a.shape[:2]
        if h > 1 and w > 1:
            b = (a[0:h - h % 2:2, 0:w - w % 2:2] + a[1:h:2, 0:w - w % 2:2] + a[0:h - h % 2:2, 1:w:2] + a[1:h:2, 1:w:2] + 2) >> 2
        elif h > 1:
            b = (a[0:h - h % 2:2] + a[1:h:2] + 1) >> 1
        else:
            b = (a[:, 0:w - w % 2:2] + a[:, 1:w:2] + 1) >> 1
        levels.append(np.ascontiguousarray(b.astype(np.uint8)))
    return levels


class Texture:
    def __init__(self, img):
        img = np.asarray(img, np.uint8)
        if img.ndim == 2:
            img = img[:, :, None]
        assert img.shape[2] in (1, 3, 4)
        self.height, self.width, self.channels = img.shape
        self.levels = build_mips(img)

    def packed(self):
        return np.concatenate([l.reshape(-1) for l in self.levels])


def texture_from_packed(t):
    """dict(width, height, channels, levels, pixels = all levels packed) -> Texture"""
    tex = Texture.__new__(Texture)
    tex.width, tex.height, tex.channels = t["width"], t["height"], t["channels"]
    tex.levels, off = [], 0
    for l in range(t["levels"]):
        w, h = max(1, t["width"] >> l), max(1, t["height"] >> l)
        tex.levels.append(t["pixels"][off:off + w * h * t["channels"]].reshape(h, w, t["channels"]))
        off += w * h * t["channels"]
    return tex


class Mesh:
    """One actor's geometry: vertices (n,14) f32, indices (3T,) u32 in draw order, material id per triangle."""

    def __init__(self, vertices, indices, tri_material):
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 14)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        self.tri_material = np.ascontiguousarray(tri_material, np.int32).reshape(-1)
        assert self.indices.size == 3 * self.tri_material.size


class Scene:
    def __init__(self):
        self.meshes, self.models = [], []       # one per actor
        self.materials, self.textures = [], []  # global ids
        self.lights = []

    def add_texture(self, img):
        self.textures.append(Texture(img))
        return len(self.textures) - 1

    def add_material(self, diffuse=-1, specular=-1, normal=-1, roughness=-1, metallic=-1, alpha=-1, shininess=32.0):
        m = P.Material(diffuse, specular, normal, roughness, metallic, alpha, shininess, P.F3(0, 0, 0))
        self.materials.append(m)
        return len(self.materials) - 1

    def add_actor(self, mesh, model=None):
        self.meshes.append(mesh)
        self.models.append(np.eye(4, dtype=np.float32) if model is None else np.asarray(model, np.float32))
        return len(self.meshes) - 1

    # flattened views (oracle input; the product receives the same data per actor through vct_upload_mesh)
    def flat(self):
        vbase, verts, vact, idx, tmat = 0, [], [], [], []
        for a, m in enumerate(self.meshes):
            verts.append(m.vertices); vact.append(np.full(len(m.vertices), a, np.int32))
            idx.append(m.indices + np.uint32(vbase)); tmat.append(m.tri_material)
            vbase += len(m.vertices)
        return (np.concatenate(verts), np.concatenate(vact), np.concatenate(idx), np.concatenate(tmat),
                np.stack([m.reshape(16) for m in self.models]).astype(np.float32))

    @property
    def n_tris(self):
        return sum(m.tri_material.size for m in self.meshes)


# ------------------------------------------------------------------------------------------ baked assets
def baked_available(name):
    return os.path.isfile(os.path.join(BAKED, name, "vertices.f32"))


def load_baked(scene, name, model=None):
    """Append the baked mesh `name` as a new actor (its materials/textures are appended to the scene).
    Texels come from textures.npz, decoded once at bake time (tools/bake_assets.py) by the library's own PNG reader
    (vct_ingest_image: the reference's stb_image behaviour, tests/test_ingest.py) with its generated mips; loading them
    here needs numpy only, so a process that must not load libvct_b200.so (bench.py --impl reference) can build the scene.
    Without the cache the PNG files are decoded through the library."""
    d = os.path.join(BAKED, name)
    npz = os.path.join(d, "textures.npz")
    cached = np.load(npz) if os.path.isfile(npz) else None
    verts = np.fromfile(os.path.join(d, "vertices.f32"), np.float32).reshape(-1, 14)
    idx = np.fromfile(os.path.join(d, "indices.u32"), np.uint32)
    tmat = np.fromfile(os.path.join(d, "tri_material.i32"), np.int32)
    mat_base, cache = len(scene.materials), {}

    def tex(fn):
        if not fn:
            return -1
        if fn not in cache:
            if cached is not None and fn + "|px" in cached:
                w, h, ch, lv = (int(x) for x in cached[fn + "|meta"])
                t = {"width": w, "height": h, "channels": ch, "levels": lv, "pixels": cached[fn + "|px"]}
            else:
                from . import ingest
                t = ingest.load_image(os.path.join(d, "textures", fn))
            if t["channels"] not in (1, 3, 4):                     # grey + alpha: the reference allocates no storage for it (GLHelper.cpp:194-205)
                cache[fn] = -1
            else:
                scene.textures.append(texture_from_packed(t)); cache[fn] = len(scene.textures) - 1
        return cache[fn]

    for line in open(os.path.join(d, "materials.txt")):
        f = (line.rstrip("\n").split("|") + [""] * 7)[:7]
        scene.add_material(diffuse=tex(f[1]), specular=tex(f[2]), normal=tex(f[3]), roughness=tex(f[4]),
                           metallic=tex(f[5]), alpha=tex(f[6]), shininess=32.0)   # Material(material_t) self-assigns: always 32 (Mesh.h:33-43)
    return scene.add_actor(Mesh(verts, idx, tmat + mat_base), model)


# ------------------------------------------------------------------------------------- procedural scenes
def checker_texture(size=64, cells=8, a=(230, 60, 40), b=(40, 90, 220), alpha_holes=False, seed=0):
    y, x = np.mgrid[0:size, 0:size]
    c = ((x * cells // size) + (y * cells // size)) % 2
    rng = np.random.default_rng(seed)
    img = np.where(c[..., None] == 0, np.array(a, np.uint8), np.array(b, np.uint8)).astype(np.int16)
    img = np.clip(img + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
    if alpha_holes:
        hole = (((x * cells * 2 // size) + (y * cells * 2 // size)) % 3 == 0)
        img = np.dstack([img, np.where(hole, 0, 255).astype(np.uint8)])
    return img


def _tangent_frame(verts, idx):
    """Un-weighted per-face tangent/bitangent accumulation (reference src/Graphics/Mesh.cpp:174-201)."""
    v = verts.copy(); v[:, 8:14] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        for a, b, c in idx.reshape(-1, 3):
            e1, e2 = v[b, 0:3] - v[a, 0:3], v[c, 0:3] - v[a, 0:3]
            d1, d2 = v[b, 6:8] - v[a, 6:8], v[c, 6:8] - v[a, 6:8]
            inv = np.float32(1.0) / (d1[0] * d2[1] - d2[0] * d1[1])
            t = inv * (d2[1] * e1 - d1[1] * e2); bt = inv * (d2[0] * e1 - d1[0] * e2)
            for k in (a, b, c):
                v[k, 8:11] += t; v[k, 11:14] += bt
        for s in (slice(8, 11), slice(11, 14)):
            n = np.sqrt((v[:, s] ** 2).sum(1, dtype=np.float32), dtype=np.float32)
            v[:, s] = v[:, s] / n[:, None]
    return v.astype(np.float32)


def cube_mesh(material=0):
    """Unit cube [-1,1]^3, 24 vertices / 12 triangles, outward CCW faces, per-face normals and UVs."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 1, 0), (0, 0, -1)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (0, 0, 1), (-1, 0, 0)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (-1, 0, 0), (0, 1, 0))]
    verts, idx = [], []
    for n, u, w in faces:
        n, u, w = map(np.array, (n, u, w))
        base = len(verts)
        for (su, sw) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            p = n + su * u + sw * w
            verts.append(list(p) + list(n) + [(su + 1) / 2 * 0.9 + 0.05, (sw + 1) / 2 * 0.9 + 0.05] + [0] * 6)
        # winding: (u x w) must equal n for CCW when viewed from outside
        if np.dot(np.cross(u, w), n) > 0:
            idx += [base, base + 1, base + 2, base, base + 2, base + 3]
        else:
            idx += [base, base + 2, base + 1, base, base + 3, base + 2]
    verts = _tangent_frame(np.array(verts, np.float32), np.array(idx, np.uint32))
    return Mesh(verts, idx, np.full(12, material, np.int32))


def quad_mesh(corners, normal, material=0, uv_scale=1.0):
    """Two triangles over 4 corners given CCW w.r.t. `normal`."""
    uv = [(0, 0), (uv_scale, 0), (uv_scale, uv_scale), (0, uv_scale)]
    verts = [list(c) + list(normal) + list(t) + [0] * 6 for c, t in zip(corners, uv)]
    idx = [0, 1, 2, 0, 2, 3]
    return Mesh(_tangent_frame(np.array(verts, np.float32), np.array(idx, np.uint32)), idx, np.full(2, material, np.int32))


def merge_meshes(meshes):
    vb, vs, ids, tm = 0, [], [], []
    for m in meshes:
        vs.append(m.vertices); ids.append(m.indices + np.uint32(vb)); tm.append(m.tri_material); vb += len(m.vertices)
    return Mesh(np.concatenate(vs), np.concatenate(ids), np.concatenate(tm))


def soup_mesh(n_tris, seed=0x5EED, extent=19.0, sigma=0.117, material=0):
    """Synthetic triangle soup: centres uniform in [-extent,extent]^3, edge vectors
    N(0, sigma^2 I), vertex normals = geometric normal, uv = 0."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n_tris, 3)).astype(np.float32)
    e1 = rng.normal(0, sigma, (n_tris, 3)).astype(np.float32); e2 = rng.normal(0, sigma, (n_tris, 3)).astype(np.float32)
    p = np.stack([c, c + e1, c + e2], 1)
    n = np.cross(e1, e2); n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    v = np.zeros((n_tris, 3, 14), np.float32)
    v[:, :, 0:3] = p; v[:, :, 3:6] = n[:, None, :]
    v[:, :, 8:11] = np.nan; v[:, :, 11:14] = np.nan          # no UVs -> NaN tangents (Mesh.cpp:178,196)
    return Mesh(v.reshape(-1, 14), np.arange(3 * n_tris, dtype=np.uint32), np.full(n_tris, material, np.int32))


def room_scene(seed=1, n_clutter=6, alpha_quad=True):
    """Small procedural test scene inside [-1.5,1.5]^3: a floor, two walls, a few cubes, one alpha-masked quad.
    Everything the hot path branches on is present: textures with mips, alpha test, several materials, 2 lights."""
    s = Scene()
    t0 = s.add_texture(checker_texture(64, 8, seed=seed))
    t1 = s.add_texture(checker_texture(32, 4, (200, 200, 60), (60, 200, 120), seed=seed + 1))
    t2 = s.add_texture(checker_texture(64, 4, (250, 250, 250), (90, 90, 90), alpha_holes=True, seed=seed + 2))
    ta = s.add_texture(checker_texture(64, 4, alpha_holes=True, seed=seed + 2)[:, :, 3])
    m0, m1 = s.add_material(diffuse=t0), s.add_material(diffuse=t1)
    m2 = s.add_material(diffuse=t2, alpha=ta)
    parts = [quad_mesh([(-1.4, -1.0, 1.4), (1.4, -1.0, 1.4), (1.4, -1.0, -1.4), (-1.4, -1.0, -1.4)], (0, 1, 0), m0, 4.0),
             quad_mesh([(-1.4, -1.0, -1.4), (1.4, -1.0, -1.4), (1.4, 1.3, -1.4), (-1.4, 1.3, -1.4)], (0, 0, 1), m1, 2.0),
             quad_mesh([(-1.4, -1.0, 1.4), (-1.4, -1.0, -1.4), (-1.4, 1.3, -1.4), (-1.4, 1.3, 1.4)], (1, 0, 0), m1, 2.0)]
    if alpha_quad:
        parts.append(quad_mesh([(-0.2, -0.9, 0.9), (1.1, -0.9, 0.5), (1.1, 0.6, 0.5), (-0.2, 0.6, 0.9)], (0.29, 0, 0.957), m2, 1.0))
    s.add_actor(merge_meshes(parts))
    rng = np.random.default_rng(seed)
    for i in range(n_clutter):
        sc = rng.uniform(0.12, 0.3); pos = rng.uniform(-0.9, 0.9, 3); pos[1] = -1.0 + sc
        s.add_actor(cube_mesh(m0 if i % 2 else m1), P.matmul(P.translate_matrix(pos), P.scale_matrix(sc)))
    s.lights = [P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1),
                P.make_light(position=(0.0, 0.2, 0.0), color=(1.0, 0.0, 1.0), range_=2.0, type_=0)]
    return s


def room_params(width, height, **kw):
    cam = P.Camera(position=(1.1, 0.3, 1.2), front=(-0.65, -0.25, -0.72))
    s_light = P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1)
    p = P.default_params(width, height, cam, s_light, voxel_min=-1.5, voxel_max=1.5, **kw)
    # the room is ~3 units wide: shrink the light frustum so the shadow map has useful resolution
    lp = P.ortho(-2.5, 2.5, -2.5, 2.5, 0.0, 10.0)
    lpos, ldir = np.array(s_light.position[:], np.float32), np.array(s_light.direction[:], np.float32)
    lv = P.look_at(lpos, lpos + ldir, [0, 1, 0]); ls = P.matmul(lp, lv)
    p.lp, p.lv, p.ls, p.ls_inverse = P.mat_to_c(lp), P.mat_to_c(lv), P.mat_to_c(ls), P.mat_to_c(P.inverse(ls))
    return p


# ---------------------------------------------------------------------------- BASELINE.json configurations
def config_scene(idx):
    """Scenes of SURVEY.md §8(d) configs 1-5.  Returns (scene, camera, voxel (min,max,center), dim, (W,H), extra)."""
    if idx == 1:      # cube, 64^3, 512x512
        s = Scene()
        if baked_available("cube"):
            load_baked(s, "cube", P.scale_matrix(0.5))
        else:
            s.add_actor(cube_mesh(s.add_material(diffuse=s.add_texture(checker_texture(256, 8)))), P.scale_matrix(0.5))
        s.lights = P.reference_lights(); s.lights[1].enabled = 0
        cam = P.Camera(position=(2.5, 1.5, 2.5), front=(-2.5, -1.5, -2.5))
        return s, cam, (-1.5, 1.5, (0, 0, 0)), 64, (512, 512), {}
    if idx == 2:      # bunny, 128^3, 720p, diffuse cones only
        s = Scene()
        if not baked_available("bunny"):
            raise FileNotFoundError("assets/_baked/bunny missing: run tools/bake_assets.py")
        load_baked(s, "bunny")
        s.lights = P.reference_lights(); s.lights[1].enabled = 0
        cam = P.Camera(position=(0.9, 1.6, 1.9), front=(-1.068, -0.4985, -1.915))
        return s, cam, (-0.854, 0.854, (-0.168, 1.1015, -0.015)), 128, (1280, 720), {"enable_reflections": 0}
    if idx in (3, 4):  # Sponza (+ nanosuits), reference camera and lights (Application.cpp:96-141)
        s = Scene()
        if baked_available("sponza_pbr"):
            load_baked(s, "sponza_pbr", P.scale_matrix(0.01))
            data = "sponza_pbr (baked from reference assets)"
        else:
            s = atrium_scene(); data = "procedural atrium (sponza assets absent)"
        if idx == 4 and baked_available("nanosuit"):
            load_baked(s, "nanosuit", P.scale_matrix(0.25)); load_baked(s, "nanosuit", P.scale_matrix(0.2))
        s.lights = P.reference_lights()
        cam = P.Camera(position=(5, 1, 0), yaw=180.0)
        if idx == 3:
            return s, cam, (-20.0, 20.0, (0, 0, 0)), 256, (1920, 1080), {"data": data}
        return s, cam, (-20.0, 20.0, (0, 0, 0)), 512, (3840, 2160), {"data": data, "warp_texture": 1, "temporal_filter_radiance": 1}
    if idx == 5:
        s = Scene()
        s.add_actor(soup_mesh(1 << 20, material=s.add_material(diffuse=s.add_texture(np.full((1, 1, 3), 255, np.uint8)))))
        s.lights = P.reference_lights()
        cam = P.Camera(position=(0, 0, 30), front=(0, 0, -1))
        return s, cam, (-20.0, 20.0, (0, 0, 0)), 512, (3840, 2160), {}
    raise ValueError(idx)


def nanosuit_models(t):
    """Actor animation of config 4 (reference src/Application.cpp:101-116, commented-out controllers)."""
    import math
    a = P.matmul(P.translate_matrix((math.cos(t), 0.0, math.sin(t))), P.scale_matrix(0.25))
    b = P.matmul(P.translate_matrix((2 * math.sin(0.4 * t) - 2, 4.2, 3.0)), P.scale_matrix(0.2))
    return a, b


def atrium_scene(seed=7):
    """Procedural stand-in with Sponza's proportions (floor 30x12, two arcades of columns, upper gallery, curtains):
    used by bench.py only when the baked Sponza is absent."""
    s = Scene()
    tx = [s.add_texture(checker_texture(256, 16, seed=seed + i, a=(200 - 20 * i, 170, 140), b=(120, 100 + 20 * i, 90))) for i in range(4)]
    m = [s.add_material(diffuse=t) for t in tx]
    parts = [quad_mesh([(-15, 0, 6), (15, 0, 6), (15, 0, -6), (-15, 0, -6)], (0, 1, 0), m[0], 16.0),
             quad_mesh([(-15, 0, -6), (15, 0, -6), (15, 12, -6), (-15, 12, -6)], (0, 0, 1), m[1], 8.0),
             quad_mesh([(15, 0, 6), (-15, 0, 6), (-15, 12, 6), (15, 12, 6)], (0, 0, -1), m[1], 8.0),
             quad_mesh([(-15, 0, 6), (-15, 0, -6), (-15, 12, -6), (-15, 12, 6)], (1, 0, 0), m[2], 4.0),
             quad_mesh([(15, 0, -6), (15, 0, 6), (15, 12, 6), (15, 12, -6)], (-1, 0, 0), m[2], 4.0)]
    s.add_actor(merge_meshes(parts))
    for i in range(10):
        for z in (-3.0, 3.0):
            x = -12.0 + i * 2.6
            s.add_actor(cube_mesh(m[3]), P.matmul(P.translate_matrix((x, 2.5, z)), P.scale_matrix((0.35, 2.5, 0.35))))
            s.add_actor(cube_mesh(m[1]), P.matmul(P.translate_matrix((x, 5.2, z)), P.scale_matrix((1.3, 0.2, 0.5))))
    s.lights = P.reference_lights()
    return s
