"""Python face of the library's scene ingest (include/vct_b200.h "scene ingest").

The work is done by the C++ in vct_b200/host/vct_ingest{,_image}.hpp behind the C ABI of libvct_b200.so — there is no
Python re-implementation: OBJ/MTL -> Vertex/index/material arrays as Mesh::loadMesh produces them (reference
src/Graphics/Mesh.cpp:42-206), PNG decode as stb_image (GLHelper.cpp:165-211), DDS/S3TC with the file's mips
(ResourceLoader.h:26-108).  `load_obj` appends one actor to a vct_b200.scene.Scene; `load_image` decodes one file.
"""
import ctypes as C

import numpy as np

from . import lib as L
from . import params as P
from . import scene as S


class Ingest:
    """One loaded OBJ (vct_ingest*).  Arrays are copied out, so the handle can be closed at any time."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @property
    def log(self):
        return self._lib.vct_ingest_log(self._h).decode(errors="replace")

    def mesh(self):
        m = P.IngestMesh()
        if self._lib.vct_ingest_get_mesh(self._h, C.byref(m)):
            raise L.VctError("vct_ingest_get_mesh failed")
        nv, ni = m.n_vertices, m.n_indices
        v = np.ctypeslib.as_array(m.vertices, (nv * 14,)).copy().reshape(nv, 14) if nv else np.zeros((0, 14), np.float32)
        i = np.ctypeslib.as_array(m.indices, (ni,)).copy() if ni else np.zeros(0, np.uint32)
        t = np.ctypeslib.as_array(m.material_of_triangle, (ni // 3,)).copy() if ni else np.zeros(0, np.int32)
        return v, i, t, m

    def materials(self):
        out, n = [], P.IngestMesh()
        self._lib.vct_ingest_get_mesh(self._h, C.byref(n))
        for k in range(n.n_materials):
            m, name = P.Material(), C.c_char_p()
            if self._lib.vct_ingest_get_material(self._h, k, C.byref(m), C.byref(name)):
                raise L.VctError("vct_ingest_get_material failed")
            out.append((name.value.decode(errors="replace"), m))
        return out

    def textures(self):
        out, k = [], 0
        while True:
            t = P.IngestTexture()
            if self._lib.vct_ingest_get_texture(self._h, k, C.byref(t)):
                return out
            px = np.ctypeslib.as_array(C.cast(t.pixels, C.POINTER(C.c_uint8)), (t.bytes,)).copy() if t.bytes else np.zeros(0, np.uint8)
            out.append({"name": t.name.decode(errors="replace"), "width": t.width, "height": t.height, "channels": t.channels,
                        "levels": t.levels, "pixels": px})
            k += 1

    def upload(self, ctx, actor, material_base=0, texture_base=0, model=None):
        m = None if model is None else np.ascontiguousarray(model, np.float32).reshape(16).ctypes.data_as(C.POINTER(C.c_float))
        return self._lib.vct_ingest_upload(ctx, self._h, actor, material_base, texture_base, m)

    def close(self):
        if self._h:
            self._lib.vct_ingest_free(self._h)
            self._h = None

    __del__ = close


def open_obj(path, resource_dir="", decode_textures=True):
    lib = L.load()
    h = C.c_void_p()
    rc = lib.vct_ingest_obj(str(path).encode(), str(resource_dir).encode(), 0 if decode_textures else 1, C.byref(h))
    g = Ingest(h, lib)
    if rc:
        msg = g.log if h else "out of memory"
        g.close()
        raise L.VctError(f"vct_ingest_obj({path}): {msg}")
    return g


def load_image(path, generate_mips=True):
    """-> dict(width, height, channels, levels, pixels) for one PNG / DDS file."""
    lib = L.load()
    h = C.c_void_p()
    rc = lib.vct_ingest_image(str(path).encode(), 1 if generate_mips else 0, C.byref(h))
    g = Ingest(h, lib)
    try:
        if rc:
            raise L.VctError(f"vct_ingest_image({path}): {g.log if h else 'out of memory'}")
        return g.textures()[0]
    finally:
        g.close()


_texture_from_packed = S.texture_from_packed


def load_obj(scene, path, resource_dir="", model=None):
    """Append the OBJ at `path` as a new actor of `scene` (materials and textures are appended, ids rebased)."""
    g = open_obj(path, resource_dir)
    try:
        v, i, t, _ = g.mesh()
        tex_base, mat_base = len(scene.textures), len(scene.materials)
        for tx in g.textures():
            scene.textures.append(_texture_from_packed(tx))
        for _, m in g.materials():
            ids = [x + tex_base if x >= 0 else -1 for x in (m.diffuse_tex, m.specular_tex, m.normal_tex, m.roughness_tex, m.metallic_tex, m.alpha_tex)]
            scene.add_material(*ids, shininess=m.shininess)
        return scene.add_actor(S.Mesh(v, i, t + mat_base), model), g.log
    finally:
        g.close()
