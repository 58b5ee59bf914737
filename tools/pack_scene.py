#!/usr/bin/env python3
"""Write a vct_b200.scene.Scene as the flat "VCTS" file the C++ host (vct_b200/host/vct_host.hpp, vct_headless) loads.

Layout (little-endian):  u32 magic "VCTS", u32 version 1, u32 n_textures, n_materials, n_actors, n_lights
  per texture : u32 width, height, channels, levels, nbytes ; nbytes of pixels (all mip levels, level 0 first)
  per material: vct_material (40 bytes: 6 x i32 texture ids, f32 shininess, 3 x f32 diffuse)
  per actor   : u32 n_vertices, n_triangles ; f32[n_vertices*14] ; u32[n_triangles*3] ; i32[n_triangles] ; f32[16] model
  per light   : vct_light (80 bytes)
usage: pack_scene.py <room|1|2|3|4|5> <out.vcts>"""
import ctypes as C
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from vct_b200 import scene as S  # noqa: E402


def pack(scene, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<6I", 0x53544356, 1, len(scene.textures), len(scene.materials), len(scene.meshes), len(scene.lights)))
        for t in scene.textures:
            px = np.ascontiguousarray(t.packed(), np.uint8)
            f.write(struct.pack("<5I", t.width, t.height, t.channels, min(16, len(t.levels)), px.nbytes)); f.write(px.tobytes())
        for m in scene.materials:
            f.write(bytes(m))
        for mesh, model in zip(scene.meshes, scene.models):
            v = np.ascontiguousarray(mesh.vertices, np.float32); i = np.ascontiguousarray(mesh.indices, np.uint32)
            f.write(struct.pack("<2I", len(v), i.size // 3)); f.write(v.tobytes()); f.write(i.tobytes())
            f.write(np.ascontiguousarray(mesh.tri_material, np.int32).tobytes())
            f.write(np.ascontiguousarray(np.asarray(model, np.float32).reshape(16)).tobytes())
        for l in scene.lights:
            f.write(bytes(l))


if __name__ == "__main__":
    which, out = sys.argv[1], sys.argv[2]
    sc = S.room_scene() if which == "room" else S.config_scene(int(which))[0]
    pack(sc, out)
    print(out, os.path.getsize(out), "bytes")
