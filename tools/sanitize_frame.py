#!/usr/bin/env python3
"""A small whole frame (room scene, 32^3, 128x96, 256^2 shadow map) in one voxeliser mode, for compute-sanitizer:
usage: sanitize_frame.py det|cas|max|tess|warp|long|huge|cubes [frames]   (prints a checksum of the volumes and the image)
long / huge: 300 / 1300 coincident quads on a 32^3 grid -> per-voxel lists for k_voxel_resolve_medium / the k_voxel_huge_* kernels."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vct_b200 import params as P
from vct_b200 import scene as S
from vct_b200.pipeline import Pipeline

mode = sys.argv[1] if len(sys.argv) > 1 else "det"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D, L, SS, W, H = 32, 5, 256, 128, 96
sc = S.room_scene()
if mode in ("long", "huge"):
    sc = S.Scene()
    mat = sc.add_material(diffuse=sc.add_texture(S.checker_texture(16, 4, a=(200, 60, 40), b=(40, 90, 220), seed=1)))
    n = 300 if mode == "long" else 1300
    parts = [S.quad_mesh([(-0.6, -0.2, 0.5), (0.7, -0.2, 0.5), (0.7, -0.2, -0.6), (-0.6, -0.2, -0.6)], (0, 1, 0), mat, 1.0 + 0.1 * (k % 7)) for k in range(n)]
    parts.append(S.quad_mesh([(-1.4, -1.0, 1.4), (1.4, -1.0, 1.4), (1.4, -1.0, -1.4), (-1.4, -1.0, -1.4)], (0, 1, 0), mat, 4.0))
    sc.add_actor(S.merge_meshes(parts))
    sc.lights = [P.make_light(position=(1.2, 4.0, 0.7), direction=(-0.28, -0.9, -0.2), shadow_caster=True, type_=1)]
    W, H = 96, 64
p = S.room_params(W, H)
if mode == "cas": p.deterministic = 0
if mode == "max": p.voxelize_atomic_max = 1
if mode == "tess": p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1
if mode == "warp": p.warp_texture = 1; p.temporal_filter_radiance = 1
g = Pipeline(sc, D, L, SS, W, H)
for k in range(frames):
    try:
        g.frame(p)
    except Exception as e:                     # huge: the first frame meets a > 1024 list without the long-list kernels and says so once
        if mode != "huge" or k != 1:
            raise
        print("expected once:", str(e)[:90])
        g.frame(p)
if mode == "cubes":                             # Application::debugVoxels after the frames (vct_debug_voxels)
    g.debug_voxels(p)
g.sync()
crc = zlib.crc32(g.read_volume(P.VOL_COLOR).tobytes()) ^ zlib.crc32(g.read_volume(P.VOL_RADIANCE).tobytes()) ^ zlib.crc32(g.read_image().tobytes())
i = g.counters()
print(f"sanitize_frame {mode}: {frames} frames, fragments {i.total_fragments}, voxels {i.unique_voxels}, crc {crc:08x}")
g.close()
