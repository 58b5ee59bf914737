#!/usr/bin/env python3
"""A small whole frame (room scene, 32^3, 128x96, 256^2 shadow map) in one voxeliser mode, for compute-sanitizer:
usage: sanitize_frame.py det|cas|max|tess|warp [frames]   (prints a checksum of the volumes and the image)"""
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
p = S.room_params(W, H)
if mode == "cas": p.deterministic = 0
if mode == "max": p.voxelize_atomic_max = 1
if mode == "tess": p.voxelize_tesselation = 1; p.voxelize_atomic_max = 1
if mode == "warp": p.warp_texture = 1; p.temporal_filter_radiance = 1
g = Pipeline(sc, D, L, SS, W, H)
for _ in range(frames):
    g.frame(p)
g.sync()
crc = zlib.crc32(g.read_volume(P.VOL_COLOR).tobytes()) ^ zlib.crc32(g.read_volume(P.VOL_RADIANCE).tobytes()) ^ zlib.crc32(g.read_image().tobytes())
i = g.counters()
print(f"sanitize_frame {mode}: {frames} frames, fragments {i.total_fragments}, voxels {i.unique_voxels}, crc {crc:08x}")
g.close()
