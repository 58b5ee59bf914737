#!/usr/bin/env python3
"""Hardware smoke of a few frame configurations through vct_headless (the C++ host: no Python start-up on the box, a whole
run takes ~8 s of GPU-box time) against the CPU oracle.

  python tools/hw_smoke.py prepare     here: packs the room scene into gpurun_in/room.vcts, renders the oracle's frames into gpurun_in/
  gpurun --timeout 25 -- 'bash tools/hw_smoke.sh'        on the box: one vct_headless run per configuration -> gpurun_out/hw_<name>.{ppm,json,err}
  python tools/hw_smoke.py compare     here: PSNR / differing pixels / counters per configuration

The two hosts build their matrices with different float libraries (numpy fp32 here, C++ on the box), so the comparison is the
final image's gate (PSNR >= 45 dB) plus the VoxelizeInfo counters, not bit equality."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
D, L, SS, W, H = 64, 6, 1024, 256, 256
CONFIGS = {   # name -> (frame parameters, vct_headless flags)
    "plain": ({}, ""),
    "tess_warp": ({"voxelize_tesselation": 1, "voxelize_atomic_max": 1, "voxelize_tesselation_warp": 1}, "--tesselation --atomic-max --tesselation-warp"),
    "raster_tess_warp": ({"voxelize_tesselation_warp": 1}, "--tesselation-warp"),
    "view_tess_warp": ({"voxelize_tesselation": 1, "voxelize_atomic_max": 1, "voxelize_tesselation_warp": 1, "debug_view": 1, "miplevel": 0.0},
                       "--tesselation --atomic-max --tesselation-warp --view voxels --miplevel 0"),
    "both_warps": ({"warp_texture": 1, "warp_voxels": 1}, "--warp-texture --warp-voxels"),
}


def prepare():
    from tests.oracle_lib import Oracle
    from tools import pack_scene
    from vct_b200 import params as P, scene as S
    os.makedirs(os.path.join(ROOT, "gpurun_in"), exist_ok=True)
    sc = S.room_scene()
    pack_scene.pack(sc, os.path.join(ROOT, "gpurun_in", "room.vcts"))
    cam = P.Camera(position=(1.1, 0.3, 1.2), front=(-0.65, -0.25, -0.72))
    for name, (kw, _) in CONFIGS.items():
        p = P.default_params(W, H, cam, sc.lights[0], voxel_min=-1.5, voxel_max=1.5)
        for k, v in kw.items():
            setattr(p, k, v)
        o = Oracle(sc, D, L, SS, W, H)
        o.frame(p)
        np.save(os.path.join(ROOT, "gpurun_in", f"oracle_{name}.npy"), o.image.view(np.uint8).reshape(H, W, 4)[::-1, :, :3])   # PPM rows run top-down
        np.save(os.path.join(ROOT, "gpurun_in", f"oracle_{name}_unique.npy"), np.array([int(((o.color[0] >> 24) != 0).sum())]))
        print("oracle frame:", name)


def compare():
    ok = True
    for name in CONFIGS:
        ppm = open(os.path.join(ROOT, "gpurun_out", f"hw_{name}.ppm"), "rb").read()
        head = f"P6\n{W} {H}\n255\n".encode()
        got = np.frombuffer(ppm[len(head):], np.uint8).reshape(H, W, 3).astype(np.float64)
        ref = np.load(os.path.join(ROOT, "gpurun_in", f"oracle_{name}.npy")).astype(np.float64)
        uniq = int(np.load(os.path.join(ROOT, "gpurun_in", f"oracle_{name}_unique.npy"))[0])
        mse = ((got - ref) ** 2).mean()
        psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
        rep = json.loads(open(os.path.join(ROOT, "gpurun_out", f"hw_{name}.json")).read().strip().splitlines()[-1])
        good = psnr >= 45.0 and rep["ok"] and rep["unique_voxels"] == uniq
        ok &= good
        print(f"{name:18s} PSNR {psnr:6.2f} dB  max|d| {np.abs(got - ref).max():3.0f}  pixels differing {(np.abs(got - ref).max(axis=2) > 0).mean():.4f}  "
              f"unique voxels B200 {rep['unique_voxels']} oracle {uniq}  fragments {rep['total_fragments']}  {'ok' if good else 'FAIL'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(prepare() if sys.argv[1:] == ["prepare"] else compare() if sys.argv[1:] == ["compare"] else print(__doc__))
