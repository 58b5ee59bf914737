#!/usr/bin/env python3
"""Bake the reference's benchmark scenes into flat arrays under assets/_baked/ (git-ignored, shipped by gpurun).

Harness tooling only: runs in the build container where /root/reference exists; the GPU box
only ever sees the baked output.  Geometry goes through tools/bake_mesh.cpp (the reference's own vendored
tinyobjloader, compiled in place); textures are copied as the PNG files the reference's loader would bind
(reference src/Graphics/Mesh.cpp:81-86 — diffuse/specular/normal/roughness/metallic/alpha only).
A texture file that does not exist is treated as "no map" (the reference's behaviour there is undefined:
src/Graphics/GLHelper.cpp:171-174 logs and carries on with uninitialised sizes).
"""
import os, shutil, subprocess, sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "assets", "_baked")
SCENES = {
    "cube": "resources/cube.obj",
    "bunny": "resources/bunny.obj",
    "sponza_pbr": "resources/sponza/sponza_pbr.obj",
    "nanosuit": "resources/nanosuit/nanosuit.obj",
}

def main():
    if not os.path.isdir(REF):
        sys.exit("bake_assets: /root/reference not present (run in the build container)")
    os.makedirs(OUT, exist_ok=True)
    tool = os.path.join(OUT, "bake_mesh")
    subprocess.check_call(["g++", "-O2", "-std=c++17", f"-I{REF}/ext/include",
                           os.path.join(ROOT, "tools", "bake_mesh.cpp"), "-o", tool])
    for name, rel in SCENES.items():
        dst = os.path.join(OUT, name)
        os.makedirs(os.path.join(dst, "textures"), exist_ok=True)
        subprocess.check_call([tool, os.path.join(REF, rel), dst])
        base = os.path.dirname(os.path.join(REF, rel))
        lines = []
        for line in open(os.path.join(dst, "materials.txt")):
            f = line.rstrip("\n").split("|")[:7]
            f += [""] * (7 - len(f))
            for i in range(1, 7):
                if not f[i]:
                    continue
                src = (os.path.join(REF, "resources", f[i][1:]) if f[i].startswith("@")
                       else os.path.join(base, f[i].replace("\\", "/")))
                if not os.path.isfile(src):
                    print(f"  {name}: missing texture {f[i]} -> no map")
                    f[i] = ""
                    continue
                flat = os.path.basename(src)
                shutil.copyfile(src, os.path.join(dst, "textures", flat))
                f[i] = flat
            lines.append("|".join(f))
        open(os.path.join(dst, "materials.txt"), "w").write("\n".join(lines) + "\n")
    os.remove(tool)
    for name in SCENES:
        decode_textures(name)
    print("baked into", OUT)


def decode_textures(name):
    """Decode every texture of a baked scene ONCE, here in the build container, into textures.npz (all mip levels packed, as
    vct_upload_texture takes them).  The harness (bench.py, tests, both bench arms) then loads texels with numpy alone: the
    `--impl reference` process never loads libvct_b200.so ("the reference arm loads the repo's .so").  The decoder is
    the library's own PNG reader, texel-identical to the reference's stb_image on every shipped file (tests/test_ingest.py)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from vct_b200 import ingest
    d = os.path.join(OUT, name, "textures")
    out = {}
    for fn in sorted(os.listdir(d)):
        t = ingest.load_image(os.path.join(d, fn))
        out[fn + "|meta"] = np.array([t["width"], t["height"], t["channels"], t["levels"]], np.int32)
        out[fn + "|px"] = t["pixels"]
    np.savez_compressed(os.path.join(OUT, name, "textures.npz"), **out)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--decode-only":
        for n in SCENES:
            decode_textures(n)
    else:
        main()
