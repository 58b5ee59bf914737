#!/usr/bin/env python3
"""Regenerates the reference-side fixtures of tests/test_ingest.py.  Run where /root/reference exists, after
`make -C oracle` (which compiles the reference's vendored tinyobjloader and stb_image IN PLACE into oracle/_ref/).

  tests/golden/ingest/ref/<name>/{vertices.f32,indices.u32,tri_material.i32,materials.txt}
        output of oracle/_ref/bake_mesh (tinyobj::LoadObj + Mesh::loadMesh restated in tools/bake_mesh.cpp) for the
        hand-written OBJ/MTL files tests/golden/ingest/*.obj (fan triangulation, relative indices, exponents, CRLF,
        texture options, unknown materials, meshes without UVs)
  tests/golden/ingest/reference_objs.json
        sha256 of the same four files for every OBJ the reference ships (8 files)
  tests/golden/ingest/png_ref.json
        "<w> <h> <channels> <fnv1a-64>" as printed by oracle/_ref/stb_dump (= stbi_load(..., STBI_default)) for the
        synthetic PNGs of tests/png_writer.py (every colour type, bit depth, tRNS, Adam7) and for every PNG the
        reference ships
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import png_writer  # noqa: E402

REF = "/root/reference/resources"
BAKE = os.path.join(ROOT, "oracle", "_ref", "bake_mesh")
STB = os.path.join(ROOT, "oracle", "_ref", "stb_dump")
FILES = ("vertices.f32", "indices.u32", "tri_material.i32", "materials.txt")


def main():
    ing = os.path.join(HERE, "ingest")
    for name in ("fan", "nouv"):
        out = os.path.join(ing, "ref", name)
        os.makedirs(out, exist_ok=True)
        subprocess.check_call([BAKE, os.path.join(ing, name + ".obj"), out], stderr=subprocess.DEVNULL)
    digests = {}
    for dirpath, _, files in sorted(os.walk(REF)):
        for f in sorted(files):
            if not f.endswith(".obj"):
                continue
            with tempfile.TemporaryDirectory() as tmp:
                subprocess.check_call([BAKE, os.path.join(dirpath, f), tmp], stderr=subprocess.DEVNULL)
                digests[os.path.relpath(os.path.join(dirpath, f), REF)] = {
                    x: hashlib.sha256(open(os.path.join(tmp, x), "rb").read()).hexdigest() for x in FILES}
    json.dump(digests, open(os.path.join(ing, "reference_objs.json"), "w"), indent=1, sort_keys=True)

    png = {"synthetic": {}, "reference": {}}
    with tempfile.TemporaryDirectory() as tmp:
        names = png_writer.write_all(tmp)
        lines = subprocess.check_output([STB] + [os.path.join(tmp, n) for n in names]).decode().splitlines()
        png["synthetic"] = dict(zip(names, lines))
    paths = sorted(os.path.join(d, f) for d, _, fs in os.walk(REF) for f in fs if f.endswith(".png"))
    lines = subprocess.check_output([STB] + paths).decode().splitlines()
    png["reference"] = {os.path.relpath(p, REF): l for p, l in zip(paths, lines)}
    json.dump(png, open(os.path.join(ing, "png_ref.json"), "w"), indent=1, sort_keys=True)
    print(len(digests), "OBJ digests,", len(png["synthetic"]), "synthetic +", len(png["reference"]), "reference PNG lines")


if __name__ == "__main__":
    main()
