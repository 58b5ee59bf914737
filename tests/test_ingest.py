"""Scene ingest: the library's OBJ/MTL, PNG and DDS readers against the reference's own loaders.

Reference side = the reference's vendored tinyobjloader 1.0.7 and stb_image 2.15 compiled IN PLACE into oracle/_ref/
(oracle/Makefile: bake_mesh, stb_dump).  Their outputs for the hand-written fixtures and their digests for every asset
the reference ships are committed under tests/golden/ingest/ (generator: tests/golden/make_ingest_ref.py), so the
comparisons also run where /root/reference does not exist; where it does, the binaries are run live as well.
Everything here is host code: no GPU.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests import png_writer
from vct_b200 import ingest
from vct_b200 import lib as L
from vct_b200 import params as P
from vct_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ingest")
DUMP = os.path.join(ROOT, "vct_b200", "lib", "vct_ingest_dump")
BAKE = os.path.join(ROOT, "oracle", "_ref", "bake_mesh")
STB = os.path.join(ROOT, "oracle", "_ref", "stb_dump")
REF = "/root/reference/resources"
FILES = ("vertices.f32", "indices.u32", "tri_material.i32", "materials.txt")
have_ref = os.path.isdir(REF) and os.path.isfile(BAKE) and os.path.isfile(STB)
needs_ref = pytest.mark.skipif(not have_ref, reason="reference checkout / oracle/_ref binaries not present")


def fnv1a(data):
    h = 1469598103934665603
    for b in bytes(data):                                      # small inputs only
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def dump_obj(path, out):
    subprocess.check_call([DUMP, path, str(out)], stderr=subprocess.DEVNULL)
    return {f: open(os.path.join(out, f), "rb").read() for f in FILES}


# ------------------------------------------------------------------------------------------------- OBJ / MTL
@pytest.mark.parametrize("name", ["fan", "nouv"])
def test_obj_fixture_equals_reference_loader_output(name, tmp_path):
    """tinyobj::LoadObj + Mesh::loadMesh (reference Mesh.cpp:42-206) vs vct::load_obj, byte for byte."""
    ours = dump_obj(os.path.join(GOLD, name + ".obj"), tmp_path)
    for f in FILES:
        assert ours[f] == open(os.path.join(GOLD, "ref", name, f), "rb").read(), f"{name}/{f} differs from the reference loader"
    if have_ref:
        live = tmp_path / "live"; live.mkdir()
        subprocess.check_call([BAKE, os.path.join(GOLD, name + ".obj"), str(live)], stderr=subprocess.DEVNULL)
        for f in FILES:
            assert ours[f] == open(live / f, "rb").read()


def test_obj_fixture_semantics_through_the_abi():
    g = ingest.open_obj(os.path.join(GOLD, "fan.obj"), GOLD, decode_textures=False)
    v, i, t, m = g.mesh()
    ref = os.path.join(GOLD, "ref", "fan")
    assert np.array_equal(v.view(np.uint32), np.fromfile(os.path.join(ref, "vertices.f32"), np.uint32).reshape(-1, 14))
    assert np.array_equal(i, np.fromfile(os.path.join(ref, "indices.u32"), np.uint32))
    assert np.array_equal(t, np.fromfile(os.path.join(ref, "tri_material.i32"), np.int32))
    assert "missing.mtl" in g.log                                         # first mtllib entry is absent: a warning, not an error
    mats = g.materials()
    assert [n for n, _ in mats] == ["brick", "glass", "default"]
    names = [x["name"] for x in g.textures()]
    brick, glass, default = (m for _, m in mats)
    assert [names[k] for k in (brick.diffuse_tex, brick.specular_tex, brick.normal_tex, brick.roughness_tex, brick.metallic_tex)] == \
        ["brick_d.png", "brick_s.png", "brick_n.png", "brick_r.png", "brick_m.png"] and brick.alpha_tex == -1
    assert names[glass.diffuse_tex] == "glass.png" and names[glass.alpha_tex] == "glass_a.png" and glass.normal_tex == -1
    assert names[default.diffuse_tex] == "@default_texture.png"
    assert all(abs(m.shininess - 32.0) < 1e-9 and tuple(m.diffuse) == (0.0, 0.0, 0.0) for _, m in mats)   # Mesh.h:33-43 self-assignment
    # quad + pentagon + 2 triangles + relative-index triangle = 2 + 3 + 2 + 1 fan triangles; draw order is by material
    assert i.size == 3 * 8 and list(t) == sorted(t)
    assert m.n_materials == 3 and tuple(m.bounds_min) == (-25.0, 0.0, -0.125) and tuple(m.bounds_max) == (1.5, 1.0, 3.0) and m.radius == 13.25
    g.close()
    # no UVs -> NaN tangent frame, the reference's behaviour (Mesh.cpp:178,196)
    g = ingest.open_obj(os.path.join(GOLD, "nouv.obj"), GOLD, decode_textures=False)
    v, i, t, _ = g.mesh()
    assert np.isnan(v[:, 8:14]).all() and (t == 0).all() and [n for n, _ in g.materials()] == ["default"]
    g.close()


def test_obj_errors_are_reported_not_thrown(tmp_path):
    with pytest.raises(L.VctError, match="cannot open"):
        ingest.open_obj(str(tmp_path / "absent.obj"))
    lib = L.load()
    assert lib.vct_ingest_obj(None, None, 0, None) == 1
    h = C.c_void_p()
    assert lib.vct_ingest_obj(b"/nonexistent.obj", None, 0, C.byref(h)) == 1 and h.value   # handle carries the log
    assert b"cannot open" in lib.vct_ingest_log(h)
    lib.vct_ingest_free(h)
    lib.vct_ingest_free(None)
    assert lib.vct_ingest_get_mesh(None, None) == 1 and lib.vct_ingest_upload(None, None, 0, 0, 0, None) == 1
    # garbage in: lines that are not OBJ are ignored, faces with out-of-range corners must not crash the process
    p = tmp_path / "junk.obj"
    p.write_bytes(b"\x00\x01garbage\nv 1 2\nv a b c\nf\nf 1\nusemtl\nmtllib\nvt\n" + bytes(range(256)))
    g = ingest.open_obj(str(p), decode_textures=False)
    assert g.mesh()[1].size == 0
    g.close()


@needs_ref
def test_every_reference_obj_equals_the_reference_loader(tmp_path):
    """All 8 OBJ files the reference ships (cube … three Sponza variants, nanosuit): arrays and material tables identical
    to tinyobjloader + Mesh::loadMesh, checked against the committed digests and a live run of oracle/_ref/bake_mesh."""
    digests = json.load(open(os.path.join(GOLD, "reference_objs.json")))
    assert len(digests) == 8
    for rel, want in sorted(digests.items()):
        out = tmp_path / rel.replace("/", "_"); out.mkdir()
        ours = dump_obj(os.path.join(REF, rel), out)
        got = {f: hashlib.sha256(ours[f]).hexdigest() for f in FILES}
        assert got == want, f"{rel}: ingest differs from the reference loader"
    live = tmp_path / "live"; live.mkdir()
    subprocess.check_call([BAKE, os.path.join(REF, "nanosuit/nanosuit.obj"), str(live)], stderr=subprocess.DEVNULL)
    assert {f: hashlib.sha256(open(live / f, "rb").read()).hexdigest() for f in FILES} == digests["nanosuit/nanosuit.obj"]


# ------------------------------------------------------------------------------------------------------- PNG
def test_png_decoder_equals_stb_image_on_every_colour_type_and_depth(tmp_path):
    """32 synthetic PNGs (grey/RGB/palette/grey+alpha/RGBA x 1-16 bit, tRNS, Adam7, stored/fixed/dynamic deflate, split
    IDAT, all five row filters) decoded by vct::decode_png vs stbi_load(STBI_default) of the reference's stb_image."""
    want = json.load(open(os.path.join(GOLD, "png_ref.json")))["synthetic"]
    names = png_writer.write_all(str(tmp_path))
    assert sorted(names) == sorted(want)
    ours = subprocess.check_output([DUMP, "--images"] + [str(tmp_path / n) for n in names]).decode().splitlines()
    for n, line in zip(names, ours):
        assert line == want[n], f"{n}: ours {line} != stb_image {want[n]}"
    if have_ref:
        live = subprocess.check_output([STB] + [str(tmp_path / n) for n in names]).decode().splitlines()
        assert live == ours
    # through the C ABI, pixel for pixel, against the source samples of one RGBA8 file
    rng = np.random.default_rng(7)
    px = rng.integers(0, 256, (19, 23, 4)).astype(np.uint16)
    (tmp_path / "abi.png").write_bytes(png_writer.encode(px, 6, 8, rng, interlace=True))
    t = ingest.load_image(tmp_path / "abi.png", generate_mips=False)
    assert (t["width"], t["height"], t["channels"], t["levels"]) == (23, 19, 4, 1)
    assert np.array_equal(t["pixels"].reshape(19, 23, 4), px.astype(np.uint8))


@needs_ref
def test_every_reference_png_equals_stb_image():
    want = json.load(open(os.path.join(GOLD, "png_ref.json")))["reference"]
    assert len(want) == 152
    rels = sorted(want)
    ours = subprocess.check_output([DUMP, "--images"] + [os.path.join(REF, r) for r in rels]).decode().splitlines()
    bad = [r for r, line in zip(rels, ours) if line != want[r]]
    assert not bad, f"decode differs from stb_image for {bad[:5]}"


def test_png_damage_is_an_error_never_a_crash(tmp_path):
    rng = np.random.default_rng(11)
    good = png_writer.encode(rng.integers(0, 256, (40, 40, 3)).astype(np.uint16), 2, 8, rng, level=6)
    lib = L.load()

    def status(data):
        p = tmp_path / "x.png"; p.write_bytes(data)
        h = C.c_void_p()
        rc = lib.vct_ingest_image(str(p).encode(), 1, C.byref(h))
        log = lib.vct_ingest_log(h).decode()
        lib.vct_ingest_free(h)
        return rc, log

    assert status(good)[0] == 0
    assert status(b"")[0] == 1 and status(good[:20])[0] == 1 and "signature" in status(b"JFIF" * 8)[1]
    for cut in (60, len(good) // 2, len(good) - 20):                     # truncated inside IDAT / before IEND
        assert status(good[:cut])[0] == 1
    for k in range(200):                                                 # random byte damage: any status, no crash, no hang
        d = bytearray(good)
        for _ in range(1 + k % 4):
            d[int(rng.integers(8, len(d)))] = int(rng.integers(0, 256))
        status(bytes(d))
    assert status(good)[0] == 0
    h = C.c_void_p()
    assert lib.vct_ingest_image(b"/nonexistent.png", 1, C.byref(h)) == 1 and b"cannot open" in lib.vct_ingest_log(h)
    lib.vct_ingest_free(h)


def test_generated_mips_equal_the_harness_box_filter(tmp_path):
    """build_mips (C++) == scene.build_mips (numpy): 2x2 box, round half up, odd sizes, 1-pixel-wide tails."""
    rng = np.random.default_rng(3)
    for (h, w, ch, color) in ((37, 64, 3, 2), (1, 9, 1, 0), (16, 1, 4, 6), (5, 5, 4, 6), (128, 128, 1, 0)):
        px = rng.integers(0, 256, (h, w, ch)).astype(np.uint16)
        (tmp_path / "m.png").write_bytes(png_writer.encode(px, color, 8, rng))
        t = ingest.load_image(tmp_path / "m.png")
        want = S.Texture(px.astype(np.uint8))
        assert t["levels"] == len(want.levels) and np.array_equal(t["pixels"], want.packed()), (h, w, ch)


# ------------------------------------------------------------------------------------------------------- DDS
def _expand565(c):
    r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
    return np.array([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], np.int64)


def _decode_block_numpy(block, kind):
    """Independent restatement of EXT_texture_compression_s3tc block decode (truncating thirds, as Mesa evaluates it)."""
    col = block[8:] if kind != 1 else block
    c0, c1 = int(col[0]) | int(col[1]) << 8, int(col[2]) | int(col[3]) << 8
    p = [_expand565(c0), _expand565(c1)]
    if kind != 1 or c0 > c1:
        p += [(2 * p[0] + p[1]) // 3, (p[0] + 2 * p[1]) // 3]
    else:
        p += [(p[0] + p[1]) // 2, np.zeros(3, np.int64)]
    bits = int.from_bytes(bytes(col[4:8]), "little")
    out = np.zeros((16, 4), np.uint8)
    for i in range(16):
        out[i, :3] = p[(bits >> (2 * i)) & 3]
        out[i, 3] = 255
    if kind == 3:
        for i in range(16):
            out[i, 3] = ((int(block[i >> 1]) >> ((i & 1) * 4)) & 15) * 17
    elif kind == 5:
        a0, a1 = int(block[0]), int(block[1])
        a = [a0, a1] + ([((7 - k) * a0 + k * a1) // 7 for k in range(1, 7)] if a0 > a1 else [((5 - k) * a0 + k * a1) // 5 for k in range(1, 5)] + [0, 255])
        bits = int.from_bytes(bytes(block[2:8]), "little")
        for i in range(16):
            out[i, 3] = a[(bits >> (3 * i)) & 7]
    return out.reshape(4, 4, 4)


def _dds(width, height, fourcc, mips, payload):
    hdr = np.zeros(31, np.uint32)
    hdr[0], hdr[2], hdr[3], hdr[6] = 124, height, width, mips
    hdr[18] = 32; hdr[20] = int.from_bytes(fourcc, "little")
    return b"DDS " + hdr.tobytes() + payload


@pytest.mark.parametrize("fourcc,kind", [(b"DXT1", 1), (b"DXT3", 3), (b"DXT5", 5)])
def test_dds_block_decode_and_file_mips(fourcc, kind, tmp_path):
    rng = np.random.default_rng(kind)
    w, h, block = 20, 12, 8 if kind == 1 else 16                          # not a multiple of 4 at level >= 1
    sizes, payload = [], b""
    for l in range(5):
        lw, lh = max(1, w >> l), max(1, h >> l)
        sizes.append((lw, lh)); payload += rng.integers(0, 256, ((lw + 3) // 4) * ((lh + 3) // 4) * block, dtype=np.uint8).tobytes()
    (tmp_path / "t.dds").write_bytes(_dds(w, h, fourcc, 5, payload))
    t = ingest.load_image(tmp_path / "t.dds")
    ch = 3 if kind == 1 else 4
    assert (t["width"], t["height"], t["channels"], t["levels"]) == (w, h, ch, 5)     # DXT1 = GL_COMPRESSED_RGB_S3TC_DXT1_EXT
    raw, off, at = np.frombuffer(payload, np.uint8), 0, 0
    for lw, lh in sizes:
        got = t["pixels"][at:at + lw * lh * ch].reshape(lh, lw, ch); at += lw * lh * ch
        for by in range((lh + 3) // 4):
            for bx in range((lw + 3) // 4):
                ref = _decode_block_numpy(raw[off:off + block], kind); off += block
                tile = got[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4]
                assert np.array_equal(tile, ref[:tile.shape[0], :tile.shape[1], :ch])
    assert at == t["pixels"].size
    with pytest.raises(L.VctError, match="DXT"):
        (tmp_path / "u.dds").write_bytes(_dds(4, 4, b"ATI2", 1, bytes(16)))
        ingest.load_image(tmp_path / "u.dds")
    with pytest.raises(L.VctError, match="truncated"):
        (tmp_path / "v.dds").write_bytes(_dds(64, 64, fourcc, 1, bytes(100)))
        ingest.load_image(tmp_path / "v.dds")


@needs_ref
def test_reference_dds_textures_decode_like_an_independent_decoder():
    """Level 0 of every DDS the reference ships (75 files, DXT1 and DXT5) vs Pillow's S3TC decoder: identical."""
    from PIL import Image
    paths = sorted(os.path.join(d, f) for d, _, fs in os.walk(REF) for f in fs if f.endswith(".dds"))
    assert len(paths) == 75
    for p in paths[::5]:                                                  # every fifth file keeps the test short
        t = ingest.load_image(p)
        ours = t["pixels"][:t["width"] * t["height"] * t["channels"]].reshape(t["height"], t["width"], t["channels"])
        ref = np.asarray(Image.open(p).convert("RGBA"))
        assert np.array_equal(ours, ref[..., :t["channels"]]), p
        assert t["levels"] == int.from_bytes(open(p, "rb").read(32)[28:32], "little")


# ------------------------------------------------------------------------------------- whole scene through the ABI
def test_scene_ingest_binds_textures_like_mesh_draw(tmp_path):
    """An OBJ with PNG textures next to it: texture de-duplication by MTL name, '\\\\' -> '/', missing file -> no map + log,
    two-channel image -> no map + log, default material from <resource_dir>/default_texture.png (Mesh.cpp:57-107)."""
    rng = np.random.default_rng(5)
    (tmp_path / "tex").mkdir()
    rgb = rng.integers(0, 256, (8, 8, 3)).astype(np.uint16)
    (tmp_path / "tex" / "a.png").write_bytes(png_writer.encode(rgb, 2, 8, rng))
    (tmp_path / "ga.png").write_bytes(png_writer.encode(rng.integers(0, 256, (4, 4, 2)).astype(np.uint16), 4, 8, rng))
    (tmp_path / "default_texture.png").write_bytes(png_writer.encode(np.full((2, 2, 3), 255, np.uint16), 2, 8, rng))
    (tmp_path / "s.mtl").write_text("newmtl one\nmap_Kd tex\\a.png\nmap_Ks tex\\a.png\nmap_d gone.png\nnorm ga.png\nnewmtl two\nmap_Kd tex\\a.png\n")
    (tmp_path / "s.obj").write_text("mtllib s.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf 1 2 3\nusemtl one\nf 1/1 2/2 3/3\nusemtl two\nf 3/3 2/2 1/1\n")
    sc = S.Scene()
    actor, log = ingest.load_obj(sc, tmp_path / "s.obj", tmp_path)
    assert actor == 0 and len(sc.textures) == 2 and len(sc.materials) == 3
    one, two, default = sc.materials
    assert one.diffuse_tex == one.specular_tex == two.diffuse_tex == 0 and one.alpha_tex == -1 and one.normal_tex == -1
    assert default.diffuse_tex == 1 and "gone.png" in log and "ga.png" in log
    assert np.array_equal(sc.textures[0].levels[0], rgb.astype(np.uint8)) and len(sc.textures[0].levels) == 4
    assert list(sc.meshes[0].tri_material) == [0, 1, 2]
    assert P.Material is type(one)


def _write_box_scene(d, rng):
    """A small lit box with two textured materials + untextured faces, as OBJ/MTL/PNG files (for the ABI and GPU tests)."""
    os.makedirs(os.path.join(d, "tex"), exist_ok=True)
    tex_a = S.checker_texture(32, 4, seed=1)
    tex_b = S.checker_texture(16, 2, a=(40, 200, 60), b=(220, 220, 40), seed=2)
    open(os.path.join(d, "tex", "a.png"), "wb").write(png_writer.encode(tex_a.astype(np.uint16), 2, 8, rng))
    open(os.path.join(d, "tex", "b.png"), "wb").write(png_writer.encode(tex_b.astype(np.uint16), 2, 8, rng, interlace=True))
    open(os.path.join(d, "default_texture.png"), "wb").write(png_writer.encode(np.full((4, 4, 3), 200, np.uint16), 2, 8, rng))
    open(os.path.join(d, "box.mtl"), "w").write("newmtl floor\nmap_Kd tex\\a.png\nnewmtl wall\nmap_Kd tex/b.png\n")
    v = ["v -4 0 -4", "v 4 0 -4", "v 4 0 4", "v -4 0 4", "v -4 5 -4", "v 4 5 -4", "v 4 5 4", "v -4 5 4",
         "v -1 0 -1", "v 1 0 -1", "v 1 2 -1", "v -1 2 -1", "v -1 0 1", "v 1 0 1", "v 1 2 1", "v -1 2 1"]
    vt = ["vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1"]
    vn = ["vn 0 1 0", "vn 0 0 1", "vn 1 0 0", "vn -1 0 0", "vn 0 0 -1"]
    f = ["usemtl floor", "f 1/1/1 4/2/1 3/3/1 2/4/1",
         "usemtl wall", "f 1/1/2 2/2/2 6/3/2 5/4/2", "f 1/1/3 5/2/3 8/3/3 4/4/3",
         "usemtl none", "f 9/1/5 12/2/5 11/3/5 10/4/5", "f 13/1/2 14/2/2 15/3/2 16/4/2", "f 12/1/1 16/2/1 15/3/1 11/4/1",
         "f 10/1/3 11/2/3 15/3/3 14/4/3", "f 9/1/4 13/2/4 16/3/4 12/4/4"]
    open(os.path.join(d, "box.obj"), "w").write("mtllib box.mtl\n" + "\n".join(v + vt + vn + f) + "\n")
    return os.path.join(d, "box.obj")


def test_headless_host_loads_an_obj_through_the_ingest(tmp_path):
    """vct_headless mesh.obj: Scene::addObj (C++ host) parses the scene, then vct_create fails loudly without a GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    obj = _write_box_scene(str(tmp_path), np.random.default_rng(1))
    exe = os.path.join(ROOT, "vct_b200", "lib", "vct_headless")
    r = subprocess.run([exe, obj, "--dim", "32", "--size", "64x64", "--shadow", "128"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and "Failed to load mesh" not in r.stderr
    r = subprocess.run([exe, str(tmp_path / "absent.obj")], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to load mesh" in r.stderr


@needs_ref
@pytest.mark.parametrize("name,rel", [("cube", "cube.obj"), ("sponza_pbr", "sponza/sponza_pbr.obj")])
def test_ingested_scene_equals_the_baked_scene_the_benchmark_reads(name, rel):
    """The scene bench.py reads from assets/_baked (arrays written by the reference's tinyobjloader, textures decoded by the
    library) and the scene the library's own ingest builds from the reference's files are the same inputs: arrays identical,
    and for every material slot the same texels in every mip level."""
    if not S.baked_available(name):
        pytest.skip("assets/_baked missing")
    a, b = S.Scene(), S.Scene()
    S.load_baked(a, name)
    ingest.load_obj(b, os.path.join(REF, rel), REF)
    ma, mb = a.meshes[0], b.meshes[0]
    assert np.array_equal(ma.vertices.view(np.uint32), mb.vertices.view(np.uint32)) and np.array_equal(ma.indices, mb.indices)
    assert np.array_equal(ma.tri_material, mb.tri_material) and len(a.materials) == len(b.materials)
    checked = 0
    for x, y in zip(a.materials, b.materials):
        for slot in ("diffuse_tex", "specular_tex", "normal_tex", "roughness_tex", "metallic_tex", "alpha_tex"):
            i, j = getattr(x, slot), getattr(y, slot)
            assert (i < 0) == (j < 0), slot
            if i < 0:
                continue
            ta, tb = a.textures[i], b.textures[j]
            assert (ta.width, ta.height, len(ta.levels)) == (tb.width, tb.height, len(tb.levels))
            assert ta.channels == tb.channels
            for la, lb in zip(ta.levels, tb.levels):
                assert np.array_equal(la, lb)
            checked += 1
    assert checked >= 1


@pytest.mark.gpu
def test_vct_ingest_upload_renders_like_the_oracle(tmp_path):
    """OBJ + MTL + PNG files -> vct_ingest_obj -> vct_ingest_upload -> one frame: volumes bit-exact and image >= 45 dB against
    the oracle fed with the same scene (and identical to the frame of a context that received the scene array by array)."""
    from tests.oracle_lib import Oracle
    from vct_b200.pipeline import Pipeline
    obj = _write_box_scene(str(tmp_path), np.random.default_rng(1))
    D, Lv, SS, W, H = 64, 5, 512, 320, 240
    sc = S.Scene()
    ingest.load_obj(sc, obj, str(tmp_path))
    sc.lights = P.reference_lights(); sc.lights[0].position = P.F3(3.0, 10.0, 2.0); sc.lights[0].direction = P.F3(-0.3, -0.9, -0.25)
    sc.lights[1].position = P.F3(0.0, 3.0, 2.5)
    cam = P.Camera(position=(3.2, 2.6, 3.4), front=(-3.2, -1.6, -3.4))
    p = P.default_params(W, H, cam, sc.lights[0], voxel_min=-5.0, voxel_max=5.0, voxel_center=(0, 2.5, 0))
    ref = Pipeline(sc, D, Lv, SS, W, H)
    ref.frame(p)
    g = Pipeline(None, D, Lv, SS, W, H)
    h = ingest.open_obj(obj, str(tmp_path))
    assert h.upload(g.h, actor=0) == 0
    h.close()
    g.set_lights(sc.lights)
    g.frame(p)
    o = Oracle(sc, D, Lv, SS, W, H)
    o.frame(p)
    col, rad, img = g.read_volume(P.VOL_COLOR), g.read_volume(P.VOL_RADIANCE), g.read_image()
    assert (col >> 24 != 0).sum() > 500
    assert np.array_equal(col, o.color[0]) and np.array_equal(rad, o.radiance[0])
    assert np.array_equal(col, ref.read_volume(P.VOL_COLOR)) and np.array_equal(img, ref.read_image())
    a = img.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    b = o.image.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    mse = ((a - b) ** 2).mean()
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 45.0
    g.close(); ref.close()


def test_hostile_files_never_crash_the_host_process(tmp_path):
    """The reference's loaders index unchecked (tinyobjloader, Mesh::loadMesh) and trust image headers; a library must not:
    byte-mutated OBJ/MTL/DDS files, out-of-range and huge relative indices, PNG headers promising terabytes and a
    decompression bomb all end in a status + log line (or a smaller mesh), in-process."""
    import struct
    import zlib
    rng = np.random.default_rng(2024)
    lib = L.load()

    def obj_status(data, name="f.obj"):
        p = tmp_path / name; p.write_bytes(data)
        h = C.c_void_p()
        rc = lib.vct_ingest_obj(str(p).encode(), str(tmp_path).encode(), 1, C.byref(h))      # 1 = VCT_INGEST_NO_TEXTURES
        m = P.IngestMesh(); lib.vct_ingest_get_mesh(h, C.byref(m))
        n = (m.n_vertices, m.n_indices)
        if m.n_indices:                                                     # every index must address a vertex
            idx = np.ctypeslib.as_array(m.indices, (m.n_indices,))
            assert idx.max() < m.n_vertices
        log = lib.vct_ingest_log(h).decode(errors="replace")
        lib.vct_ingest_free(h)
        return rc, n, log

    good = open(os.path.join(GOLD, "fan.obj"), "rb").read()
    (tmp_path / "mats.mtl").write_bytes(open(os.path.join(GOLD, "mats.mtl"), "rb").read())
    assert obj_status(good)[1] == (18, 24)
    rc, n, log = obj_status(b"v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1 2 3\nf 1 2 9\nf -7 1 2\nf 1/5/1 2/1/9 3//-4\nf 2147483647 1 2\nf -2147483648 1 2\n")
    assert rc == 0 and n == (4, 6) and "dropped" in log                     # two valid triangles survive; out-of-range vt / vn read as absent
    for k in range(300):
        d = bytearray(good)
        for _ in range(1 + k % 6):
            d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        obj_status(bytes(d))
    mtl = bytearray(open(os.path.join(GOLD, "mats.mtl"), "rb").read())
    for k in range(100):
        d = bytearray(mtl)
        for _ in range(1 + k % 4):
            d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        (tmp_path / "mats.mtl").write_bytes(bytes(d))
        obj_status(good)

    def img_status(data, name):
        p = tmp_path / name; p.write_bytes(data)
        h = C.c_void_p()
        rc = lib.vct_ingest_image(str(p).encode(), 1, C.byref(h))
        log = lib.vct_ingest_log(h).decode(errors="replace")
        lib.vct_ingest_free(h)
        return rc, log

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
    sig = b"\x89PNG\r\n\x1a\n"
    huge = sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 23, 1 << 23, 16, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    rc, log = img_status(huge, "huge.png")
    assert rc == 1 and "too large" in log
    bomb = sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 4, 4, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * (64 << 20), 9)) + chunk(b"IEND", b"")
    assert len(bomb) < 100000
    rc, log = img_status(bomb, "bomb.png")
    assert rc == 1 and "more data than the image needs" in log
    short = sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 64, 64, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 100)) + chunk(b"IEND", b"")
    assert img_status(short, "short.png") == (1, "png: not enough pixel data")
    dds = bytearray(_dds(16, 16, b"DXT5", 3, bytes(rng.integers(0, 256, 16 * 16 + 64 + 16, dtype=np.uint8))))
    assert img_status(bytes(dds), "ok.dds")[0] == 0
    for k in range(200):
        d = bytearray(dds)
        for _ in range(1 + k % 5):
            d[int(rng.integers(0, 128))] = int(rng.integers(0, 256))       # header bytes: sizes, mip counts, fourcc
        img_status(bytes(d), "m.dds")
