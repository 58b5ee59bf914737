"""CPU-only checks of the drop-in boundary and the host-side mirror (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from vct_b200 import params as P
from vct_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "vct_b200.h")).read()


def declared_symbols():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(vct_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    from vct_b200 import lib
    so = lib.load()
    decl = declared_symbols()
    assert len(decl) >= 40
    assert sorted(lib.SYMBOLS) == decl, "vct_b200/lib.py SYMBOLS must list exactly what include/vct_b200.h declares"
    for name in decl:
        assert hasattr(so, name), f"libvct_b200.so does not export {name}"


def test_library_has_no_torch_or_oracle_dependency():
    import subprocess
    from vct_b200 import lib
    out = subprocess.run(["ldd", lib.SO], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcudart" not in out      # cudart is linked statically


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vct_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in src and "vct_oracle" not in src and "libvct_oracle" not in src, f


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from vct_b200.lib import VctError
    from vct_b200.pipeline import Pipeline
    with pytest.raises(VctError, match="no CUDA device|no CPU fallback|CUDA"):
        Pipeline(None, 64, 5, 256, 64, 64)


def test_struct_layouts_match_the_header(tmp_path):
    """Compile the public header with gcc (plain C) and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess
    structs = {"vct_config": P.Config, "vct_light": P.Light, "vct_material": P.Material, "vct_voxelize_info": P.VoxelizeInfo,
               "vct_cone_settings": P.ConeSettings, "vct_frame_params": P.FrameParams, "vct_timings": P.Timings, "vct_peer": P.Peer}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "vct_b200.h")}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"; src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert C.sizeof(P.Light) == 80                                     # Scene.h:33 glslSize
    assert C.sizeof(P.VoxelizeInfo) == 12                              # Application.h:195-198


def test_glm_equivalents():
    # perspective with fov=45.0 used as RADIANS (Camera.h:18) -> SURVEY §3.2 pass 0 constants
    m = P.perspective(45.0, 16.0 / 9.0, 0.1, 100.0)
    assert m[1][1] == pytest.approx(1.792591, rel=2e-6) and m[0][0] == pytest.approx(1.008332, rel=2e-6)
    assert m[2][2] == pytest.approx(-1.002002, rel=1e-6) and m[3][2] == pytest.approx(-0.2002002, rel=1e-6) and m[2][3] == -1
    o = P.ortho(-25, 25, -25, 25, 0, 100)
    assert o[0][0] == pytest.approx(0.04) and o[2][2] == pytest.approx(-0.02) and o[3][2] == pytest.approx(-1.0)
    v = P.look_at((12, 40, -7), (12 - 0.38, 40 - 0.88, -7 + 0.2), (0, 1, 0))
    r = np.array(v)[:3, :3]
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-6)
    eye_h = np.array([12, 40, -7, 1], np.float32)
    assert np.allclose(eye_h @ np.array(v), [0, 0, 0, 1], atol=1e-4)       # the eye maps to the view-space origin
    a, b = P.translate_matrix((1, 2, 3)), P.scale_matrix(2.0)
    ab = P.matmul(a, b)                                                # GLM a*b: scale first, then translate
    assert np.allclose(np.array([1, 1, 1, 1], np.float32) @ ab, [3, 4, 5, 1])
    assert np.allclose(P.matmul(ab, P.inverse(ab)), np.eye(4), atol=1e-6)


def test_voxelisation_views_agree_with_voxel_linear_position():
    """for a symmetric cube volume, un-swizzling each axis view (voxelize.frag:87-98) gives exactly
    (world - center - min)/(max - min)."""
    cam = P.Camera(position=(5, 1, 0), yaw=180.0)
    light = P.reference_lights()[0]
    p = P.default_params(64, 64, cam, light, voxel_min=-3.0, voxel_max=3.0, voxel_center=(0.5, -1.0, 2.0))
    rng = np.random.default_rng(0)
    for _ in range(50):
        w = rng.uniform(-2.5, 2.5, 3).astype(np.float32) + np.array([0.5, -1.0, 2.0], np.float32)
        expect = (w - np.array([0.5, -1.0, 2.0], np.float32) + 3.0) / 6.0
        for axis, name in enumerate(("mvp_x", "mvp_y", "mvp_z")):
            m = np.array(getattr(p, name)[:], np.float32).reshape(4, 4)
            ndc = (np.append(w, np.float32(1)) @ m)[:3]
            u = (ndc + 1) * 0.5
            if axis == 0:
                u = np.array([1 - u[2], u[1], u[0]])
            elif axis == 1:
                u = np.array([u[0], 1 - u[2], u[1]])
            u[2] = 1 - u[2]
            assert np.allclose(u, expect, atol=2e-6), (axis, u, expect)


def test_reference_camera_and_lights():
    cam = P.Camera(position=(5, 1, 0), yaw=180.0)                      # Application.cpp:139-141
    assert np.allclose(cam.front, [-1, 0, 0], atol=1e-6)
    lights = P.reference_lights()
    assert lights[0].type == 1 and lights[0].shadow_caster == 1 and lights[1].type == 0 and lights[1].range == 5.0


def test_scene_containers_and_mips():
    sc = S.room_scene()
    verts, vact, idx, tmat, models = sc.flat()
    assert verts.shape[1] == 14 and idx.max() < len(verts) and len(tmat) * 3 == len(idx) == 3 * sc.n_tris
    assert models.shape == (len(sc.meshes), 16)
    t = S.Texture(np.arange(64, dtype=np.uint8).reshape(8, 8))
    assert [l.shape[:2] for l in t.levels] == [(8, 8), (4, 4), (2, 2), (1, 1)]
    assert t.levels[1][0, 0, 0] == (0 + 1 + 8 + 9 + 2) >> 2
    soup = S.soup_mesh(1000)
    assert soup.indices.size == 3000 and np.isnan(soup.vertices[:, 8:14]).all()   # no UVs -> NaN tangents (Mesh.cpp:178,196)


def test_default_params_follow_reference_settings():
    cam = P.Camera(position=(5, 1, 0), yaw=180.0)
    p = P.default_params(1920, 1080, cam, P.reference_lights()[0])
    assert (p.voxelize_atomic_max, p.deterministic, p.voxelize_lighting) == (0, 1, 1)
    assert p.voxel_set_opacity == 0.5 and p.temporal_decay == pytest.approx(0.8)
    assert (p.diffuse_cone.steps, p.specular_cone.steps) == (16, 32)       # Application.h:96-97
    assert p.diffuse_cone.cone_angle == pytest.approx(np.radians(60.0)) and p.specular_cone.lod_offset == pytest.approx(0.1)
    ref_default = P.default_params(1920, 1080, cam, P.reference_lights()[0], parity=False)
    assert ref_default.voxelize_atomic_max == 1                        # Application.h:100
