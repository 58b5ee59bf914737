"""The C++ host mirror (vct_b200/host/vct_host.hpp + vct_headless): the reference's Application::render with the GL
blocks replaced by vct_* calls.  CPU: it builds, links only the C-ABI library, and fails loudly without a GPU.
GPU: its frame equals the frame the Python mirror produces for the same configuration."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "vct_b200", "lib", "vct_headless")


def _pack(tmp_path, which):
    from tools import pack_scene
    from vct_b200 import scene as S
    sc = S.room_scene() if which == "room" else S.config_scene(which)[0]
    path = str(tmp_path / f"{which}.vcts")
    pack_scene.pack(sc, path)
    return sc, path


def test_headless_builds_and_prints_usage():
    assert os.path.isfile(EXE), "run __graft_entry__.build()"
    r = subprocess.run([EXE, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage: vct_headless" in r.stdout
    needed = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libvct_b200.so" in needed and "torch" not in needed and "oracle" not in needed


def test_headless_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, path = _pack(tmp_path, 1)
    r = subprocess.run([EXE, path, "--dim", "64", "--size", "64x64", "--shadow", "256"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not r.stdout.strip()


def test_scene_file_round_trip(tmp_path):
    """tools/pack_scene.py writes exactly the bytes vct_host::Scene::load expects (header + sizes re-derived here)."""
    sc, path = _pack(tmp_path, "room")
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:24], np.uint32)
    assert hdr[0] == 0x53544356 and hdr[1] == 1 and list(hdr[2:]) == [len(sc.textures), len(sc.materials), len(sc.meshes), len(sc.lights)]
    size = 24 + sum(20 + t.packed().nbytes for t in sc.textures) + 40 * len(sc.materials) + 80 * len(sc.lights)
    size += sum(8 + m.vertices.nbytes + m.indices.nbytes + m.tri_material.nbytes + 64 for m in sc.meshes)
    assert size == len(raw)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_headless_frame_equals_python_mirror(tmp_path, fused):
    from vct_b200 import params as P
    from vct_b200 import scene as S
    from vct_b200.pipeline import Pipeline
    sc, cam, (vmin, vmax, vc), D, (W, H), _ = S.config_scene(1)
    _, path = _pack(tmp_path, 1)
    out = str(tmp_path / "frame.ppm")
    cmd = [EXE, path, "--dim", str(D), "--size", f"{W}x{H}", "--shadow", "1024", "--frames", "2", "--eye", "2.5", "1.5", "2.5",
           "--front", "-2.5", "-1.5", "-2.5", "--volume", str(vmin), str(vmax), "--out", out] + (["--fused"] if fused else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["ok"] and rep["unique_voxels"] > 50
    p = P.default_params(W, H, cam, sc.lights[0], voxel_min=vmin, voxel_max=vmax, voxel_center=vc)
    g = Pipeline(sc, D, 6, 1024, W, H)
    try:
        g.frame(p); g.frame(p)
        info = g.counters()
        assert (rep["total_fragments"], rep["unique_voxels"], rep["max_fragments_per_voxel"]) == (info.total_fragments, info.unique_voxels, info.max_fragments_per_voxel)
        img = g.image_rgba()[:, :, :3].astype(np.float64)
    finally:
        g.close()
    ppm = open(out, "rb").read()
    head = f"P6\n{W} {H}\n255\n".encode()
    assert ppm.startswith(head)
    got = np.frombuffer(ppm[len(head):], np.uint8).reshape(H, W, 3).astype(np.float64)
    mse = ((got - img) ** 2).mean()
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    # the two hosts build the same matrices with different float libraries (numpy fp32 vs C++): not bit-identical inputs
    assert psnr >= 45.0, psnr
    if fused:
        assert rep["timers_ms"]["total"] > 0


FRAME_PARAMS_MAIN = r"""
#include <cstdio>
#include "vct_host.hpp"
using namespace vct_host;
int main(int argc, char** argv) {
    Scene sc; sc.addReferenceLights();
    Application app; app.scene = &sc; app.width = 640; app.height = 360;
    app.camera.position = {2.5f, 1.5f, -3.0f}; app.camera.hasFront = true; app.camera.frontOverride = {-2.0f, -1.0f, 2.5f};
    app.vct.min = {-3.f, -3.f, -3.f}; app.vct.max = {3.f, 3.f, 3.f}; app.vct.center = {0.5f, 1.0f, -0.25f};
    if (argc > 1) {   // every Settings member that reaches the kernels, moved off its default
        Settings& s = app.settings;
        s.voxelizeLighting = 0; s.voxelizeAtomicMax = 1; s.axisOverride = 2; s.voxelSetOpacity = 0.25f; s.temporalFilterRadiance = 1; s.temporalDecay = 0.6f;
        s.radianceLighting = 1; s.voxelFillHoles = 1; s.warpVoxels = 1; s.warpTexture = 1; s.warpTextureLinear = 1; s.warpTextureAxes[1] = 0;
        s.useWarpmapWeightsTexture = 0; s.warpTextureHighResolution = 3.0f; s.warpTextureLowResolution = 0.25f;
        s.drawRadiance = 0; s.drawOcclusion = 0; s.cooktorrance = 0; s.enablePostprocess = 0; s.enableNormalMap = 0;
        s.enableIndirect = 0; s.enableDiffuse = 0; s.enableSpecular = 0; s.enableReflections = 0; s.ambientScale = 0.5f; s.reflectScale = 2.0f;
        s.diffuseConeSettings.steps = 9; s.specularConeSettings.bias = 2.5f; s.specularConeAngleFromRoughness = 0;
        s.debugOcclusion = 1; s.drawNormals = 1;          // the shader tests drawNormals first
        s.miplevel = 1.5f; s.voxelizeTesselation = 1; s.voxelizeTesselationWarp = 1; s.voxelizeMultiplier = 2.0f; s.conservativeRasterization = Settings::MSAA;
    }
    const vct_frame_params p = app.frameParams();
    std::fwrite(&p, sizeof p, 1, stdout);
    return 0;
}
"""


def test_cpp_host_frame_parameters_equal_the_python_mirror(tmp_path):
    """vct_host.hpp::Application::frameParams (the C++ mirror of the uniforms Application::render sets) against
    vct_b200.params.default_params (the Python mirror the parity tests use): same integers, same scalars, matrices equal to a few
    ulp of their largest entry (the two hosts use different float libraries).  Runs without a GPU: no context is created."""
    import ctypes as C
    from vct_b200 import params as P
    src = tmp_path / "fp.cpp"; exe = tmp_path / "fp"
    src.write_text(FRAME_PARAMS_MAIN)
    lib = os.path.join(ROOT, "vct_b200", "lib")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f"-I{ROOT}/vct_b200/host", f"-I{ROOT}/include", str(src), "-o", str(exe),
                           f"-L{lib}", "-lvct_b200", f"-Wl,-rpath,{lib}"])
    cam = P.Camera(position=(2.5, 1.5, -3.0), front=(-2.0, -1.0, 2.5))
    light = P.reference_lights()[0]
    for changed in (False, True):
        raw = subprocess.run([str(exe)] + (["x"] if changed else []), capture_output=True, check=True).stdout
        assert len(raw) == C.sizeof(P.FrameParams)
        got = P.FrameParams.from_buffer_copy(raw)
        want = P.default_params(640, 360, cam, light, voxel_min=-3.0, voxel_max=3.0, voxel_center=(0.5, 1.0, -0.25))
        if changed:
            for k, v in dict(voxelize_lighting=0, voxelize_atomic_max=1, axis_override=2, voxel_set_opacity=0.25, temporal_filter_radiance=1, temporal_decay=0.6,
                             radiance_lighting=1, voxel_fill_holes=1, warp_voxels=1, warp_texture=1, warp_texture_linear=1, use_warpmap_weights_texture=0,
                             warp_texture_high_resolution=3.0, warp_texture_low_resolution=0.25, draw_radiance=0, draw_occlusion=0, cooktorrance=0,
                             enable_postprocess=0, enable_normal_map=0, enable_indirect=0, enable_diffuse=0, enable_specular=0, enable_reflections=0,
                             ambient_scale=0.5, reflect_scale=2.0, specular_cone_angle_from_roughness=0, debug_view=P.VIEW_NORMALS, miplevel=1.5,
                             voxelize_tesselation=1, voxelize_tesselation_warp=1, voxelize_multiplier=2.0, conservative_raster=P.RASTER_MSAA).items():
                setattr(want, k, v)
            want.warp_texture_axes[1] = 0; want.diffuse_cone.steps = 9; want.specular_cone.bias = 2.5
        seen = 0
        for name, ctype in P.FrameParams._fields_:
            a, b = getattr(got, name), getattr(want, name)
            if isinstance(a, C.Array) and a._type_ is C.c_float:
                x, y = np.array(a[:], np.float32), np.array(b[:], np.float32)
                assert np.all(np.abs(x - y) <= 8 * np.finfo(np.float32).eps * max(1.0, float(np.abs(y).max()))), (name, x, y)
            elif isinstance(a, C.Array):
                assert a[:] == b[:], (name, a[:], b[:])
            elif isinstance(a, C.Structure):
                for f, _ in a._fields_:
                    assert getattr(a, f) == pytest.approx(getattr(b, f), rel=1e-6), (name, f)
            elif ctype is C.c_float:
                assert a == pytest.approx(b, rel=1e-6), (name, a, b)
            else:
                assert a == b, (name, a, b)
            seen += 1
        assert seen == len(P.FrameParams._fields_)
