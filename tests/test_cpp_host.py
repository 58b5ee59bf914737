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
