"""Multi-GPU parity on real devices (needs >= 2 GPUs; skipped on a 1-GPU box): the z-slab sharded frame with the slab exchange
inside the library (csrc/exchange.cu: peer memory + device-side flags) reproduces the single-GPU pyramid, counters and image word
for word — one process per GPU (torchrun, cudaIpc) and one process driving all GPUs (vct_config.n_devices, C++ host --gpus N).
The N>1 host logic itself is covered on CPU by tests/test_sharded_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


@pytest.mark.parametrize("workload", ["room", "room_msaa", "sponza", "animated"])
def test_sharded_frame_equals_single_gpu(workload):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    if workload != "room":
        from vct_b200 import scene as S
        if not S.baked_available("sponza_pbr") or (workload == "animated" and not S.baked_available("nanosuit")):
            pytest.skip("baked assets missing")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "sharded_parity.py"), workload]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert rep["ok"], rep


def test_single_process_multi_device_handle_equals_single_gpu():
    """vct_config.n_devices: ONE handle, the library fans every call out to a context per device (worker threads), shards the frame
    and leaves the image on device 0.  Same words as the single-GPU frame: volumes (assembled from the ranks' slabs), counters, image."""
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np
    from vct_b200 import params as P
    from vct_b200 import scene as S
    from vct_b200.pipeline import Pipeline
    D, L, SS, W, H = 64, 5, 512, 320, 240
    sc = S.room_scene()
    one = Pipeline(sc, D, L, SS, W, H)
    grp = Pipeline(sc, D, L, SS, W, H, devices=[0, 1])
    try:
        for k, mods in enumerate(({}, {}, {"warp_texture": 1, "temporal_filter_radiance": 1}, {"warp_texture": 1, "temporal_filter_radiance": 1}, {})):
            p = S.room_params(W, H)
            for key, v in mods.items():
                setattr(p, key, v)
            one.frame(p); grp.frame(p)
            for which in (P.VOL_COLOR, P.VOL_NORMAL, P.VOL_RADIANCE):
                assert np.array_equal(one.read_volume(which), grp.read_volume(which)), (k, which)
            for l in range(1, L):
                assert np.array_equal(one.read_volume(P.VOL_RADIANCE, l), grp.read_volume(P.VOL_RADIANCE, l)), (k, l)
            assert np.array_equal(one.read_image(), grp.read_image()), k
            a, b = one.counters(), grp.counters()
            assert (a.total_fragments, a.unique_voxels, a.max_fragments_per_voxel) == (b.total_fragments, b.unique_voxels, b.max_fragments_per_voxel), k
            assert one.cone_steps() == grp.cone_steps(), k
    finally:
        grp.close(); one.close()


def test_cpp_host_gpus_flag_renders_the_same_image(tmp_path):
    """vct_headless --gpus 2 (no torch in the process) against --gpus 1: identical image hash and counters."""
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(ROOT, "vct_b200", "lib", "vct_headless")
    scene = str(tmp_path / "room.vcts")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "pack_scene.py"), "room", scene], cwd=ROOT)
    outs = []
    for gpus in ("1", "2"):
        r = subprocess.run([exe, scene, "--dim", "64", "--levels", "5", "--size", "320x240", "--shadow", "512", "--volume", "-1.5", "1.5",
                            "--eye", "1.1", "0.3", "1.2", "--front", "-0.65", "-0.25", "-0.72", "--frames", "3", "--fused", "--gpus", gpus],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]))
    a, b = outs
    assert b["gpus"] == 2 and a["ok"] and b["ok"]
    for key in ("image_fnv1a", "image_byte_sum", "total_fragments", "unique_voxels", "max_fragments_per_voxel"):
        assert a[key] == b[key], (key, a[key], b[key])
