"""Multi-GPU parity on real devices (needs >= 2 GPUs; skipped on a 1-GPU box): the z-slab sharded frame + NCCL
exchange of vct_b200/sharded.py reproduces the single-GPU pyramid, counters and image word for word.
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


@pytest.mark.parametrize("workload", ["room", "sponza"])
def test_sharded_frame_equals_single_gpu(workload):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    if workload == "sponza":
        from vct_b200 import scene as S
        if not S.baked_available("sponza_pbr"):
            pytest.skip("assets/_baked/sponza_pbr missing")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "sharded_parity.py"), workload]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert rep["ok"], rep
