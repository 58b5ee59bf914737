"""Sparse frames (segment masks): from the second vct_frame on, clear / transferVoxels / the mip chain visit
only the x-row segments flagged by the voxeliser.  The results must be what the dense kernels and the oracle produce —
bit for bit — while geometry moves between frames (segments go stale and must be cleared), with the temporal filter
(support only grows), and after something wrote a volume behind the masks' back."""
import os

import numpy as np
import pytest

from tests.oracle_lib import Oracle
from vct_b200 import params as P
from vct_b200 import scene as S

pytestmark = pytest.mark.gpu

D, L, SS, W, H = 64, 5, 512, 160, 120


def _psnr(a, b):
    a = a.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64); b = b.view(np.uint8).reshape(-1, 4)[:, :3].astype(np.float64)
    mse = ((a - b) ** 2).mean()
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _dense_pipeline(sc):
    from vct_b200.pipeline import Pipeline
    os.environ["VCT_SPARSE"] = "0"
    try:
        return Pipeline(sc, D, L, SS, W, H)
    finally:
        del os.environ["VCT_SPARSE"]


def _move(sc, o, pipes, actor, frame):
    m = P.matmul(P.translate_matrix((-0.8 + 0.37 * frame, -0.75 + 0.11 * frame, 0.6 - 0.29 * frame)), P.scale_matrix(0.22))
    o.s.models[actor] = np.asarray(m, np.float32).reshape(16)
    for g in pipes:
        g.set_actor_transform(actor, m)


def _check(o, g, gd, tag):
    for which, ref in ((P.VOL_COLOR, o.color), (P.VOL_RADIANCE, o.radiance)):
        for l in range(L):
            got = g.read_volume(which, l)
            assert np.array_equal(got, ref[l]), f"{tag}: volume {which} level {l} differs from the oracle in {(got != ref[l]).sum()} voxels"
            assert np.array_equal(got, gd.read_volume(which, l)), f"{tag}: volume {which} level {l} differs from the dense kernels"
    assert np.array_equal(g.read_volume(P.VOL_NORMAL), o.normal), tag
    a, b = g.counters(), o.info
    assert (a.total_fragments, a.unique_voxels, a.max_fragments_per_voxel) == (b.total_fragments, b.unique_voxels, b.max_fragments_per_voxel), tag
    img = g.read_image()
    assert np.array_equal(img, gd.read_image()), f"{tag}: image differs from the dense kernels (stale texels in the texture array?)"
    assert _psnr(img, o.image) >= 45.0, tag
    assert g.cone_steps() == gd.cone_steps()


@pytest.mark.parametrize("temporal", [0, 1])
def test_sparse_frames_with_moving_actor(temporal):
    from vct_b200.pipeline import Pipeline
    sc = S.room_scene(seed=7)
    p = S.room_params(W, H)
    p.temporal_filter_radiance = temporal
    o = Oracle(sc, D, L, SS, W, H)
    g = Pipeline(sc, D, L, SS, W, H)
    gd = _dense_pipeline(sc)
    try:
        actor = len(sc.meshes) - 1
        for f in range(5):
            _move(sc, o, (g, gd), actor, f)
            o.frame(p); g.frame(p); gd.frame(p)
            _check(o, g, gd, f"temporal={temporal} frame {f}")
    finally:
        g.close(); gd.close()


def test_sparse_frames_survive_outside_writes_and_mode_changes():
    from vct_b200.pipeline import Pipeline
    sc = S.room_scene(seed=9)
    p = S.room_params(W, H)
    o = Oracle(sc, D, L, SS, W, H)
    g = Pipeline(sc, D, L, SS, W, H)
    gd = _dense_pipeline(sc)
    try:
        for _ in range(2):
            o.frame(p); g.frame(p); gd.frame(p)
        _check(o, g, gd, "warm")
        # garbage written straight into the volumes (all levels): the masks know nothing about it
        rng = np.random.default_rng(1)
        for which in (P.VOL_COLOR, P.VOL_RADIANCE):
            for l in range(L):
                junk = rng.integers(0, 2 ** 32, (D >> l) ** 3, dtype=np.uint32)
                g.write_volume(which, l, junk); gd.write_volume(which, l, junk)
        g.write_volume(P.VOL_NORMAL, 0, rng.integers(0, 2 ** 32, D ** 3, dtype=np.uint32))
        o.frame(p); g.frame(p); gd.frame(p)
        _check(o, g, gd, "after outside writes")
        # pass-by-pass calls in between, then frames again
        g.voxelize(p); g.transfer(p); gd.voxelize(p); gd.transfer(p)
        o.frame(p); g.frame(p); gd.frame(p)
        _check(o, g, gd, "after pass-by-pass calls")
        # atomic-max voxelisation and the colour pyramid as the traced volume
        q = type(p).from_buffer_copy(p); q.voxelize_atomic_max = 1; q.draw_radiance = 0
        for f in range(3):
            _move(sc, o, (g, gd), len(sc.meshes) - 1, f)
            o.frame(q); g.frame(q); gd.frame(q)
            _check(o, g, gd, f"atomic max / colour pyramid frame {f}")
        o.frame(p); g.frame(p); gd.frame(p)
        _check(o, g, gd, "back to radiance")
        # fill holes dilates the support: dense frame, masks rebuilt afterwards
        q = type(p).from_buffer_copy(p); q.voxel_fill_holes = 1
        o.frame(q); g.frame(q); gd.frame(q)
        _check(o, g, gd, "fill holes")
        o.frame(p); g.frame(p); gd.frame(p)
        _check(o, g, gd, "after fill holes")
    finally:
        g.close(); gd.close()


def test_cuda_graph_replay_matches_eager_steps():
    """bench.py replays a captured PAIR of steps (the masks swap roles every frame); the replayed frames must leave
    exactly the volumes, counters and image of host-launched frames, also when eager steps come in between."""
    import torch
    from vct_b200.pipeline import Pipeline
    from vct_b200.sharded import ShardedFrame
    sc = S.room_scene(seed=11)
    p = S.room_params(W, H)
    o = Oracle(sc, D, L, SS, W, H)
    o.frame(p)
    g = Pipeline(sc, D, L, SS, W, H)
    try:
        fr = ShardedFrame(g, p, 1, 0)
        fr.producers()
        for _ in range(3):
            fr.step()
        assert fr.enable_graph(), getattr(fr, "graph_error", None)
        assert fr.run_steps(4) == 4
        fr.step()                                  # odd eager step: parity differs from the captured one
        assert fr.run_steps(3) == 2                # one eager step restores it, then one replay
        torch.cuda.synchronize()
        for l in range(L):
            assert np.array_equal(g.read_volume(P.VOL_RADIANCE, l), o.radiance[l]), f"level {l}"
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])
        a, b = g.counters(), o.info
        assert (a.total_fragments, a.unique_voxels, a.max_fragments_per_voxel) == (b.total_fragments, b.unique_voxels, b.max_fragments_per_voxel)
        assert _psnr(g.read_image(), o.image) >= 45.0
    finally:
        g.close()
