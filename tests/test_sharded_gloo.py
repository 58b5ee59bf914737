"""Multi-rank host logic on CPU: world_size 2 and 4 with the gloo backend (the N>1 data path of vct_b200/sharded.py).

What is checked without a GPU: the z-slab / level-chunk / screen-band partition maths, that BOX2 mips of a slab need
no halo (the slab of the full-volume mip == the mip computed from a volume holding only that slab), and that the
in-place per-level all-gather reassembles exactly the single-process pyramid.  The oracle stands in for the kernels.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vct_b200 import sharded as SH

D, L, WORLD = 32, 4, 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


def _full_pyramid(seed=5):
    from tests import oracle_lib as ol
    rng = np.random.default_rng(seed)
    lv = [np.where(rng.random(D ** 3) < 0.2, rng.integers(0, 2 ** 32, D ** 3, dtype=np.uint64), 0).astype(np.uint32)]
    for l in range(L - 1):
        d = D >> l
        dst = np.zeros((d // 2) ** 3, np.uint32)
        ol.lib().orc_mip(d, ol.ptr(lv[-1]), ol.ptr(dst), 0)
        lv.append(dst)
    return lv


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_lib as ol
        full = _full_pyramid()
        z_lo, z_hi = SH.slab_range(D, world, rank)
        chunks = SH.level_chunks(D, L, world)
        # this rank's level 0: only its slab
        mine = [np.zeros_like(v) for v in full]
        n0 = chunks[0]
        mine[0][rank * n0:(rank + 1) * n0] = full[0][rank * n0:(rank + 1) * n0]
        assert rank * n0 == z_lo * D * D and (rank + 1) * n0 == z_hi * D * D
        # slab-local mip chain: level l+1 from the rank's own level l (rest of the volume is zero)
        for l in range(L - 1):
            d = D >> l
            dst = np.zeros((d // 2) ** 3, np.uint32)
            ol.lib().orc_mip(d, ol.ptr(mine[l]), ol.ptr(dst), 0)
            n = chunks[l + 1]
            mine[l + 1][rank * n:(rank + 1) * n] = dst[rank * n:(rank + 1) * n]
            assert np.array_equal(dst[rank * n:(rank + 1) * n], full[l + 1][rank * n:(rank + 1) * n]), "BOX2 slab mip must not need a halo"
        tens = [torch.from_numpy(v.view(np.int32)) for v in mine]
        SH.all_gather_levels(dist, tens, chunks, rank)
        ok = all(np.array_equal(t.numpy().view(np.uint32), f) for t, f in zip(tens, full))
        # image bands: every rank contributes its band, all-gather gives the whole (padded) image
        band, bands = SH.image_bands(20, world)
        W = 4
        img = torch.zeros(band * world * W, dtype=torch.int32)
        y0, y1 = bands[rank]
        img[y0 * W:y1 * W] = rank + 1
        dist.all_gather_into_tensor(img, img[rank * band * W:(rank + 1) * band * W].clone())
        expect = np.concatenate([np.full((b1 - b0) * W, r + 1) for r, (b0, b1) in enumerate(bands)])
        ok = ok and np.array_equal(img.numpy()[:20 * W], expect)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [WORLD, 4])          # 4: the coarsest level (4^3) is exactly one texel per slab
def test_slab_exchange_reassembles_single_process_pyramid(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}


def _stripe_worker(rank, world, port, q, stripe, levels):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_lib as ol
        rng = np.random.default_rng(9)
        full = [np.where(rng.random(D ** 3) < 0.2, rng.integers(0, 2 ** 32, D ** 3, dtype=np.uint64), 0).astype(np.uint32)]
        for l in range(levels - 1):
            dst = np.zeros(((D >> l) // 2) ** 3, np.uint32)
            ol.lib().orc_mip(D >> l, ol.ptr(full[-1]), ol.ptr(dst), 0)
            full.append(dst)
        own = SH.stripes(D, stripe, world, rank)
        top = SH.top_sharded_level(stripe, levels)
        mine = [np.zeros_like(v) for v in full]
        for lo, hi in own:
            mine[0][lo * D * D:hi * D * D] = full[0][lo * D * D:hi * D * D]
        for l in range(top):                                 # own stripes only: BOX2 needs no halo while texel layers stay inside a stripe
            d = D >> l
            dst = np.zeros((d // 2) ** 3, np.uint32)
            ol.lib().orc_mip(d, ol.ptr(mine[l]), ol.ptr(dst), 0)
            dd = d // 2
            for lo, hi in own:
                a, b = (lo >> (l + 1)) * dd * dd, (hi >> (l + 1)) * dd * dd
                assert b > a
                mine[l + 1][a:b] = dst[a:b]
                assert np.array_equal(dst[a:b], full[l + 1][a:b])
        tens = [torch.from_numpy(v.view(np.int32)) for v in mine[:top + 1]]
        for t in tens:                                       # the exchange: every layer has exactly one owner, the others hold zeros
            dist.all_reduce(t)
        got = [t.numpy().view(np.uint32) for t in tens]
        for l in range(top, levels - 1):                     # the tail: every rank, from the complete level below
            dst = np.zeros(((D >> l) // 2) ** 3, np.uint32)
            ol.lib().orc_mip(D >> l, ol.ptr(got[l]), ol.ptr(dst), 0)
            got.append(dst)
        q.put((rank, len(got) == levels and all(np.array_equal(a, b) for a, b in zip(got, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stripe,levels", [(2, 4, 5), (2, 8, 5), (4, 2, 4)])
def test_interleaved_stripes_reassemble_single_process_pyramid(world, stripe, levels):
    """Layers dealt out in stripes (vct_config.slab_stripe): levels up to log2(stripe) from the own stripes, the rest after the exchange."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stripe_worker, args=(r, world, port, q, stripe, levels)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}


def test_partition_maths():
    assert SH.stripes(256, 32, 8, 3) == [(96, 128)] == [SH.slab_range(256, 8, 3)]
    assert SH.stripes(256, 16, 8, 3) == [(48, 64), (176, 192)] and SH.stripes(512, 16, 2, 1)[:2] == [(16, 32), (48, 64)]
    assert all(sum(SH.stripe_owner(z, 16, 8) == r for z in range(256)) == 32 for r in range(8))
    assert all(SH.stripe_owner(z, 16, 8) == 3 for lo, hi in SH.stripes(256, 16, 8, 3) for z in range(lo, hi))
    assert SH.top_sharded_level(16, 6) == 4 and SH.top_sharded_level(32, 6) == 5 and SH.top_sharded_level(128, 6) == 5 and SH.top_sharded_level(16, 3) == 2
    with pytest.raises(ValueError):
        SH.stripes(256, 24, 8, 0)
    assert [SH.slab_range(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]
    assert SH.level_chunks(256, 6, 8) == [256 ** 3 // 8, 128 ** 3 // 8, 64 ** 3 // 8, 32 ** 3 // 8, 16 ** 3 // 8, 8 ** 3 // 8]
    with pytest.raises(ValueError):
        SH.level_chunks(64, 6, 8)                 # coarsest level 2^3 is thinner than 8 slabs
    with pytest.raises(ValueError):
        SH.slab_range(100, 8, 0)
    # screen tiles of the sharded cone trace (cone_trace.cu ensure_trace_tiles): 64x64 pixels, diagonal stripes, every tile owned once
    for (w, h, n) in ((1920, 1080, 8), (3840, 2160, 8), (320, 240, 2), (100, 70, 4)):
        per = [SH.screen_tiles(w, h, n, r) for r in range(n)]
        allt = sorted(t for p in per for t in p)
        assert allt == sorted((x, y) for y in range(0, h, 64) for x in range(0, w, 64)), (w, h, n)
        assert max(len(p) for p in per) - min(len(p) for p in per) <= max(1, -(-h // 64)), (w, h, n)      # balanced to within one tile per row
    assert SH.tile_owner(3, 2, 4) == 1 and SH.screen_tiles(128, 128, 2, 0) == [(0, 0), (64, 64)]
    band, bands = SH.image_bands(1080, 8)
    assert band == 136 and bands[0] == (0, 136) and bands[7] == (952, 1080) and sum(b - a for a, b in bands) == 1080
    band, bands = SH.image_bands(2160, 8)
    assert band * 8 >= 2160 and band % 8 == 0 and bands[-1][1] == 2160
    band, bands = SH.image_bands(8, 4)             # more ranks than tiles: trailing ranks get empty bands
    assert bands == [(0, 8), (8, 8), (8, 8), (8, 8)]
