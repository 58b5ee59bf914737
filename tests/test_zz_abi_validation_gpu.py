"""Hostile arguments at the C ABI (needs a context, hence a GPU): every call returns non-zero with a message, nothing reaches a
kernel, and the context keeps working afterwards — the reference's convention of logging and carrying on (log.h:29-47).
Written after this round's GPU budget was spent (host-side checks only, no kernel involved); sorts after the other GPU tests."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_lib import Oracle
from vct_b200 import params as P
from vct_b200 import scene as S

pytestmark = pytest.mark.gpu


def test_bad_scene_arguments_are_rejected_and_the_context_survives():
    from vct_b200.pipeline import Pipeline
    D, L, SS, W, H = 32, 4, 128, 64, 48
    sc = S.room_scene()
    g = Pipeline(sc, D, L, SS, W, H)
    try:
        lib, h = g.lib, g.h

        def rejected(status, needle):
            assert status != 0
            msg = lib.vct_last_error(h).decode()
            assert needle in msg, msg

        m = P.Material(10 ** 6, -1, -1, -1, -1, -1, 32.0, P.F3(0, 0, 0))
        rejected(lib.vct_set_material(h, 7, C.byref(m)), "texture id out of range")
        m = P.Material(-5, -1, -1, -1, -1, -1, 32.0, P.F3(0, 0, 0))
        rejected(lib.vct_set_material(h, 7, C.byref(m)), "texture id out of range")
        v = np.zeros((3, 14), np.float32); v[:, 3:6] = [0, 0, 1]; v[1, 0] = v[2, 1] = 0.5
        idx = np.arange(3, dtype=np.uint32)
        for bad in (-1, 256, 10 ** 6):
            tm = np.full(1, bad, np.int32)
            rejected(lib.vct_upload_mesh(h, 40, v.ctypes.data, 3, 56, idx.ctypes.data, 3, tm.ctypes.data), "material id out of range")
        rejected(lib.vct_upload_mesh(h, 10 ** 6, v.ctypes.data, 3, 56, idx.ctypes.data, 3, None), "actor ids")
        rejected(lib.vct_upload_mesh(h, 40, v.ctypes.data, 3, 56, np.array([0, 1, 7], np.uint32).ctypes.data, 3, None), "index out of range")
        rejected(lib.vct_set_actor_transform(h, 0, None), "null matrix")
        # nothing above changed the scene: the frame still equals the oracle's
        p = S.room_params(W, H)
        o = Oracle(sc, D, L, SS, W, H); o.frame(p)
        g.frame(p)
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0]) and np.array_equal(g.read_volume(P.VOL_RADIANCE), o.radiance[0])
        # a material that names a texture slot nothing was uploaded to is caught before any kernel runs ...
        m = P.Material(200, -1, -1, -1, -1, -1, 32.0, P.F3(0, 0, 0))
        assert lib.vct_set_material(h, 9, C.byref(m)) == 0
        rejected(lib.vct_voxelize(h, C.byref(p)), "never uploaded")
        # ... and uploading it afterwards heals the context
        tex = np.full((2, 2, 3), 128, np.uint8)
        assert lib.vct_upload_texture(h, 200, 2, 2, 3, 1, tex.ctypes.data) == 0
        g.frame(p)
        assert np.array_equal(g.read_volume(P.VOL_COLOR), o.color[0])
    finally:
        g.close()
