"""Debug views of phong.frag through the C ABI: vct_frame_params::debug_view selects what the reference's
Settings::drawVoxels / drawNormals / drawDominantAxis / debug* flags select (src/Application.cpp:992-1005, phong.frag:346-447,
489-505).  The oracle's version of every view is pinned to the reference's phong.frag compiled as C++
(tests/test_glsl_ref.py, modes view_*); here the CUDA kernel is compared with the oracle: same PSNR gate as the final image
(>= 45 dB; the texture unit filters with 8-bit weights), and the shaded frame must be untouched by the extra instantiation."""
import numpy as np
import pytest

from tests.oracle_lib import Oracle
from tests.test_glsl_ref import pbr_room
from tests.test_gpu_parity import psnr
from vct_b200 import params as P
from vct_b200 import scene as S
from vct_b200.lib import VctError

pytestmark = pytest.mark.gpu

D, L, SS, W, H = 64, 5, 512, 320, 240

VIEWS = [
    ("voxels_point", {"debug_view": P.VIEW_VOXELS, "miplevel": 0.0}),
    ("voxels_lod1_7", {"debug_view": P.VIEW_VOXELS, "miplevel": 1.7}),
    ("voxels_color_volume", {"debug_view": P.VIEW_VOXELS, "miplevel": 0.8, "draw_radiance": 0}),
    ("material_diffuse", {"debug_view": P.VIEW_MATERIAL_DIFFUSE}),
    ("material_roughness", {"debug_view": P.VIEW_MATERIAL_ROUGHNESS}),
    ("material_metallic", {"debug_view": P.VIEW_MATERIAL_METALLIC}),
    ("normals", {"debug_view": P.VIEW_NORMALS}),
    ("dominant_axis", {"debug_view": P.VIEW_DOMINANT_AXIS}),
    ("indirect", {"debug_view": P.VIEW_INDIRECT}),
    ("indirect_no_occlusion", {"debug_view": P.VIEW_INDIRECT, "draw_occlusion": 0}),
    ("occlusion", {"debug_view": P.VIEW_OCCLUSION}),
    ("reflections", {"debug_view": P.VIEW_REFLECTIONS}),
    ("voxels_warp_voxels", {"debug_view": P.VIEW_VOXELS, "miplevel": 1.2, "warp_voxels": 1}),
    ("indirect_warp_texture", {"debug_view": P.VIEW_INDIRECT, "warp_texture": 1}),
    ("voxel_normals", {"debug_view": P.VIEW_VOXEL_NORMALS}),
    ("voxel_normals_warp_voxels", {"debug_view": P.VIEW_VOXEL_NORMALS, "miplevel": 2.3, "warp_voxels": 1}),
    ("warp_texture", {"debug_view": P.VIEW_WARP_TEXTURE, "warp_texture": 1}),
    ("warp_texture_tc", {"debug_view": P.VIEW_WARP_TEXTURE_TC, "warp_texture": 1}),
]


def test_debug_views_match_the_oracle():
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    report = {}
    base = S.room_params(W, H)
    g.frame(base)
    shaded = g.read_image().copy()
    for name, kw in VIEWS:
        p = S.room_params(W, H)
        for k, v in kw.items():
            setattr(p, k, v)
        o = Oracle(sc, D, L, SS, W, H)
        o.frame(p)
        g.frame(p)
        img = g.read_image()
        assert np.unique(o.image).size > 2, name                     # the view shows something
        report[name] = psnr(img, o.image)
    print({k: round(v, 2) for k, v in report.items()})
    bad = {k: v for k, v in report.items() if v < 45.0}
    assert not bad, f"debug views below 45 dB: {bad}"
    # 2D material lookups are filtered in software with fp32 weights on both sides: only fast-math rounding of the interpolated
    # uv separates them (measured 76 dB and better)
    for name in ("material_diffuse", "material_roughness", "material_metallic"):
        assert report[name] >= 60.0, (name, report[name])
    g.frame(base)
    assert np.array_equal(g.read_image(), shaded), "the shaded frame changed after debug views were rendered"
    p = S.room_params(W, H); p.debug_view = 77
    with pytest.raises(VctError, match="debug_view"):
        g.frame(p)
    g.close()


def test_pipelined_image_readback_is_ordered_with_the_next_frame():
    """vct_read_image_async: frame A's image is copied on the copy stream while frame B (a different view, so every pixel
    differs) is already being computed; B's cone trace must wait for A's copy.  Both host buffers hold their own frame."""
    import torch
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    pa = S.room_params(W, H)
    pb = S.room_params(W, H); pb.debug_view = P.VIEW_NORMALS
    g.frame(pa); want_a = g.read_image().copy()
    g.frame(pb); want_b = g.read_image().copy()
    assert (want_a != want_b).mean() > 0.5
    bufs = [torch.zeros(W * H, dtype=torch.int32).pin_memory() for _ in range(2)]
    for rounds in range(20):                                     # many back-to-back pairs: a missing wait shows up as torn images
        g.frame(pa); g.read_image_async(bufs[0].data_ptr())
        g.frame(pb); g.read_image_async(bufs[1].data_ptr())
        g.gi_passes(pa)                                          # a third frame in flight behind both copies
        g.read_image_wait(block_host=True)
        g.sync()
        assert np.array_equal(bufs[0].numpy().view(np.uint32), want_a), f"round {rounds}: buffer A torn"
        assert np.array_equal(bufs[1].numpy().view(np.uint32), want_b), f"round {rounds}: buffer B torn"
        bufs[0].zero_(); bufs[1].zero_()
    g.read_image_wait(block_host=False)                          # nothing pending: a no-op
    assert np.array_equal(g.read_image(), want_a)                # the blocking call still works (image of the last frame: pa)
    g.close()


def test_debug_voxels_cubes_match_the_oracle():
    """Application::debugVoxels (Application.cpp:1222-1275, debugVoxels.vert/.geom/.frag): the traced pyramid's non-empty voxels as depth-tested,
    back-face-culled cubes.  At lod <= 0.5 the voxel colours are texel fetches: the image must equal the oracle's to the last bit (same
    rasteriser, same depth rule, same draw order); at a fractional lod the texture unit filters with 8-bit weights: >= 45 dB."""
    from vct_b200.pipeline import Pipeline
    sc = pbr_room()
    g = Pipeline(sc, D, L, SS, W, H)
    o = Oracle(sc, D, L, SS, W, H)
    try:
        base = S.room_params(W, H)
        o.frame(base); g.frame(base)
        shaded = g.read_image().copy()
        report = {}
        for name, kw in (("point", {"miplevel": 0.0}), ("color_volume", {"miplevel": 0.3, "draw_radiance": 0}), ("lod1_4", {"miplevel": 1.4}), ("lod3", {"miplevel": 3.0})):
            p = S.room_params(W, H)
            out = P.default_params(W, H, P.Camera(position=(3.4, 2.2, 3.9), front=(-0.62, -0.38, -0.69)), sc.lights[0], voxel_min=-1.5, voxel_max=1.5)
            p.projection, p.view, p.pv, p.eye = out.projection, out.view, out.pv, out.eye       # from outside: the whole volume as cubes
            for k, v in kw.items():
                setattr(p, k, v)
            o.debug_voxels(p); g.debug_voxels(p)
            img = g.read_image()
            assert np.unique(o.image).size > 20, name
            report[name] = (psnr(img, o.image), int((img != o.image).sum()))
        print(report)
        assert report["point"][1] == 0 and report["color_volume"][1] == 0, report
        assert report["lod1_4"][0] >= 45.0 and report["lod3"][0] >= 45.0, report
        # the frame's own camera, inside the volume: cubes cut by the near plane and triangles hundreds of pixels wide (the queued path)
        p = S.room_params(W, H)
        o.debug_voxels(p); g.debug_voxels(p)
        img = g.read_image()
        assert int((img != o.image).sum()) == 0 and np.unique(img).size > 20
        g.frame(base)
        assert np.array_equal(g.read_image(), shaded)
    finally:
        g.close()
