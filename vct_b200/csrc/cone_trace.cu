// cone_trace.cu — per-pixel shading + voxel cone tracing (a7).  Tolerance-gated (final image PSNR), compiled
// with FMA contraction on; the voxel pyramid is read through HARDWARE 3D mipmapped texture objects (the 32^3
// warp map is filtered in software with fp32 weights, see warp_sample below).
//
// Replaces the GL_EQUAL colour pass of phong.vert/phong.frag (reference src/Application.cpp:967-1067):
//   phong.frag:427-439  normal (normal map through the interpolated TBN, or the vertex normal)
//   phong.frag:305-344  direct lighting (Cook-Torrance :230-258 / Blinn :260-301, 5-tap PCF :183-207)
//   phong.frag:455-512  6 weighted diffuse cones + 1 specular cone through traceCone (:135-180)
//   phong.frag:210-218  Reinhard + gamma
// Inputs: the visibility buffer (triangle id per pixel) — attributes are re-interpolated here with the
// perspective-correct barycentrics of the unclipped triangle, so no fat G-buffer is stored.
// Sampler state reproduced (src/Application.cpp:1094-1098, Application.h:146): min LINEAR_MIPMAP_LINEAR, mag
// NEAREST, CLAMP_TO_BORDER(0): lambda <= 0.5 is a magnification -> point fetch of level 0 (GL 4.5 §8.14),
// otherwise trilinear + mip-linear, lambda clamped to the last level.
// Thread mapping: a warp shades an 8x4 pixel tile so that the 32 cones marched in lock-step (same cone index,
// same step) stay spatially coherent in the texture cache; a CTA of 4 warps covers 32x4 pixels.
// Device code: cone_trace.cuh (shared with cone_trace_debug.cu, which holds the debug views in an exactly rounded build).
#include "cone_trace.cuh"

// rows of the image buffer: whole 64-row screen tiles (the sharded trace addresses pixels by tile)
size_t vctk_image_rows(const vct_ctx* c) { return (size_t)((c->H + 63) / 64) * 64; }

int vctk_cone_trace_debug(vct_ctx* c, const TraceArgs& a, int wm, dim3 grid);          // cone_trace_debug.cu

// Screen sharding: 64x64-pixel tiles, tile (tx, ty) belongs to rank (tx + ty) mod N — diagonal stripes, so that every rank
// gets the same mix of sky, floor and walls.  Mirrored by vct_b200/sharded.py::tile_owner for the host-side tests.
constexpr int kScreenTile = 64;
static int ensure_trace_tiles(vct_ctx* c) {
    if (c->d_trace_tiles) return 0;
    std::vector<uint32_t> tiles;
    const int ntx = (c->W + kScreenTile - 1) / kScreenTile, nty = (c->H + kScreenTile - 1) / kScreenTile;
    for (int ty = 0; ty < nty; ++ty)
        for (int tx = 0; tx < ntx; ++tx)
            if ((tx + ty) % c->cfg.world_size == c->cfg.rank) tiles.push_back((uint32_t)(tx * kScreenTile) | (uint32_t)(ty * kScreenTile) << 16);
    c->n_trace_tiles = (int)tiles.size();
    VCT_CHECK(c, cudaMalloc(&c->d_trace_tiles, std::max<size_t>(tiles.size(), 1) * 4));
    if (!tiles.empty()) VCT_CHECK(c, cudaMemcpy(c->d_trace_tiles, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

static TraceArgs make_trace_args(vct_ctx* c) {
    TraceArgs a{};
    a.fc = c->d_fc; a.W = c->W; a.H = c->H;
    a.y_lo = 0; a.y_hi = c->H;                               // the whole image, unless this is a sharded frame (vctk_cone_trace)
    a.vis = c->d_vis; a.indices = c->d_indices; a.trimat = c->d_trimat; a.verts = c->d_vertices;
    a.wpos = c->d_wpos; a.wnrm = c->d_wnrm; a.wT = c->d_wT; a.wB = c->d_wB; a.tex = c->d_tex; a.mats = c->d_mat; a.shadow = c->d_shadow;
    const bool rad = c->h_fc.p.draw_radiance != 0;
    a.vol = rad ? c->radiance_tex : c->color_tex; a.vol_point = rad ? c->radiance_tex_point : c->color_tex_point;
    a.vol_last = rad ? c->radiance_tex_last : c->color_tex_last; a.warp = reinterpret_cast<const float4*>(c->d_warpmap + 4 * (size_t)VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM);
    a.level0 = rad ? c->d_radiance : c->d_color; a.normal0 = c->d_normal;
    a.image = c->image_of(c->image_parity ^ 1); a.counters = c->d_counters;       // the half the last trace did not write
    return a;
}

int vctk_cone_trace(vct_ctx* c) {
    TraceArgs a = make_trace_args(c);
    dim3 grid((c->W + kThreads / 4 - 1) / (kThreads / 4), (c->H + 3) / 4);
    const int par = c->image_parity ^ 1;
    if (vctk_xchg_ready(c)) {                                   // sharded frame: own tiles only, pixels also into rank 0's image
        if (ensure_trace_tiles(c)) return 1;
        if (!c->n_trace_tiles) { c->image_parity = par; return 0; }      // (more ranks than tiles: stay in step with the others)
        a.tiles = c->d_trace_tiles;
        a.image_remote = c->cfg.rank ? reinterpret_cast<uint32_t*>(c->peer[0].image) + (size_t)par * vctk_image_rows(c) * (size_t)c->W : nullptr;
        grid = dim3((unsigned)c->n_trace_tiles * 32u, 1);
    }
    if (c->copy_pending[par]) {                                 // vct_read_image_async: this half is still being read back (the frame before the last)
        VCT_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done[par], 0));
        c->copy_pending[par] = false;
    }
    const vct_frame_params& p = c->h_fc.p;
    // Which mapping the kernel applies.  traceCone tests warpTexture first (phong.frag:150-162), common.glsl's getVoxelPosition —
    // all the voxel view uses (:348) — tests warpVoxels first (common.glsl:44-60); the two differ only when both are switched on.
    const bool voxel_view = p.debug_view == VCT_VIEW_VOXELS || p.debug_view == VCT_VIEW_VOXEL_NORMALS;
    const int wm = voxel_view ? (p.warp_voxels ? WARP_VOXELS : p.warp_texture ? WARP_TEXTURE : p.voxelize_tesselation_warp ? WARP_TESS : WARP_NONE)
                              : (p.warp_texture ? WARP_TEXTURE : p.warp_voxels ? WARP_VOXELS : p.voxelize_tesselation_warp ? WARP_TESS : WARP_NONE);
    if (p.debug_view != VCT_VIEW_SHADED) {                      // debug views: own instantiations in the exactly rounded unit
        if (p.debug_view < 0 || p.debug_view > VCT_VIEW_LAST) { c->error = "vct_cone_trace: unknown debug_view"; return 1; }
        if (p.debug_view == VCT_VIEW_VOXEL_NORMALS && c->cfg.world_size > 1) { c->error = "VCT_VIEW_VOXEL_NORMALS: voxelNormal is not exchanged between the ranks (world_size must be 1)"; return 1; }
        if (vctk_cone_trace_debug(c, a, wm, grid)) return 1;
        c->image_parity = par;
        return 0;
    }
    if (wm == WARP_VOXELS) k_cone_trace<WARP_VOXELS><<<grid, kThreads, 0, c->stream>>>(a);
    else if (wm == WARP_TEXTURE) k_cone_trace<WARP_TEXTURE><<<grid, kThreads, 0, c->stream>>>(a);
    else if (wm == WARP_TESS) k_cone_trace<WARP_TESS><<<grid, kThreads, 0, c->stream>>>(a);
    else k_cone_trace<WARP_NONE><<<grid, kThreads, 0, c->stream>>>(a);
    VCT_LAUNCH_CHECK(c, "k_cone_trace");
    c->image_parity = par;
    return 0;
}
