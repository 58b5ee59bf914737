// raster_passes.cu — depth producers: the light-space shadow map (a0) and the camera visibility buffer (a0').
// Compiled with -fmad=false (coverage and depth are bit-exact vs the oracle).
//
//   shadow map   reference src/Application.cpp:212-233, simple.vert:15-22, reflectiveShadowMap.frag:34-38
//                depth test LESS, back-face culling, depth = ndc.z*0.5+0.5 as float32, clear 1.0
//   visibility   depth prepass + GL_EQUAL colour pass (src/Application.cpp:936-977, dither.frag:25-31):
//                per pixel the nearest alpha-tested fragment, last-drawn wins among equal depths, stored as
//                depthbits<<32 | (0xFFFFFFFF - drawIndex) and resolved with one 64-bit atomicMin
//
// Two kernels per pass.  k_raster_bin: one thread per triangle — transform, (near-)clip, fixed-point setup;
// (sub-)triangles whose bounding box is <= kInlineArea pixels are rasterised on the spot, the others publish their
// setup and are cut into 8x4-pixel tiles (trivially rejected tiles skipped) pushed to a device work queue.
// k_raster_tiles: a persistent grid (CTAs = multiple of 148 SMs) pops tiles, one warp per tile, ONE LANE PER PIXEL.
#include "raster.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kInlineArea = 4;

struct __align__(16) TileSetup { // 96 bytes, written once per queued (sub-)triangle
    TriSetup s;
    uint32_t tri; int alpha_tex; uint32_t pad[2];
};
static_assert(sizeof(TileSetup) == 96, "TileSetup is broadcast-loaded as 6 x 16 bytes");

struct RasterArgs {
    const FrameConst* fc; int W, H;
    const uint32_t* indices; const int32_t* trimat; const float* verts; uint32_t n_tris;
    const float4* wpos; const DevTexture* tex; const DevMaterial* mats;
    unsigned* depth_bits; unsigned long long* vis;
    TileSetup* setups; TileQueues q; unsigned setup_cap;
    Counters* counters; unsigned* setup_count;
};

__device__ __forceinline__ void clip_verts(const RasterArgs& a, bool camera, uint32_t t, RV cv[3]) {
    const FrameConst& fc = *a.fc;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 w = __ldg(a.wpos + __ldg(a.indices + 3 * (size_t)t + k));
        const V4 c = camera ? mul44(fc.projection, mul44(fc.view, mk4(w.x, w.y, w.z, 1.0f))) : mul44(fc.lp, mul44(fc.lv, mk4(w.x, w.y, w.z, 1.0f)));
        cv[k].x = c.x; cv[k].y = c.y; cv[k].z = c.z; cv[k].w = c.w;
    }
}
__device__ __forceinline__ void load_uv(const RasterArgs& a, uint32_t t, float uv[3][2]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t vi = __ldg(a.indices + 3 * (size_t)t + k);
        uv[k][0] = __ldg(a.verts + 14 * (size_t)vi + 6); uv[k][1] = __ldg(a.verts + 14 * (size_t)vi + 7);
    }
}
// perspective-correct barycentrics of the unclipped triangle (2D homogeneous edge functions)
struct Homog { float a[3], b[3], c[3]; };
__device__ __forceinline__ Homog homog_setup(const RV v[3]) {
    Homog h;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const RV p = v[(i + 1) % 3], q = v[(i + 2) % 3];
        h.a[i] = p.y * q.w - q.y * p.w; h.b[i] = q.x * p.w - p.x * q.w; h.c[i] = p.x * q.y - q.x * p.y;
    }
    return h;
}
__device__ __forceinline__ void homog_eval(const Homog& h, float nx, float ny, float l[3]) {
    float b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = (h.a[i] * nx + h.b[i] * ny) + h.c[i];
    const float s = (b[0] + b[1]) + b[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) l[i] = b[i] / s;
}

// alpha test of reflectiveShadowMap.frag:35 / dither.frag:26-31: discard iff alpha.r < 0.1
struct AlphaCtx { bool on; const DevTexture* tex; float uv[3][2]; float rho2; Homog hg; };
template <bool CAMERA>
__device__ __forceinline__ void alpha_setup(const RasterArgs& a, uint32_t t, int alpha_tex, const RV cv_unclipped[3], AlphaCtx& ac) {
    ac.on = alpha_tex >= 0;
    if (!ac.on) return;
    ac.tex = a.tex + alpha_tex;
    load_uv(a, t, ac.uv);
    if (CAMERA) ac.hg = homog_setup(cv_unclipped);
    else ac.rho2 = tri_rho2_affine(cv_unclipped, ac.uv, a.W, a.H, *ac.tex);
}
template <bool CAMERA>
__device__ __forceinline__ bool alpha_pass(const RasterArgs& a, const AlphaCtx& ac, int px, int py, const float l[3]) {
    if (!ac.on) return true;
    if (!CAMERA) {
        const float u = interp1(l, ac.uv[0][0], ac.uv[1][0], ac.uv[2][0]), v = interp1(l, ac.uv[0][1], ac.uv[1][1], ac.uv[2][1]);
        return !(sample2d(*ac.tex, u, v, ac.rho2).x < 0.1f);
    }
    const float nx = ((float)px + 0.5f) / (float)a.W * 2.0f - 1.0f, ny = ((float)py + 0.5f) / (float)a.H * 2.0f - 1.0f;
    float lb[3], lx[3], ly[3];
    homog_eval(ac.hg, nx, ny, lb); homog_eval(ac.hg, nx + 2.0f / (float)a.W, ny, lx); homog_eval(ac.hg, nx, ny + 2.0f / (float)a.H, ly);
    const float u = interp1(lb, ac.uv[0][0], ac.uv[1][0], ac.uv[2][0]), v = interp1(lb, ac.uv[0][1], ac.uv[1][1], ac.uv[2][1]);
    const float ux = interp1(lx, ac.uv[0][0], ac.uv[1][0], ac.uv[2][0]) - u, vx = interp1(lx, ac.uv[0][1], ac.uv[1][1], ac.uv[2][1]) - v;
    const float uy = interp1(ly, ac.uv[0][0], ac.uv[1][0], ac.uv[2][0]) - u, vy = interp1(ly, ac.uv[0][1], ac.uv[1][1], ac.uv[2][1]) - v;
    const float ax = ux * (float)ac.tex->w, bx = vx * (float)ac.tex->h, ay = uy * (float)ac.tex->w, by = vy * (float)ac.tex->h;
    return !(sample2d(*ac.tex, u, v, maxsel(ax * ax + bx * bx, ay * ay + by * by)).x < 0.1f);
}

// out-of-line so that the tile kernel's common path stays light on registers
template <bool CAMERA>
__device__ __noinline__ bool alpha_test_slow(const RasterArgs& a, uint32_t tri, int alpha_tex, int px, int py, const float l[3]) {
    RV cv[3]; clip_verts(a, CAMERA, tri, cv);
    AlphaCtx ac; alpha_setup<CAMERA>(a, tri, alpha_tex, cv, ac);
    return alpha_pass<CAMERA>(a, ac, px, py, l);
}

template <bool CAMERA>
__device__ __forceinline__ void depth_write(const RasterArgs& a, int px, int py, float z, uint32_t t) {
    const float d = z * 0.5f + 0.5f;
    const unsigned db = __float_as_uint(d);
    const size_t o = (size_t)py * a.W + px;
    if (CAMERA) {
        const unsigned long long key = (unsigned long long)db << 32 | (0xFFFFFFFFu - t);
        if (key < a.vis[o]) atomicMin(a.vis + o, key);
    } else {
        if (db < a.depth_bits[o]) atomicMin(a.depth_bits + o, db);
    }
}

template <bool CAMERA>
__global__ void __launch_bounds__(kThreads) k_raster_bin(RasterArgs a) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (a.n_tris + 31u) & ~31u;                       // whole warps stay converged for the votes
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        RV cv[3]; RV sub[2][3]; int nsub = 0; int alpha_tex = -1;
        TriSetup S[2]; bool valid[2] = {false, false}, tiny[2] = {false, false};
        if (t < a.n_tris) {
            clip_verts(a, CAMERA, t, cv);
            if (CAMERA) nsub = clip_near(cv, sub);
            else { nsub = 1; sub[0][0] = cv[0]; sub[0][1] = cv[1]; sub[0][2] = cv[2]; }
            for (int q = 0; q < nsub; ++q) {
                valid[q] = tri_setup(sub[q], a.W, a.H, true, S[q]);
                if (valid[q]) tiny[q] = (S[q].x1 - S[q].x0 + 1) * (S[q].y1 - S[q].y0 + 1) <= kInlineArea;
            }
            if (valid[0] || valid[1]) alpha_tex = a.mats[__ldg(a.trimat + t)].alpha_tex;
        }
        // ---- tiny (sub-)triangles: rasterise now
        if ((valid[0] && tiny[0]) || (valid[1] && tiny[1])) {
            AlphaCtx ac; alpha_setup<CAMERA>(a, t, alpha_tex, cv, ac);
            for (int q = 0; q < nsub; ++q) {
                if (!valid[q] || !tiny[q]) continue;
                const TriSetup& s = S[q];
                for (int py = s.y0; py <= s.y1; ++py)
                    for (int px = s.x0; px <= s.x1; ++px) {
                        float l[3];
                        if (!pixel_owned(a.q, px, py) || !tri_cover(s, px, py, l)) continue;
                        const float z = interp1(l, s.z[0], s.z[1], s.z[2]);
                        if (z < -1.0f || z > 1.0f) continue;
                        if (!alpha_pass<CAMERA>(a, ac, px, py, l)) continue;
                        depth_write<CAMERA>(a, px, py, z, t);
                    }
            }
        }
        // ---- the others: publish the setup, cut the bounding box into tiles
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!CAMERA && q == 1) break;
            const bool queued = valid[q] && !tiny[q];
            const uint32_t slot = reserve_slots(queued, a.setup_count);
            bool stored = false;
            if (queued) {
                if (slot < a.setup_cap) { TileSetup ts; ts.s = S[q]; ts.tri = t; ts.alpha_tex = alpha_tex; ts.pad[0] = ts.pad[1] = 0u; a.setups[slot] = ts; stored = true; }
                else vct_flag_overflow(a.counters);
            }
            enqueue_tiles(stored, S[q], slot, a.q);
        }
    }
}

template <bool CAMERA>
__global__ void __launch_bounds__(kThreads) k_raster_tiles(RasterArgs a) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = min(*a.q.tile_count, a.q.tile_cap);
    const unsigned warps = gridDim.x * (kThreads / 32);
    for (unsigned item = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); item < n_items; item += warps) {
        const uint2 it = __ldg(a.q.tiles + item);
        const TileSetup ts = a.setups[it.x];                                // same address in every lane: broadcast
        const TriSetup& s = ts.s;
        const int px = (int)(it.y & 0xFFFFu) + (lane & (kTileW - 1)), py = (int)(it.y >> 16) + (lane >> 3);
        float l[3];
        if (px > s.x1 || py > s.y1 || !pixel_owned(a.q, px, py) || !tri_cover(s, px, py, l)) continue;
        const float z = interp1(l, s.z[0], s.z[1], s.z[2]);
        if (z < -1.0f || z > 1.0f) continue;
        if (ts.alpha_tex >= 0 && !alpha_test_slow<CAMERA>(a, ts.tri, ts.alpha_tex, px, py, l)) continue;   // rare: masked materials only
        depth_write<CAMERA>(a, px, py, z, ts.tri);
    }
}
__global__ void __launch_bounds__(kThreads) k_raster_expand(const TileSetup* __restrict__ setups, TileQueues q) {
    expand_items(reinterpret_cast<const unsigned char*>(setups), sizeof(TileSetup), q);
}

__global__ void __launch_bounds__(kThreads) k_fill_u32(unsigned* __restrict__ p, size_t n, unsigned v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(kThreads) k_fill_u64(unsigned long long* __restrict__ p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_reset_queue(Counters* c) { c->tile_queue_count = 0; c->setup_count = 0; c->expand_count = 0; c->pixel_count = 0; }

template <bool CAMERA>
int run_raster(vct_ctx* c) {
    RasterArgs a{};
    a.fc = c->d_fc; a.W = CAMERA ? c->W : c->S; a.H = CAMERA ? c->H : c->S;
    a.indices = c->d_indices; a.trimat = c->d_trimat; a.verts = c->d_vertices; a.n_tris = (uint32_t)c->n_tris;
    a.wpos = c->d_wpos; a.tex = c->d_tex; a.mats = c->d_mat;
    a.depth_bits = reinterpret_cast<unsigned*>(c->d_shadow); a.vis = c->d_vis;
    a.setups = reinterpret_cast<TileSetup*>(c->d_setup); a.q = vctk_tile_queues(c);
    a.setup_cap = (unsigned)c->setup_cap; a.counters = c->d_counters;
    a.setup_count = &c->d_counters->setup_count;
    // sharded frame (peers attached): a rank's cone trace reads the visibility of its own screen tiles only; the shadow map is
    // rasterised in bands of rows, one per rank, into the map the last frame did not use, and the bands are exchanged (exchange.cu)
    const bool sharded = vctk_xchg_ready(c);
    int row_lo = 0, row_hi = c->S;
    if (sharded) {
        a.q.own_world = c->cfg.world_size; a.q.own_rank = c->cfg.rank;
        if (!CAMERA) {
            c->shadow_parity ^= 1; c->d_shadow = c->d_shadow_base + (size_t)c->shadow_parity * c->S * c->S;
            a.depth_bits = reinterpret_cast<unsigned*>(c->d_shadow);
            a.q.own_band = (c->S + c->cfg.world_size - 1) / c->cfg.world_size;
            row_lo = std::min(c->S, c->cfg.rank * a.q.own_band); row_hi = std::min(c->S, row_lo + a.q.own_band);
        }
    }
    const size_t npx = CAMERA ? (size_t)a.W * a.H : (size_t)(row_hi - row_lo) * c->S;
    const int fill_grid = (int)std::min<size_t>((npx + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 32);
    if (CAMERA) k_fill_u64<<<fill_grid, kThreads, 0, c->stream>>>(c->d_vis, npx, ~0ull);
    else k_fill_u32<<<std::max(fill_grid, 1), kThreads, 0, c->stream>>>(a.depth_bits + (size_t)row_lo * c->S, npx, 0x3F800000u);
    VCT_LAUNCH_CHECK(c, CAMERA ? "k_fill_u64" : "k_fill_u32");
    k_reset_queue<<<1, 1, 0, c->stream>>>(c->d_counters); VCT_LAUNCH_CHECK(c, "k_reset_queue");
    if (!c->n_tris) return (!CAMERA && sharded) ? vctk_xchg_shadow(c, row_lo, row_hi) : 0;
    const int grid = (int)std::min<size_t>((c->n_tris + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_raster_bin<CAMERA><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, CAMERA ? "k_raster_bin_camera" : "k_raster_bin_light");
    k_raster_expand<<<VCT_SM_COUNT * 4, kThreads, 0, c->stream>>>(a.setups, a.q); VCT_LAUNCH_CHECK(c, CAMERA ? "k_raster_expand_camera" : "k_raster_expand_light");
    k_raster_tiles<CAMERA><<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, CAMERA ? "k_raster_tiles_camera" : "k_raster_tiles_light");
    if (!CAMERA && sharded) return vctk_xchg_shadow(c, row_lo, row_hi);
    return 0;
}

}  // namespace

size_t vctk_tile_setup_bytes() { return sizeof(TileSetup); }
int vctk_shadowmap(vct_ctx* c) { return run_raster<false>(c); }
int vctk_visibility(vct_ctx* c) { return run_raster<true>(c); }
