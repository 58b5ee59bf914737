// debug_voxels.cu — Application::debugVoxels (reference src/Application.cpp:1222-1275, shaders/debugVoxels.vert/.geom/.frag):
// the non-empty voxels of the base grid drawn as cubes.  The reference issues glDrawArraysInstanced(GL_POINTS, 0, 1, D^3): the vertex shader turns
// gl_InstanceID into a voxel (two modf calls on float(gl_InstanceID) — beyond 2^24 instances ids collapse, a quirk that is kept) and samples the
// pyramid at the voxel's centre with lod = Settings::miplevel; the geometry shader emits a 21-vertex triangle strip (19 triangles, the repeated
// indices make degenerate stitches) for voxels with alpha > 0; depth test GL_LESS, back faces culled, colour = the voxel's.
//   k_dbgvox_list     one thread per instance: voxel colour, alpha test, warp-aggregated append
//   k_dbgvox_raster   one thread per (listed voxel, strip triangle): clip-space vertices exactly as the two shaders compute them, near-plane clip,
//                     fixed-point set-up with back-face culling (raster.cuh: the rasteriser of every other pass), pixels of small triangles at
//                     once, large ones queued
//   k_dbgvox_raster_big  one CTA per queued triangle
//   k_dbgvox_resolve  depth-tested winner per pixel -> colour.  Depth test = atomicMin on (bits of the fp32 window depth) << 32 | draw order:
//                     among equal depths the first drawn stays, like GL_LESS with in-order primitives.
// Compiled with -fmad=false like every unit that decides where a value lands; bit-exact against the tests' CPU restatement of the same passes at
// lod <= 0.5, tolerance-gated above it (the texture unit's 8-bit filter weights).
#include "raster.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kStripVerts = 21, kStripTris = 19;
__constant__ float c_cube[8][3] = {{-0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, 0.5f}, {-0.5f, 0.5f, 0.5f},
                                   {-0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, 0.5f}, {-0.5f, -0.5f, 0.5f}};   // debugVoxels.geom:16-26
__constant__ int c_strip[kStripVerts] = {5, 4, 1, 0, 0, 0, 0, 3, 1, 2, 5, 6, 4, 7, 0, 3, 3, 3, 2, 7, 6};                      // :27-33

struct DbgArgs {
    const FrameConst* fc; Mat4 mvp; int D, L, W, H; float lod;
    cudaTextureObject_t vol; const uint32_t* level0;
    unsigned long long* vis; uint32_t* image; uint32_t clear;
    uint32_t* list; unsigned* count; unsigned cap;                 // instance ids of the voxels that passed the alpha test
    uint4* big; unsigned* big_count; unsigned big_cap;             // (list entry, strip triangle, clipped part, 0) of triangles with a large pixel box
    Counters* counters;
};

// debugVoxels.vert:19-30
__device__ __forceinline__ void instance_voxel(const vct_frame_params& fp, int D, uint32_t id, V3& tc, V3& world) {
    const float dim = (float)D;
    float instance = (float)id;
    float t = instance / dim; instance = truncf(t); const float x = t - instance;         // modf(instance / voxelDim, instance)
    t = instance / dim; instance = truncf(t); const float y = t - instance;
    const float z = instance / dim;
    const float h = 0.5f / dim;
    tc = mk3(x + h, y + h, z + h);
    world = mk3((0.0f + fp.voxel_center[0]) + (fp.voxel_min[0] * (1.0f - tc.x) + fp.voxel_max[0] * tc.x),
                (0.0f + fp.voxel_center[1]) + (fp.voxel_min[1] * (1.0f - tc.y) + fp.voxel_max[1] * tc.y),
                (0.0f + fp.voxel_center[2]) + (fp.voxel_min[2] * (1.0f - tc.z) + fp.voxel_max[2] * tc.z));
}
// textureLod(voxels, tc, level): lambda <= 0.5 magnifies -> NEAREST on level 0, read from the linear level by floor index (exact); else the texture unit
__device__ __forceinline__ V4 voxel_color(const DbgArgs& a, V3 tc) {
    if (!(a.lod > 0.5f)) {
        const float Df = (float)a.D;
        const float fx = floorf(tc.x * Df), fy = floorf(tc.y * Df), fz = floorf(tc.z * Df);
        if (!(fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < Df && fy < Df && fz < Df)) return mk4(0.f, 0.f, 0.f, 0.f);
        return unpack_unorm(__ldg(a.level0 + ((size_t)(int)fz * a.D + (int)fy) * a.D + (int)fx));
    }
    const float4 s = tex3DLod<float4>(a.vol, tc.x, tc.y, tc.z, fminf(a.lod, (float)(a.L - 1)));
    return mk4(s.x, s.y, s.z, s.w);
}

__global__ void __launch_bounds__(kThreads) k_dbgvox_clear(unsigned long long* vis, size_t n, unsigned* count, unsigned* big_count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) vis[i] = ~0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *count = 0u; *big_count = 0u; }
}
__global__ void __launch_bounds__(kThreads) k_dbgvox_list(DbgArgs a) {
    const vct_frame_params& fp = a.fc->p;
    const size_t n = (size_t)a.D * a.D * a.D;
    const int lane = threadIdx.x & 31;
    const size_t n_round = (n + 31) & ~(size_t)31;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += (size_t)gridDim.x * blockDim.x) {
        bool keep = false;
        if (i < n) { V3 tc, world; instance_voxel(fp, a.D, (uint32_t)i, tc, world); keep = voxel_color(a, tc).w > 0.0f; }   // debugVoxels.geom:46
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (!m) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(a.count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) { const unsigned pos = base + __popc(m & ((1u << lane) - 1u)); if (pos < a.cap) a.list[pos] = (uint32_t)i; else vct_flag_overflow(a.counters); }
    }
}
// clip-space vertices of strip triangle k of instance id (odd triangles: winding reversed, OpenGL 4.5 section 10.1.8)
__device__ __forceinline__ void strip_triangle(const DbgArgs& a, uint32_t id, int k, RV cv[3]) {
    const vct_frame_params& fp = a.fc->p;
    V3 tc, world; instance_voxel(fp, a.D, id, tc, world);
    const float dim = (float)a.D;
    const V3 size = mk3((fp.voxel_max[0] - fp.voxel_min[0]) / dim, (fp.voxel_max[1] - fp.voxel_min[1]) / dim, (fp.voxel_max[2] - fp.voxel_min[2]) / dim);   // voxelWorldSize()
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float* q = c_cube[c_strip[k + j]];
        const V4 p = mul44(a.mvp, mk4(size.x * q[0] + world.x, size.y * q[1] + world.y, size.z * q[2] + world.z, 1.0f));   // debugVoxels.geom:51-52
        cv[j].x = p.x; cv[j].y = p.y; cv[j].z = p.z; cv[j].w = p.w;
    }
    if (k & 1) { const RV t = cv[0]; cv[0] = cv[1]; cv[1] = t; }
}
__device__ __forceinline__ void shade_pixel(const DbgArgs& a, const TriSetup& s, int px, int py, uint32_t order) {
    float l[3];
    if (!tri_cover(s, px, py, l)) return;
    const float zn = (l[0] * s.z[0] + l[1] * s.z[1]) + l[2] * s.z[2];
    if (zn > 1.0f) return;                                                               // far plane (the near plane was clipped in clip space)
    const float dw = zn * 0.5f + 0.5f;
    atomicMin(a.vis + (size_t)py * a.W + px, (unsigned long long)__float_as_uint(dw) << 32 | order);   // GL_LESS, primitives in order
}
constexpr int kSmallBox = 1024;
__global__ void __launch_bounds__(kThreads) k_dbgvox_raster(DbgArgs a) {
    const unsigned n = min(*a.count, a.cap);
    const size_t items = (size_t)n * kStripTris;
    for (size_t it = blockIdx.x * (size_t)blockDim.x + threadIdx.x; it < items; it += (size_t)gridDim.x * blockDim.x) {
        const unsigned e = (unsigned)(it / kStripTris); const int k = (int)(it - (size_t)e * kStripTris);
        if (c_strip[k] == c_strip[k + 1] || c_strip[k + 1] == c_strip[k + 2] || c_strip[k] == c_strip[k + 2]) continue;   // degenerate stitch
        const uint32_t id = a.list[e];
        RV cv[3]; strip_triangle(a, id, k, cv);
        RV sub[2][3]; const int nsub = clip_near(cv, sub);
        for (int q = 0; q < nsub; ++q) {
            TriSetup s;
            if (!tri_setup(sub[q], a.W, a.H, true, s)) continue;                         // GL_CULL_FACE, back faces (CCW is front)
            if ((s.x1 - s.x0 + 1) * (s.y1 - s.y0 + 1) > kSmallBox) {
                const unsigned pos = atomicAdd(a.big_count, 1u);
                if (pos < a.big_cap) a.big[pos] = make_uint4(e, (unsigned)k, (unsigned)q, 0u); else vct_flag_overflow(a.counters);
                continue;
            }
            const uint32_t order = id * (uint32_t)kStripTris + (uint32_t)k;
            for (int py = s.y0; py <= s.y1; ++py) for (int px = s.x0; px <= s.x1; ++px) shade_pixel(a, s, px, py, order);
        }
    }
}
__global__ void __launch_bounds__(kThreads) k_dbgvox_raster_big(DbgArgs a) {
    const unsigned n = min(*a.big_count, a.big_cap);
    for (unsigned b = blockIdx.x; b < n; b += gridDim.x) {
        const uint4 it = a.big[b];
        const uint32_t id = a.list[it.x];
        RV cv[3]; strip_triangle(a, id, (int)it.y, cv);
        RV sub[2][3]; const int nsub = clip_near(cv, sub);
        if ((int)it.z >= nsub) continue;
        TriSetup s;
        if (!tri_setup(sub[it.z], a.W, a.H, true, s)) continue;
        const int bw = s.x1 - s.x0 + 1, total = bw * (s.y1 - s.y0 + 1);
        const uint32_t order = id * (uint32_t)kStripTris + it.y;
        for (int i = threadIdx.x; i < total; i += kThreads) shade_pixel(a, s, s.x0 + i % bw, s.y0 + i / bw, order);
    }
}
__global__ void __launch_bounds__(kThreads) k_dbgvox_resolve(DbgArgs a) {
    const vct_frame_params& fp = a.fc->p;
    const size_t n = (size_t)a.W * a.H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long key = a.vis[i];
        uint32_t word = a.clear;                                                          // glClear(GL_COLOR_BUFFER_BIT)
        if (key != ~0ull) {
            const uint32_t id = (uint32_t)(key & 0xFFFFFFFFu) / (uint32_t)kStripTris;
            V3 tc, world; instance_voxel(fp, a.D, id, tc, world);
            const V4 c = voxel_color(a, tc);
            word = pack_unorm(mk4(c.x, c.y, c.z, 1.0f));                                  // debugVoxels.frag:10 (alpha 1 like every image of this library)
        }
        a.image[i] = word;
    }
}
}  // namespace

// mvp = projection * view, the product the host forms with GLM (Application.cpp:926): column c = ((P0 V[c][0] + P1 V[c][1]) + P2 V[c][2]) + P3 V[c][3]
static Mat4 glm_product(const float* P, const float* V) {
    Mat4 m;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            m.m[4 * c + r] = ((P[r] * V[4 * c] + P[4 + r] * V[4 * c + 1]) + P[8 + r] * V[4 * c + 2]) + P[12 + r] * V[4 * c + 3];
    return m;
}

int vctk_debug_voxels(vct_ctx* c) {
    const vct_frame_params& p = c->h_fc.p;
    if (c->cfg.world_size > 1) { c->error = "vct_debug_voxels: one GPU only (the image is assembled from screen tiles only for the shaded frame)"; return 1; }
    if ((size_t)c->D * c->D * c->D * kStripTris > 0xFFFFFFFFull) { c->error = "vct_debug_voxels: dim too large for the 32-bit draw order (dim <= 512)"; return 1; }
    const size_t npx = (size_t)c->W * c->H, nvox = (size_t)c->D * c->D * c->D;
    if (!c->d_dbg) {                                               // vis (W H u64) | counters (64 B) | list (nvox u32) | big (1 Mi uint4)
        VCT_CHECK(c, cudaMalloc(&c->d_dbg, npx * 8 + 64 + nvox * 4 + ((size_t)1 << 20) * 16));
    }
    const bool rad = p.draw_radiance != 0;
    DbgArgs a{};
    a.fc = c->d_fc; a.mvp = glm_product(p.projection, p.view); a.D = c->D; a.L = c->L; a.W = c->W; a.H = c->H; a.lod = p.miplevel;
    a.vol = rad ? c->radiance_tex : c->color_tex; a.level0 = rad ? c->d_radiance : c->d_color;
    if (!a.vol) { c->error = "vct_debug_voxels: the pyramid has no texture array (run a frame first)"; return 1; }
    char* base = reinterpret_cast<char*>(c->d_dbg);
    a.vis = reinterpret_cast<unsigned long long*>(base);
    a.count = reinterpret_cast<unsigned*>(base + npx * 8); a.big_count = a.count + 1;
    a.list = reinterpret_cast<uint32_t*>(base + npx * 8 + 64); a.cap = (unsigned)std::min<size_t>(nvox, 0xFFFFFFFFu);
    a.big = reinterpret_cast<uint4*>(base + npx * 8 + 64 + nvox * 4); a.big_cap = 1u << 20;
    a.counters = c->d_counters;
    const float cc[3] = {p.clear_color[0], p.clear_color[1], p.clear_color[2]};
    auto u8 = [](float v) { return !(v > 0.0f) ? 0u : (v > 1.0f ? 255u : (unsigned)lrintf(v * 255.0f)); };
    a.clear = u8(cc[0]) | u8(cc[1]) << 8 | u8(cc[2]) << 16 | 255u << 24;
    const int par = c->image_parity ^ 1;
    if (c->copy_pending[par]) { VCT_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done[par], 0)); c->copy_pending[par] = false; }
    a.image = c->image_of(par);
    const unsigned wide = VCT_SM_COUNT * 16;
    k_dbgvox_clear<<<wide, kThreads, 0, c->stream>>>(a.vis, npx, a.count, a.big_count); VCT_LAUNCH_CHECK(c, "k_dbgvox_clear");
    k_dbgvox_list<<<wide, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_dbgvox_list");
    k_dbgvox_raster<<<wide, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_dbgvox_raster");
    k_dbgvox_raster_big<<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_dbgvox_raster_big");
    k_dbgvox_resolve<<<wide, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_dbgvox_resolve");
    c->image_parity = par;
    return 0;
}
