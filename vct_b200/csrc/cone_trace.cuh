// cone_trace.cuh — per-pixel shading + voxel cone tracing: device code shared by two translation units.
//   cone_trace.cu        the shaded frame (debug_view == 0): compiled with -use_fast_math / FMA contraction; tolerance-gated (PSNR)
//   cone_trace_debug.cu  the debug views of phong.frag (DBG instantiations): compiled like the "exact" units (-fmad=false, IEEE
//                        division), because the voxel view at lod <= 0.5 is a NEAREST fetch whose voxel index must equal the
//                        oracle's bit for bit — in the tessellation warp whole image rows sit exactly on voxel boundaries
//                        (profiles/r02a_diag_voxel_view.txt), where one ulp of contraction flips the voxel.
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
#pragma once
#include <algorithm>

#include "common.cuh"

struct TraceArgs {
    const FrameConst* fc; int W, H, y_lo, y_hi;
    const unsigned long long* vis;
    const uint32_t* indices; const int32_t* trimat; const float* verts;
    const float4 *wpos, *wnrm, *wT, *wB;
    const DevTexture* tex; const DevMaterial* mats; const float* shadow;
    cudaTextureObject_t vol, vol_point, vol_last; const float4* warp;     // warp map as floats (k_warpmap_floats)
    const uint32_t* level0;                      // linear level 0 of the traced pyramid: the voxel view's NEAREST fetch reads it exactly
    const uint32_t* normal0;                     // voxelNormal (VCT_VIEW_VOXEL_NORMALS)
    uint32_t* image; Counters* counters;
    // sharded frame (multi-GPU with attached peers): the CTAs walk this rank's 64x64 screen tiles (x0 | y0 << 16; 32 CTAs each) instead of the
    // whole image, and every pixel is also stored into rank 0's image over NVLink (nullptr on rank 0 and on a single GPU)
    const uint32_t* tiles; uint32_t* image_remote;
};

namespace {

constexpr int kThreads = 128;          // default CTA: 4 warps = 32x4 pixels; 128 registers/thread -> 16 warps per SM
__device__ const float kPI = 3.1415982f;                 // common.glsl:1 [sic]


__device__ __forceinline__ V3 f4to3(float4 q) { return mk3(q.x, q.y, q.z); }
__device__ __forceinline__ void put_pixel(const TraceArgs& a, size_t o, uint32_t word) { a.image[o] = word; if (a.image_remote) a.image_remote[o] = word; }
__device__ __forceinline__ float ip(const float l[3], float a, float b, float c) { return l[0] * a + l[1] * b + l[2] * c; }
__device__ __forceinline__ V3 ip3(const float l[3], V3 a, V3 b, V3 c) { return mk3(ip(l, a.x, b.x, c.x), ip(l, a.y, b.y, c.y), ip(l, a.z, b.z, c.z)); }

// ---- 2D material textures: LINEAR_MIPMAP_NEAREST / NEAREST / REPEAT, software-filtered from linear memory
__device__ __forceinline__ int wrapi(int i, int n) {
    if ((n & (n - 1)) == 0) return i & (n - 1);                            // power-of-two sizes (every Sponza map): REPEAT is a mask
    int r = i % n; return r < 0 ? r + n : r;
}
__device__ __forceinline__ V4 texel2d(const DevTexture& t, int level, int x, int y) {
    const int w = max(1, t.w >> level), h = max(1, t.h >> level);
    const uint8_t* p = t.level[level] + ((size_t)wrapi(y, h) * w + wrapi(x, w)) * t.ch;
    const float k = 1.0f / 255.0f;
    if (t.ch == 4) { const uchar4 q = *reinterpret_cast<const uchar4*>(p); return mk4(q.x * k, q.y * k, q.z * k, q.w * k); }
    if (t.ch == 3) return mk4(p[0] * k, p[1] * k, p[2] * k, 1.0f);
    return mk4(p[0] * k, 0.f, 0.f, 1.f);
}
__device__ __forceinline__ V4 lerp4(V4 a, V4 b, float t) { return mk4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t); }
__device__ __forceinline__ V4 sample2d(const DevTexture& t, float u, float v, float rho2) {
    if (!(rho2 > 2.0f)) return texel2d(t, 0, (int)floorf(u * (float)t.w), (int)floorf(v * (float)t.h));
    int d = 1; float lim = 8.0f;
    while (d < t.levels - 1 && rho2 > lim) { d++; lim *= 4.0f; }
    if (d > t.levels - 1) d = t.levels - 1;
    const int w = max(1, t.w >> d), h = max(1, t.h >> d);
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const int x0 = (int)fx0, y0 = (int)fy0; const float fx = x - fx0, fy = y - fy0;
    return lerp4(lerp4(texel2d(t, d, x0, y0), texel2d(t, d, x0 + 1, y0), fx), lerp4(texel2d(t, d, x0, y0 + 1), texel2d(t, d, x0 + 1, y0 + 1), fx), fy);
}

// ---- shadow map PCF (phong.frag:183-207), LINEAR + CLAMP_TO_BORDER(1)
__device__ __forceinline__ float sm_texel(const float* __restrict__ sm, int S, int x, int y) { return (x < 0 || y < 0 || x >= S || y >= S) ? 1.0f : __ldg(sm + (size_t)y * S + x); }
__device__ __forceinline__ float calc_shadow_factor(const float* __restrict__ sm, int S, V4 lsp) {
    const float sx = (lsp.x / lsp.w + 1.0f) * 0.5f, sy = (lsp.y / lsp.w + 1.0f) * 0.5f, sz = (lsp.z / lsp.w + 1.0f) * 0.5f;
    const float frag_depth = sz - 0.01f;
    if (frag_depth > 1.0f) return 0.0f;
    const float x = sx * (float)S - 0.5f, y = sy * (float)S - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f)) return 0.0f;       // every tap reads the border (1.0): never shadowed
    const int x0 = (int)fx0, y0 = (int)fy0; const float fx = x - fx0, fy = y - fy0;
    // the five bilinear footprints share a 4x4 texel neighbourhood: fetch the 12 distinct texels once
    float t[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) t[j][i] = ((i == 0 || i == 3) && (j == 0 || j == 3)) ? 0.0f : sm_texel(sm, S, x0 - 1 + i, y0 - 1 + j);
    auto bil = [&](int ox, int oy) {
        const int i = 1 + ox, j = 1 + oy;
        const float top = t[j][i] * (1.0f - fx) + t[j][i + 1] * fx, bot = t[j + 1][i] * (1.0f - fx) + t[j + 1][i + 1] * fx;
        return top * (1.0f - fy) + bot * fy;
    };
    float f = 0.0f;
    if (frag_depth > bil(0, 0)) f += 1.0f;
    if (frag_depth > bil(1, 0)) f += 1.0f;
    if (frag_depth > bil(0, 1)) f += 1.0f;
    if (frag_depth > bil(-1, 0)) f += 1.0f;
    if (frag_depth > bil(0, -1)) f += 1.0f;
    return f / 5.0f;
}

// ---- common.glsl
__device__ __forceinline__ V3 voxel_linear_position(V3 p, const vct_frame_params& fp) {
    return mk3((p.x - fp.voxel_center[0] - fp.voxel_min[0]) / (fp.voxel_max[0] - fp.voxel_min[0]),
               (p.y - fp.voxel_center[1] - fp.voxel_min[1]) / (fp.voxel_max[1] - fp.voxel_min[1]),
               (p.z - fp.voxel_center[2] - fp.voxel_min[2]) / (fp.voxel_max[2] - fp.voxel_min[2]));
}
// common.glsl:37-42 (voxelizeTesselationWarp): (pv * P).xyz / w * 0.5 + 0.5
__device__ __forceinline__ V3 tess_warp_position(V3 pos, const vct_frame_params& fp) {
    const float* m = fp.pv;
    const float qx = ((m[0] * pos.x + m[4] * pos.y) + m[8] * pos.z) + m[12];
    const float qy = ((m[1] * pos.x + m[5] * pos.y) + m[9] * pos.z) + m[13];
    const float qz = ((m[2] * pos.x + m[6] * pos.y) + m[10] * pos.z) + m[14];
    const float qw = ((m[3] * pos.x + m[7] * pos.y) + m[11] * pos.z) + m[15];
    return mk3((qx / qw) * 0.5f + 0.5f, (qy / qw) * 0.5f + 0.5f, (qz / qw) * 0.5f + 0.5f);
}
__device__ __forceinline__ float voxel_warp_fn1(float x) {
    const float alpha = 0.25f;
    x = alpha * x + (3.0f - 3.0f * alpha) * x * x + (2.0f * alpha - 2.0f) * x * x * x;
    return clampf(x, 0.0f, 1.0f);
}
__device__ __forceinline__ V3 voxel_warp(V3 p, V3 c) {
    V3 o = p - c;
    o = mk3(voxel_warp_fn1(0.5f * o.x + 0.5f), voxel_warp_fn1(0.5f * o.y + 0.5f), voxel_warp_fn1(0.5f * o.z + 0.5f));
    return c + mk3(2.0f * o.x - 1.0f, 2.0f * o.y - 1.0f, 2.0f * o.z - 1.0f);
}

// ---- warp map lookup: 32^3 RGBA16 unorm, LINEAR, CLAMP_TO_EDGE (reference src/Application.cpp:383-389).
// Filtered in software with fp32 weights from the 256 KiB linear copy (L1/L2 resident): the texture unit's 8-bit
// interpolation weights move the warped sample position by up to 1/256 of a warp cell, which is enough to
// flip the NEAREST (lambda <= 0.5) fetches of the specular cone onto a neighbouring voxel (measured: final
// image PSNR 41 dB with the hardware filter vs the fp32 definition of the oracle).
// One 256-bit load (LDG.E.256, sm_100) fetches texel (x, y, z) and its +x neighbour: the table holds them side by side (k_warpmap_floats), q / 65535
// per channel divided once there.  x, y, z already clamped.
__device__ __forceinline__ void warp_texel_pair(const float4* __restrict__ wm, int x, int y, int z, V3& a, V3& b) {
    const int n = VCT_WARP_DIM;
    float ax, ay, az, aw, bx, by, bz, bw;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(ax), "=f"(ay), "=f"(az), "=f"(aw), "=f"(bx), "=f"(by), "=f"(bz), "=f"(bw) : "l"(wm + 2 * ((z * n + y) * n + x)));
    a = mk3(ax, ay, az); b = mk3(bx, by, bz);
}
__device__ __forceinline__ V3 lerp3x(V3 a, V3 b, float t) {
    const float s = 1.0f - t;
    return mk3(__fadd_rn(__fmul_rn(a.x, s), __fmul_rn(b.x, t)), __fadd_rn(__fmul_rn(a.y, s), __fmul_rn(b.y, t)), __fadd_rn(__fmul_rn(a.z, s), __fmul_rn(b.z, t)));
}
__device__ __noinline__ V3 warp_sample(const float4* __restrict__ wm, V3 tc) {
    const float n = (float)VCT_WARP_DIM;
    const float x = __fmul_rn(tc.x, n) - 0.5f, y = __fmul_rn(tc.y, n) - 0.5f, z = __fmul_rn(tc.z, n) - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    const int n1 = VCT_WARP_DIM - 1;
    const int xc = min(max(x0, 0), n1), ya = min(max(y0, 0), n1), yb = min(max(y0 + 1, 0), n1), za = min(max(z0, 0), n1), zb = min(max(z0 + 1, 0), n1);
    V3 a00, b00, a10, b10, a01, b01, a11, b11;
    warp_texel_pair(wm, xc, ya, za, a00, b00); warp_texel_pair(wm, xc, yb, za, a10, b10);
    warp_texel_pair(wm, xc, ya, zb, a01, b01); warp_texel_pair(wm, xc, yb, zb, a11, b11);
    if (x0 < 0) { b00 = a00; b10 = a10; b01 = a01; b11 = a11; }                // x0 = -1: both corners clamp to texel 0
    const V3 c00 = lerp3x(a00, b00, fx), c10 = lerp3x(a10, b10, fx), c01 = lerp3x(a01, b01, fx), c11 = lerp3x(a11, b11, fx);
    return lerp3x(lerp3x(c00, c10, fy), lerp3x(c01, c11, fy), fz);
}

// The same filter with eight plain 16-byte loads (the first half of each pair entry), for the debug view that shows the warp map itself (nvcc 12.9
// crashes when the out-of-line function above gets a second call site in the debug instantiations)
__device__ __forceinline__ V3 warp_sample_plain(const float4* __restrict__ wm, V3 tc) {
    const int n = VCT_WARP_DIM;
    const float x = __fmul_rn(tc.x, (float)n) - 0.5f, y = __fmul_rn(tc.y, (float)n) - 0.5f, z = __fmul_rn(tc.z, (float)n) - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f) || !(fabsf(fz0) < 1e9f)) return mk3(0.f, 0.f, 0.f);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    auto texel = [&](int xx, int yy, int zz) {
        xx = min(max(xx, 0), n - 1); yy = min(max(yy, 0), n - 1); zz = min(max(zz, 0), n - 1);
        const float4 q = __ldg(wm + 2 * ((zz * n + yy) * n + xx));
        return mk3(q.x, q.y, q.z);
    };
    const V3 c00 = lerp3x(texel(x0, y0, z0), texel(x0 + 1, y0, z0), fx), c10 = lerp3x(texel(x0, y0 + 1, z0), texel(x0 + 1, y0 + 1, z0), fx);
    const V3 c01 = lerp3x(texel(x0, y0, z0 + 1), texel(x0 + 1, y0, z0 + 1), fx), c11 = lerp3x(texel(x0, y0 + 1, z0 + 1), texel(x0 + 1, y0 + 1, z0 + 1), fx);
    return lerp3x(lerp3x(c00, c10, fy), lerp3x(c01, c11, fy), fz);
}
// Mapping applied to every cone sample, phong.frag:150-162 (priority INSIDE traceCone: warpTexture > warpVoxels >
// voxelizeTesselationWarp; common.glsl:44-60, which the voxel view and the other passes use, has warpVoxels first)
enum { WARP_NONE = 0, WARP_TEXTURE = 1, WARP_VOXELS = 2, WARP_TESS = 3 };

// ---- traceCone, phong.frag:135-180
// (warp map and frame parameters share a slot: WARP_TESS needs pv and the volume extents, never the warp map — the context of
// the other instantiations keeps its layout)
struct ConeCtx { cudaTextureObject_t vol, vol_point, vol_last; union { const float4* warp; const vct_frame_params* fp; }; int D, L; int warp_texture, warp_voxels; V3 eye_tc; };
// phong.frag:158-162 (voxelizeTesselationWarp): the sample goes back to world space and through pv (common.glsl:37-42)
__device__ __forceinline__ V3 tess_warp_sample(const vct_frame_params& fp, V3 sp) {
    const V3 world = mk3((sp.x * (fp.voxel_max[0] - fp.voxel_min[0]) + fp.voxel_center[0]) + fp.voxel_min[0],
                         (sp.y * (fp.voxel_max[1] - fp.voxel_min[1]) + fp.voxel_center[1]) + fp.voxel_min[1],
                         (sp.z * (fp.voxel_max[2] - fp.voxel_min[2]) + fp.voxel_center[2]) + fp.voxel_min[2]);
    return tess_warp_position(world, fp);
}

template <int WM>
__device__ __noinline__ V4 trace_cone(const ConeCtx& cx, V3 position, V3 normal, V3 direction, int steps, float bias, float cone_angle,
                                         float cone_height, float lod_offset, unsigned& fetches) {
    direction = normalize3(direction);
    V3 color = mk3(0.f, 0.f, 0.f); float alpha = 0.0f;
    const float scale = 1.0f / (float)cx.D;
    const V3 start = position + (normal * bias) * scale;
    const float tan_half = tanf(cone_angle / 2.0f);
    const float max_lod = (float)(cx.L - 1);
    for (int i = 0; i < steps && alpha < 0.95f; ++i) {
        const float cone_radius = cone_height * tan_half;
        const float lod = log2f(fmaxf(1.0f, 2.0f * cone_radius));
        V3 sp = start + (direction * cone_height) * scale;
        if (!(sp.x >= 0.0f && sp.x <= 1.0f && sp.y >= 0.0f && sp.y <= 1.0f && sp.z >= 0.0f && sp.z <= 1.0f)) break;   // also NaN
        if (WM == WARP_TEXTURE) sp = warp_sample(cx.warp, sp);
        else if (WM == WARP_VOXELS) sp = voxel_warp(sp, cx.eye_tc);
        else if (WM == WARP_TESS) sp = tess_warp_sample(*cx.fp, sp);
        const float lambda = lod + lod_offset;
        float4 sc;
        if (!(lambda > 0.5f)) sc = tex3DLod<float4>(cx.vol_point, sp.x, sp.y, sp.z, 0.0f);
        else sc = tex3DLod<float4>(cx.vol, sp.x, sp.y, sp.z, fminf(lambda, max_lod));
        fetches++;
        const float a = 1.0f - alpha;
        color = color + mk3(sc.x, sc.y, sc.z) * a;
        alpha += a * sc.w;
        cone_height += cone_radius;
    }
    return mk4(color.x, color.y, color.z, alpha);
}

// ---- table-driven marching.  The step schedule of a cone (height h_i, lambda_i = lod_i + lodOffset) depends only on
// the cone settings, not on the pixel (phong.frag:145-177: r = h*tan(theta/2); lod = log2(max(1,2r)); h += r), so the
// host builds it once per frame (api.cu build_schedule) and every CTA copies it to shared memory.  lambda is
// non-decreasing along the march, which splits the steps into three runs by sampler state
// (Application.cpp:1094-1098): [0,n_point) lambda <= 0.5 magnifies -> NEAREST on level 0; [n_point,n_last)
// trilinear + mip-linear; [n_last,steps) lambda >= L-1 -> the last level only (one trilinear fetch, not two).
typedef ConeSchedule Schedule;
__device__ __forceinline__ bool inside_unit(V3 p) { return p.x >= 0.0f && p.x <= 1.0f && p.y >= 0.0f && p.y <= 1.0f && p.z >= 0.0f && p.z <= 1.0f; }   // false for NaN
// the same test for a position known to be free of NaN (cones with NaN start/direction never enter the marching loops)
__device__ __forceinline__ bool inside_unit_finite(V3 p) { return fminf(fminf(p.x, p.y), p.z) >= 0.0f && fmaxf(fmaxf(p.x, p.y), p.z) <= 1.0f; }
__device__ __forceinline__ bool has_nan(V3 a, V3 b) {     // NaN or infinity anywhere: the cone's first sample fails the s == clamp(s,0,1) test
    const float m = 3.0e38f;
    return !(fabsf(a.x) <= m && fabsf(a.y) <= m && fabsf(a.z) <= m && fabsf(b.x) <= m && fabsf(b.y) <= m && fabsf(b.z) <= m);
}
// Largest march height (in voxels) up to which start + ds*h certainly passes the reference's per-component test
// s == clamp(s,0,1): a slab test against the unit cube shrunk by 1e-5.  Steps beyond it take the exact test.
__device__ __forceinline__ float safe_height(V3 start, V3 ds) {
    const float lo = 1e-5f, hi = 1.0f - 1e-5f;
    auto axis = [&](float s, float d) {
        if (!(s > lo && s < hi)) return -1.0f;                              // also NaN
        if (d > 0.0f) return __fdividef(hi - s, d);
        if (d < 0.0f) return __fdividef(lo - s, d);
        return d == 0.0f ? 3.0e38f : -1.0f;                                 // NaN direction: never safe
    };
    return fminf(axis(start.x, ds.x), fminf(axis(start.y, ds.y), axis(start.z, ds.z)));
}
enum { SAMPLE_POINT = 0, SAMPLE_MIP = 1, SAMPLE_LAST = 2 };
template <int KIND, int WM>
__device__ __forceinline__ float4 fetch_volume(const ConeCtx& cx, V3 sp, float lambda, float max_lod) {
    if (WM == WARP_TEXTURE) sp = warp_sample(cx.warp, sp);
    else if (WM == WARP_VOXELS) sp = voxel_warp(sp, cx.eye_tc);
    else if (WM == WARP_TESS) sp = tess_warp_sample(*cx.fp, sp);
    if (KIND == SAMPLE_POINT) return tex3DLod<float4>(cx.vol_point, sp.x, sp.y, sp.z, 0.0f);
    if (KIND == SAMPLE_LAST) return tex3DLod<float4>(cx.vol_last, sp.x, sp.y, sp.z, max_lod);
    return tex3DLod<float4>(cx.vol, sp.x, sp.y, sp.z, lambda);
}
// N cones sharing one schedule, marched in lock step: N independent texture fetches are in flight per thread.
// Each cone's own accumulation sequence is exactly the reference's (front-to-back, stop at alpha >= 0.95 or when
// the sample leaves the unit cube).
template <int N> struct ConeSet { V3 ds[N]; float hsafe[N]; V4 acc[N]; unsigned alive; };     // alive: bit c = cone c still marching
// Branch-free inner loop: liveness is a bit mask, dead cones fetch nothing (predicated TEX) and accumulate with weight 0,
// so the N fetches of a step issue back to back and the compiler keeps every per-cone value in registers.
template <int N, int KIND, int WM>
__device__ __forceinline__ void march_run(const ConeCtx& cx, const Schedule& t, int i0, int i1, V3 start, ConeSet<N>& cs, unsigned& fetches) {
    const float max_lod = (float)(cx.L - 1);
    unsigned live = cs.alive;
    for (int i = i0; i < i1 && live; ++i) {
        const float hs = t.h[i], lambda = t.lambda[i];
        float4 smp[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const V3 sp = mk3(fmaf(cs.ds[c].x, hs, start.x), fmaf(cs.ds[c].y, hs, start.y), fmaf(cs.ds[c].z, hs, start.z));
            if (!((hs <= cs.hsafe[c]) | inside_unit_finite(sp))) live &= ~(1u << c);     // left the volume: phong.frag:163-165 break
            smp[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live >> c & 1u) smp[c] = fetch_volume<KIND, WM>(cx, sp, lambda, max_lod);
        }
        fetches += __popc(live);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const float a = (live >> c & 1u) ? 1.0f - cs.acc[c].w : 0.0f;
            cs.acc[c].x = fmaf(a, smp[c].x, cs.acc[c].x); cs.acc[c].y = fmaf(a, smp[c].y, cs.acc[c].y);
            cs.acc[c].z = fmaf(a, smp[c].z, cs.acc[c].z); cs.acc[c].w = fmaf(a, smp[c].w, cs.acc[c].w);
            if (!(cs.acc[c].w < 0.95f)) live &= ~(1u << c);                       // loop condition alpha < 0.95
        }
    }
    cs.alive = live;
}
template <int N, int WM>
__device__ __forceinline__ void trace_cones(const ConeCtx& cx, const Schedule& t, V3 start, ConeSet<N>& cs, unsigned& fetches) {
#pragma unroll
    cs.alive = (1u << N) - 1u;
    for (int c = 0; c < N; ++c) {
        cs.acc[c] = mk4(0.f, 0.f, 0.f, 0.f); cs.hsafe[c] = safe_height(start, cs.ds[c]);
        if (has_nan(start, cs.ds[c])) cs.alive &= ~(1u << c);               // the first sample is NaN: immediate break (phong.frag:163-165)
    }
    march_run<N, SAMPLE_POINT, WM>(cx, t, 0, t.n_point, start, cs, fetches);
    march_run<N, SAMPLE_MIP, WM>(cx, t, t.n_point, t.n_last, start, cs, fetches);
    march_run<N, SAMPLE_LAST, WM>(cx, t, t.n_last, t.steps, start, cs, fetches);
}
// One cone, K steps fetched ahead: the sample positions do not depend on earlier samples, only the decision to go on
// does, so up to K-1 fetches may be discarded when the cone saturates.  Same branch-free form: `valid` is the prefix
// mask of the steps that were inside the volume (issuing stops at the first miss).
template <int K, int KIND, int WM>
__device__ __forceinline__ void march_ahead(const ConeCtx& cx, const Schedule& t, int i0, int i1, V3 start, V3 ds, float hsafe, V4& acc, bool& alive, unsigned& fetches) {
    const float max_lod = (float)(cx.L - 1);
    for (int i = i0; i < i1 && alive; i += K) {
        float4 smp[K]; unsigned valid = 0u; bool open = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = min(i + k, i1 - 1);
            const float hs = t.h[j];
            const V3 sp = mk3(fmaf(ds.x, hs, start.x), fmaf(ds.y, hs, start.y), fmaf(ds.z, hs, start.z));
            open = open & (i + k < i1) & ((hs <= hsafe) | inside_unit_finite(sp));
            smp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (open) { smp[k] = fetch_volume<KIND, WM>(cx, sp, t.lambda[j], max_lod); valid |= 1u << k; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool use = (valid >> k & 1u) && alive;
            const float a = use ? 1.0f - acc.w : 0.0f;
            acc.x = fmaf(a, smp[k].x, acc.x); acc.y = fmaf(a, smp[k].y, acc.y); acc.z = fmaf(a, smp[k].z, acc.z); acc.w = fmaf(a, smp[k].w, acc.w);
            fetches += use ? 1u : 0u;
            if (!(acc.w < 0.95f)) alive = false;
        }
        const int nvalid = __popc(valid);
        if (nvalid < K && i + nvalid < i1) alive = false;                   // left the volume
    }
}
template <int K, int WM>
__device__ __forceinline__ V4 trace_cone_ahead(const ConeCtx& cx, const Schedule& t, V3 start, V3 ds, unsigned& fetches) {
    V4 acc = mk4(0.f, 0.f, 0.f, 0.f); bool alive = !has_nan(start, ds);
    const float hsafe = safe_height(start, ds);
    march_ahead<K, SAMPLE_POINT, WM>(cx, t, 0, t.n_point, start, ds, hsafe, acc, alive, fetches);
    march_ahead<K, SAMPLE_MIP, WM>(cx, t, t.n_point, t.n_last, start, ds, hsafe, acc, alive, fetches);
    march_ahead<K, SAMPLE_LAST, WM>(cx, t, t.n_last, t.steps, start, ds, hsafe, acc, alive, fetches);
    return acc;
}

__device__ __forceinline__ float pow2f(float x) { return x * x; }
__device__ __forceinline__ float D_ggxtr(V3 N, V3 Hh, float rough) {
    const float ndh = fmaxf(0.0f, dot3(N, Hh)), a2 = pow2f(rough);
    return a2 / (kPI * pow2f(pow2f(ndh) * (a2 - 1.0f) + 1.0f));
}
__device__ __forceinline__ float G1(V3 N, V3 V, float rough) {
    const float k = pow2f(rough + 1.0f) / 8.0f, ndv = fmaxf(0.0f, dot3(N, V));
    return ndv / (ndv * (1.0f - k) + k);
}
struct LR { V3 diffuse, specular; };
__device__ __forceinline__ LR cook_torrance(V3 dc, V3 lc, V3 N, V3 V, V3 L, V3 Hh, float rough, float metal) {
    const float ndl = fmaxf(0.0f, dot3(N, L)), ndv = fmaxf(0.0f, dot3(N, V)), hdv = fmaxf(0.0f, dot3(Hh, V));
    const V3 lambert = dc / kPI;
    const float Dg = D_ggxtr(N, Hh, rough), G = G1(N, V, rough) * G1(N, L, rough);
    const V3 F0 = mk3(mixf(0.04f, dc.x, metal), mixf(0.04f, dc.y, metal), mixf(0.04f, dc.z, metal));
    const float p5 = powf(1.0f - hdv, 5.0f);
    const V3 F = mk3(F0.x + (1.0f - F0.x) * p5, F0.y + (1.0f - F0.y) * p5, F0.z + (1.0f - F0.z) * p5);
    const float den = fmaxf(4.0f * ndl * ndv, 0.001f);
    const V3 fct = (F * (Dg * G)) / den;
    const V3 kd = mk3((1.0f - F.x) * (1.0f - metal), (1.0f - F.y) * (1.0f - metal), (1.0f - F.z) * (1.0f - metal));
    LR r; r.diffuse = ((lc * ndl) * kd) * lambert; r.specular = (lc * ndl) * fct;
    return r;
}

// kThreads per CTA (4 warps side by side, each an 8x4 pixel tile); 4 CTAs per SM = 16 warps at 128 registers.  Measured and dropped
// (profiles/r01d, r01g, r02a): CTAs of 32/64/256 threads, 3 or 5 CTAs per SM, the last level filtered from shared memory, an L2
// prefetch of the inputs, and a split into a set-up and a march kernel (march 669 us + set-up 136 us against 752 us in one kernel).
// DBG: the debug views of phong.frag (vct_frame_params::debug_view != 0; :346-447, 489-505) live in their own instantiation,
// so the shaded frame's code is the same with or without them (every DBG test below is a compile-time constant).
template <int WM, bool DBG = false>
__global__ void __launch_bounds__(kThreads, 512 / kThreads) k_cone_trace(TraceArgs a) {
    const FrameConst& fc = *a.fc;
    const vct_frame_params& fp = fc.p;
    __shared__ Schedule s_diffuse, s_specular;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&fc.sched_diffuse);          // the two tables are adjacent
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_diffuse);
        static_assert(sizeof(Schedule) % 4 == 0, "schedule copy");
        for (int i = threadIdx.x; i < (int)(sizeof(Schedule) / 4); i += kThreads) { dst[i] = __ldg(src + i); reinterpret_cast<uint32_t*>(&s_specular)[i] = __ldg(src + sizeof(Schedule) / 4 + i); }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int px = blockIdx.x * (kThreads / 4) + w * 8 + (lane & 7);
    int py = a.y_lo + blockIdx.y * 4 + (lane >> 3);
    if (a.tiles) {                                                        // sharded: CTA b works on 32x4 pixels of own tile b / 32
        const uint32_t t = __ldg(a.tiles + (blockIdx.x >> 5));
        const int sub = blockIdx.x & 31;
        px = (int)(t & 0xFFFFu) + (sub & 1) * 32 + w * 8 + (lane & 7);
        py = (int)(t >> 16) + (sub >> 1) * 4 + (lane >> 3);
    }
    unsigned fetches = 0;
    if (px < a.W && py < a.y_hi) {
        const size_t o = (size_t)py * a.W + px;
        const unsigned long long key = a.vis[o];
        if (key == ~0ull) put_pixel(a, o, pack_unorm(mk4(fp.clear_color[0], fp.clear_color[1], fp.clear_color[2], 1.0f)));
        else do {                                                         // `break` = the shader's early `return`
            const uint32_t t = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
            const uint32_t i0 = __ldg(a.indices + 3 * (size_t)t), i1 = __ldg(a.indices + 3 * (size_t)t + 1), i2 = __ldg(a.indices + 3 * (size_t)t + 2);
            const V3 w0 = f4to3(__ldg(a.wpos + i0)), w1 = f4to3(__ldg(a.wpos + i1)), w2 = f4to3(__ldg(a.wpos + i2));
            // perspective-correct barycentrics (and their forward differences for the texture LOD)
            float l[3], lx[3], ly[3];
            {
                V4 c[3];
                c[0] = mul44(fc.projection, mul44(fc.view, mk4(w0.x, w0.y, w0.z, 1.0f)));
                c[1] = mul44(fc.projection, mul44(fc.view, mk4(w1.x, w1.y, w1.z, 1.0f)));
                c[2] = mul44(fc.projection, mul44(fc.view, mk4(w2.x, w2.y, w2.z, 1.0f)));
                float ha[3], hb[3], hc[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const V4 p = c[(i + 1) % 3], q = c[(i + 2) % 3];
                    ha[i] = __fmul_rn(p.y, q.w) - __fmul_rn(q.y, p.w); hb[i] = __fmul_rn(q.x, p.w) - __fmul_rn(p.x, q.w); hc[i] = __fmul_rn(p.x, q.y) - __fmul_rn(q.x, p.y);
                }
                const float nx = ((float)px + 0.5f) / (float)a.W * 2.0f - 1.0f, ny = ((float)py + 0.5f) / (float)a.H * 2.0f - 1.0f;
                auto ev = [&](float x, float y, float out[3]) {
                    const float b0 = ha[0] * x + hb[0] * y + hc[0], b1 = ha[1] * x + hb[1] * y + hc[1], b2 = ha[2] * x + hb[2] * y + hc[2];
                    const float s = b0 + b1 + b2; out[0] = b0 / s; out[1] = b1 / s; out[2] = b2 / s;
                };
                ev(nx, ny, l); ev(nx + 2.0f / (float)a.W, ny, lx); ev(nx, ny + 2.0f / (float)a.H, ly);
            }
            const float* v0 = a.verts + 14 * (size_t)i0; const float* v1 = a.verts + 14 * (size_t)i1; const float* v2 = a.verts + 14 * (size_t)i2;
            const float u0 = __ldg(v0 + 6), t0 = __ldg(v0 + 7), u1 = __ldg(v1 + 6), t1 = __ldg(v1 + 7), u2 = __ldg(v2 + 6), t2 = __ldg(v2 + 7);
            const float u = ip(l, u0, u1, u2), v = ip(l, t0, t1, t2);
            const float ux = ip(lx, u0, u1, u2) - u, vx = ip(lx, t0, t1, t2) - v, uy = ip(ly, u0, u1, u2) - u, vy = ip(ly, t0, t1, t2) - v;
            auto fetch = [&](int ti) {
                const DevTexture& T = a.tex[ti];
                const float ax = ux * (float)T.w, bx = vx * (float)T.h, ay = uy * (float)T.w, by = vy * (float)T.h;
                return sample2d(T, u, v, fmaxf(ax * ax + bx * bx, ay * ay + by * by));
            };
            const DevMaterial mat = a.mats[__ldg(a.trimat + t)];
            const V3 Pw = ip3(l, w0, w1, w2);
            const V3 fn = ip3(l, f4to3(__ldg(a.wnrm + i0)), f4to3(__ldg(a.wnrm + i1)), f4to3(__ldg(a.wnrm + i2)));
            const V3 Tt = ip3(l, f4to3(__ldg(a.wT + i0)), f4to3(__ldg(a.wT + i1)), f4to3(__ldg(a.wT + i2)));
            const V3 Bt = ip3(l, f4to3(__ldg(a.wB + i0)), f4to3(__ldg(a.wB + i1)), f4to3(__ldg(a.wB + i2)));
            const V4 lf0 = mul44(fc.ls, mk4(w0.x, w0.y, w0.z, 1.0f)), lf1 = mul44(fc.ls, mk4(w1.x, w1.y, w1.z, 1.0f)), lf2 = mul44(fc.ls, mk4(w2.x, w2.y, w2.z, 1.0f));
            const V4 lsp = mk4(ip(l, lf0.x, lf1.x, lf2.x), ip(l, lf0.y, lf1.y, lf2.y), ip(l, lf0.z, lf1.z, lf2.z), ip(l, lf0.w, lf1.w, lf2.w));
            auto tbn = [&](V3 d) { return (Tt * d.x + Bt * d.y) + fn * d.z; };
            const int view = DBG ? fp.debug_view : 0;
            if (DBG && (view == VCT_VIEW_WARP_TEXTURE || view == VCT_VIEW_WARP_TEXTURE_TC)) {   // phong.frag:354-357 (voxelize && debugWarpTexture)
                const V3 tc = voxel_linear_position(Pw, fp);
                const V3 wv = view == VCT_VIEW_WARP_TEXTURE_TC ? tc : warp_sample_plain(a.warp, tc);
                put_pixel(a, o, pack_unorm(mk4(wv.x, wv.y, wv.z, 1.0f)));
                break;
            }
            if (DBG && (view == VCT_VIEW_VOXELS || view == VCT_VIEW_VOXEL_NORMALS)) {   // phong.frag:347-404: a volume at this fragment's voxel
                V3 gp = voxel_linear_position(Pw, fp);
                if (WM == WARP_VOXELS) gp = voxel_warp(gp, voxel_linear_position(mk3(fp.eye[0], fp.eye[1], fp.eye[2]), fp));
                else if (WM == WARP_TEXTURE) gp = warp_sample(a.warp, gp);
                else if (WM == WARP_TESS) gp = tess_warp_position(Pw, fp);
                const float Df = (float)fc.D;
                const V3 vi = mk3(__fdiv_rn(__fmul_rn(Df, gp.x), Df), __fdiv_rn(__fmul_rn(Df, gp.y), Df), __fdiv_rn(__fmul_rn(Df, gp.z), Df));   // voxelIndex(..) / voxelDim
                const float lambda = fp.miplevel;
                float4 sc;
                if (view == VCT_VIEW_VOXEL_NORMALS) {                        // :350-353: voxelNormal has one level and NEAREST filters: texel floor(s * D), border 0
                    const float fx = floorf(vi.x * Df), fy = floorf(vi.y * Df), fz = floorf(vi.z * Df);
                    V4 t4 = mk4(0.f, 0.f, 0.f, 0.f);
                    if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < Df && fy < Df && fz < Df) t4 = unpack_unorm(__ldg(a.normal0 + ((size_t)(int)fz * fc.D + (int)fy) * fc.D + (int)fx));
                    put_pixel(a, o, pack_unorm(mk4(t4.x, t4.y, t4.z, 1.0f)));
                    break;
                }
                if (!(lambda > 0.5f)) {
                    // magnification -> NEAREST on level 0 (GL 4.5 §8.14): texel floor(s * D), border 0.  Read from the linear level: the
                    // texture unit converts coordinates to fixed point before it floors, which flips texels that sit on a voxel boundary.
                    const float fx = floorf(vi.x * Df), fy = floorf(vi.y * Df), fz = floorf(vi.z * Df);
                    sc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < Df && fy < Df && fz < Df) {      // false for NaN
                        const V4 t4 = unpack_unorm(__ldg(a.level0 + ((size_t)(int)fz * fc.D + (int)fy) * fc.D + (int)fx));
                        sc = make_float4(t4.x, t4.y, t4.z, t4.w);
                    }
                }
                else sc = tex3DLod<float4>(a.vol, vi.x, vi.y, vi.z, fminf(lambda, (float)(fc.L - 1)));
                fetches++;
                put_pixel(a, o, pack_unorm(mk4(sc.x, sc.y, sc.z, 1.0f)));
                break;
            }
            if (DBG && (view == VCT_VIEW_MATERIAL_DIFFUSE || view == VCT_VIEW_MATERIAL_ROUGHNESS || view == VCT_VIEW_MATERIAL_METALLIC)) {   // :405-425
                V3 c = mk3(0.5f, 0.0f, 0.5f);
                if (view == VCT_VIEW_MATERIAL_DIFFUSE && mat.diffuse_tex >= 0) { const V4 t4 = fetch(mat.diffuse_tex); c = mk3(t4.x, t4.y, t4.z); }
                if (view == VCT_VIEW_MATERIAL_ROUGHNESS && mat.roughness_tex >= 0) { const float r = fetch(mat.roughness_tex).x; c = mk3(r, r, r); }
                if (view == VCT_VIEW_MATERIAL_METALLIC && mat.metallic_tex >= 0) { const float r = fetch(mat.metallic_tex).x; c = mk3(r, r, r); }
                put_pixel(a, o, pack_unorm(mk4(c.x, c.y, c.z, 1.0f)));
                break;
            }
            V3 N;
            if (fp.enable_normal_map && mat.normal_tex >= 0) {
                const V4 nm = fetch(mat.normal_tex);
                N = normalize3(tbn(normalize3(mk3(nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f))));
            } else N = normalize3(fn);
            if (DBG && view == VCT_VIEW_NORMALS) { put_pixel(a, o, pack_unorm(mk4(N.x, N.y, N.z, 1.0f))); break; }     // :441-443
            if (DBG && view == VCT_VIEW_DOMINANT_AXIS) {                     // :444-447  step(vec3(max component), |n|)
                const float ax = fabsf(N.x), ay = fabsf(N.y), az = fabsf(N.z), m = fmaxf(fmaxf(ax, ay), az);
                put_pixel(a, o, pack_unorm(mk4(ax < m ? 0.0f : 1.0f, ay < m ? 0.0f : 1.0f, az < m ? 0.0f : 1.0f, 1.0f)));
                break;
            }
            const V4 dc4 = mat.diffuse_tex >= 0 ? fetch(mat.diffuse_tex) : mk4(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2], 1.0f);
            const V3 dc = mk3(dc4.x, dc4.y, dc4.z);
            const V3 eye = mk3(fp.eye[0], fp.eye[1], fp.eye[2]);
            const V3 Vv = normalize3(eye - Pw);
            float rough = 0.5f; if (mat.roughness_tex >= 0) rough = fetch(mat.roughness_tex).x;
            float metal = 0.0f; if (mat.metallic_tex >= 0) metal = fetch(mat.metallic_tex).x;
            V3 dsum = mk3(0.f, 0.f, 0.f), ssum = mk3(0.f, 0.f, 0.f);
            for (int i = 0; i < fc.n_lights; ++i) {
                const vct_light& Lt = fc.lights[i];
                if (!Lt.enabled) continue;
                const V3 lc = mk3(Lt.color[0], Lt.color[1], Lt.color[2]), lpos = mk3(Lt.position[0], Lt.position[1], Lt.position[2]);
                LR r; r.diffuse = mk3(0.f, 0.f, 0.f); r.specular = mk3(0.f, 0.f, 0.f);
                if (Lt.type == 0u) {
                    const float dist = length3(lpos - Pw);
                    if (!(dist > Lt.range)) {
                        const float e0 = 0.75f * Lt.range, tt = clampf((dist - e0) / (Lt.range - e0), 0.0f, 1.0f);
                        const float att = 1.0f - tt * tt * (3.0f - 2.0f * tt);
                        const V3 Ld = normalize3(lpos - Pw);
                        if (fp.cooktorrance) { r = cook_torrance(dc, lc, N, Vv, Ld, normalize3(Vv + Ld), rough, metal); r.diffuse = r.diffuse * att; r.specular = r.specular * att; }
                        else {
                            const float df = fmaxf(dot3(N, Ld), 0.0f), sp = powf(fmaxf(dot3(N, normalize3(Vv + Ld)), 0.0f), mat.shininess);
                            r.diffuse = ((lc * (df * att)) * Lt.intensity) * dc; r.specular = ((lc * (sp * att)) * Lt.intensity) * dc;
                        }
                    }
                } else if (Lt.type == 1u) {
                    const V3 Ld = normalize3(mk3(-Lt.direction[0], -Lt.direction[1], -Lt.direction[2]));
                    if (fp.cooktorrance) r = cook_torrance(dc, lc, N, Vv, Ld, normalize3(Vv + Ld), rough, metal);
                    else {
                        const float df = fmaxf(dot3(N, Ld), 0.0f), sp = powf(fmaxf(dot3(N, normalize3(Vv + Ld)), 0.0f), mat.shininess);
                        r.diffuse = ((lc * df) * Lt.intensity) * dc; r.specular = ((lc * sp) * Lt.intensity) * dc;
                    }
                }
                if (Lt.shadow_caster) { const float sf = 1.0f - calc_shadow_factor(a.shadow, fc.S, lsp); r.diffuse = r.diffuse * sf; r.specular = r.specular * sf; }
                dsum = dsum + r.diffuse; ssum = ssum + r.specular;
            }
            if (!fp.enable_diffuse) dsum = mk3(0.f, 0.f, 0.f);
            if (!fp.enable_specular) ssum = mk3(0.f, 0.f, 0.f);
            V3 col;
            if (fp.enable_indirect) {
                ConeCtx cx; cx.vol = a.vol; cx.vol_point = a.vol_point; cx.vol_last = a.vol_last; if (WM == WARP_TESS) cx.fp = &fp; else cx.warp = a.warp; cx.D = fc.D; cx.L = fc.L;
                cx.warp_texture = fp.warp_texture; cx.warp_voxels = fp.warp_voxels; cx.eye_tc = voxel_linear_position(eye, fp);
                const V3 vp = voxel_linear_position(Pw, fp);
                const float scale = 1.0f / (float)fc.D;
                const float dirs[6][3] = {{0.f, 1.f, 0.f}, {0.f, 0.5f, 0.866025f}, {0.823639f, 0.5f, 0.267617f}, {0.509037f, 0.5f, -0.700629f},
                                          {-0.5909037f, 0.5f, -0.700629f}, {-0.823639f, 0.5f, 0.267617f}};
                const float wts[6] = {0.25f, 0.15f, 0.15f, 0.15f, 0.15f, 0.15f};
                V4 ind = mk4(0.f, 0.f, 0.f, 0.f);
                {
                    ConeSet<6> cs;
#pragma unroll
                    for (int i = 0; i < 6; ++i) cs.ds[i] = normalize3(normalize3(tbn(mk3(dirs[i][0], dirs[i][1], dirs[i][2])))) * scale;   // main() and traceCone() both normalise
                    const V3 start = vp + (N * fp.diffuse_cone.bias) * scale;
                    trace_cones<6, WM>(cx, s_diffuse, start, cs, fetches);
#pragma unroll
                    for (int i = 0; i < 6; ++i) ind = mk4(ind.x + wts[i] * cs.acc[i].x, ind.y + wts[i] * cs.acc[i].y, ind.z + wts[i] * cs.acc[i].z, ind.w + wts[i] * cs.acc[i].w);
                }
                const float occl = 1.0f - clampf(ind.w, 0.0f, 1.0f);
                if (DBG && view == VCT_VIEW_INDIRECT) {                      // :489 (before reflections and post-processing)
                    const float k = fp.draw_occlusion ? occl : 1.0f;
                    put_pixel(a, o, pack_unorm(mk4(ind.x * k, ind.y * k, ind.z * k, 1.0f)));
                    break;
                }
                if (DBG && view == VCT_VIEW_OCCLUSION) { put_pixel(a, o, pack_unorm(mk4(occl, occl, occl, 1.0f))); break; }   // :490
                if (fp.enable_reflections) {
                    const V3 I = Pw - eye;
                    const V3 R = I - N * (2.0f * dot3(N, I));
                    V4 rc;
                    if (fp.specular_cone_angle_from_roughness && mat.roughness_tex >= 0) {          // per-pixel cone angle: per-pixel schedule
                        const float ang = fetch(mat.roughness_tex).x * kPI * 0.1f;
                        rc = trace_cone<WM>(cx, vp, N, R, fp.specular_cone.steps, fp.specular_cone.bias, ang, fp.specular_cone.cone_initial_height, fp.specular_cone.lod_offset, fetches);
                    } else {
                        const V3 start = vp + (N * fp.specular_cone.bias) * scale;
                        rc = trace_cone_ahead<4, WM>(cx, s_specular, start, normalize3(R) * scale, fetches);
                    }
                    ind.x += rc.x * fp.reflect_scale; ind.y += rc.y * fp.reflect_scale; ind.z += rc.z * fp.reflect_scale;
                    if (DBG && view == VCT_VIEW_REFLECTIONS) { put_pixel(a, o, pack_unorm(mk4(rc.x, rc.y, rc.z, 1.0f))); break; }   // :505
                }
                const V3 indc = mk3(ind.x, ind.y, ind.z) * (dc * fp.ambient_scale);
                col = (indc + dsum) + ssum;
                if (fp.draw_occlusion) col = col * occl;
            } else col = (dc * fp.ambient_scale + dsum) + ssum;
            if (fp.enable_postprocess) {
                col = mk3(col.x / (col.x + 1.0f), col.y / (col.y + 1.0f), col.z / (col.z + 1.0f));
                const float g = 1.0f / 2.2f;
                col = mk3(powf(col.x, g), powf(col.y, g), powf(col.z, g));
            }
            put_pixel(a, o, pack_unorm(mk4(col.x, col.y, col.z, 1.0f)));
        } while (0);
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) fetches += __shfl_xor_sync(0xffffffffu, fetches, s);
    if (lane == 0 && fetches) atomicAdd(&a.counters->cone_steps, (unsigned long long)fetches);
}

}  // namespace
