// raster.cuh — canonical coverage rule + software samplers used by the voxeliser, shadow-map and visibility
// kernels.  Include only from translation units compiled with -fmad=false.
//
// Replaces the GL rasteriser the reference relies on (viewport D x D / S x S / W x H, reference
// src/Application.cpp:217, 668, 940).  Definition (DESIGN.md "Canonical GL semantics"):
//   window = (ndc*0.5+0.5)*size ; snapped = floor(window*256+0.5) (8 sub-pixel bits) ; 64-bit integer edge
//   functions ; samples at pixel centres ; top-left rule in y-up window space: a pixel exactly on edge a->b of
//   a CCW triangle is covered iff dy<0 || (dy==0 && dx<0) ; barycentric l_i = float(E_i)/float(area).
#pragma once
#include "common.cuh"

struct RV { float x, y, z, w; };

struct TriSetup {
    int X[3], Y[3];          // snapped window coordinates of the (re-ordered, CCW) slots
    long long area;          // > 0
    int bias[3];             // 0 for top-left edges, -1 otherwise (edge k is opposite slot k)
    int swapped;             // 1: slots (0,1,2) hold source vertices (0,2,1)
    int x0, x1, y0, y1;      // inclusive pixel bounding box, clipped to the viewport
    float z[3];              // ndc z per SOURCE vertex
};

__device__ __forceinline__ int snap_coord(float ndc, int size) {
    float wv = (ndc * 0.5f + 0.5f) * (float)size;
    float s = floorf(wv * 256.0f + 0.5f);
    if (!(s > -1073741824.0f)) s = -1073741824.0f;
    if (s > 1073741824.0f) s = 1073741824.0f;
    return (int)s;
}
__device__ __forceinline__ int cdiv256(long long a) { return (int)((a >= 0) ? (a + 255) / 256 : -((-a) / 256)); }
__device__ __forceinline__ int fdiv256(long long a) { return (int)((a >= 0) ? a / 256 : -((-a + 255) / 256)); }

__device__ __forceinline__ bool tri_setup(const RV v[3], int W, int H, bool cull_back, TriSetup& s) {
    int X[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { X[i] = snap_coord(v[i].x / v[i].w, W); Y[i] = snap_coord(v[i].y / v[i].w, H); }
    long long area = (long long)(X[1] - X[0]) * (long long)(Y[2] - Y[0]) - (long long)(Y[1] - Y[0]) * (long long)(X[2] - X[0]);
    if (area == 0) return false;
    s.swapped = 0;
    if (area < 0) { if (cull_back) return false; s.swapped = 1; area = -area; }
    s.X[0] = X[0]; s.Y[0] = Y[0];
    s.X[1] = s.swapped ? X[2] : X[1]; s.Y[1] = s.swapped ? Y[2] : Y[1];
    s.X[2] = s.swapped ? X[1] : X[2]; s.Y[2] = s.swapped ? Y[1] : Y[2];
    s.area = area;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const int dx = s.X[b] - s.X[a], dy = s.Y[b] - s.Y[a];
        s.bias[k] = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
        s.z[k] = v[k].z / v[k].w;
    }
    const int minx = min(s.X[0], min(s.X[1], s.X[2])), maxx = max(s.X[0], max(s.X[1], s.X[2]));
    const int miny = min(s.Y[0], min(s.Y[1], s.Y[2])), maxy = max(s.Y[0], max(s.Y[1], s.Y[2]));
    s.x0 = max(0, cdiv256((long long)minx - 128)); s.x1 = min(W - 1, fdiv256((long long)maxx - 128));
    s.y0 = max(0, cdiv256((long long)miny - 128)); s.y1 = min(H - 1, fdiv256((long long)maxy - 128));
    return s.x0 <= s.x1 && s.y0 <= s.y1;
}

// coverage test at pixel (px,py); on success writes the barycentrics indexed by SOURCE vertex
__device__ __forceinline__ bool tri_cover(const TriSetup& s, int px, int py, float l[3]) {
    const long long Px = 256ll * px + 128, Py = 256ll * py + 128;
    long long E[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        E[k] = (long long)(s.X[b] - s.X[a]) * (Py - s.Y[a]) - (long long)(s.Y[b] - s.Y[a]) * (Px - s.X[a]);
        if (E[k] + s.bias[k] < 0) return false;
    }
    const float fa = (float)s.area;
    const float e0 = (float)E[0] / fa, e1 = (float)E[1] / fa, e2 = (float)E[2] / fa;
    l[0] = e0; l[1] = s.swapped ? e2 : e1; l[2] = s.swapped ? e1 : e2;
    return true;
}
__device__ __forceinline__ float interp1(const float l[3], float a0, float a1, float a2) { return (l[0] * a0 + l[1] * a1) + l[2] * a2; }
__device__ __forceinline__ V3 interp3(const float l[3], V3 a, V3 b, V3 c) {
    return mk3(interp1(l, a.x, b.x, c.x), interp1(l, a.y, b.y, c.y), interp1(l, a.z, b.z, c.z));
}

// ---------------------------------------------------------------------------------- 2D material textures
// min LINEAR_MIPMAP_NEAREST, mag NEAREST, REPEAT (reference src/Graphics/GLHelper.cpp:180-183); rho2 is the
// squared GL scale factor; level selection by comparisons only (no log2).
__device__ __forceinline__ int wrapi(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }
__device__ __forceinline__ V4 texel2d(const DevTexture& t, int level, int x, int y) {
    const int w = max(1, t.w >> level), h = max(1, t.h >> level);
    const uint8_t* p = t.level[level] + ((size_t)wrapi(y, h) * w + wrapi(x, w)) * t.ch;
    V4 r = mk4(0.f, 0.f, 0.f, 1.f);
    if (t.ch == 4) { const uchar4 q = *reinterpret_cast<const uchar4*>(p); r = mk4((float)q.x / 255.0f, (float)q.y / 255.0f, (float)q.z / 255.0f, (float)q.w / 255.0f); }
    else if (t.ch == 3) r = mk4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, 1.0f);
    else r.x = (float)p[0] / 255.0f;
    return r;
}
__device__ __forceinline__ V4 lerp4(V4 a, V4 b, float t) {
    const float s = 1.0f - t;
    return mk4(a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t, a.w * s + b.w * t);
}
__device__ __forceinline__ V4 sample2d(const DevTexture& t, float u, float v, float rho2) {
    if (!(rho2 > 2.0f)) return texel2d(t, 0, (int)floorf(u * (float)t.w), (int)floorf(v * (float)t.h));
    int d = 1; float lim = 8.0f;
    while (d < t.levels - 1 && rho2 > lim) { d++; lim *= 4.0f; }
    if (d > t.levels - 1) d = t.levels - 1;
    const int w = max(1, t.w >> d), h = max(1, t.h >> d);
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const int x0 = (int)fx0, y0 = (int)fy0; const float fx = x - fx0, fy = y - fy0;
    V4 top = lerp4(texel2d(t, d, x0, y0), texel2d(t, d, x0 + 1, y0), fx);
    V4 bot = lerp4(texel2d(t, d, x0, y0 + 1), texel2d(t, d, x0 + 1, y0 + 1), fx);
    return lerp4(top, bot, fy);
}
// affine uv derivatives of an orthographically projected triangle -> rho^2
__device__ __forceinline__ float tri_rho2_affine(const RV v[3], const float uv[3][2], int W, int H, const DevTexture& t) {
    float x[3], y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x[i] = (v[i].x / v[i].w * 0.5f + 0.5f) * (float)W; y[i] = (v[i].y / v[i].w * 0.5f + 0.5f) * (float)H; }
    const float x1 = x[1] - x[0], x2 = x[2] - x[0], y1 = y[1] - y[0], y2 = y[2] - y[0];
    const float u1 = uv[1][0] - uv[0][0], u2 = uv[2][0] - uv[0][0], v1 = uv[1][1] - uv[0][1], v2 = uv[2][1] - uv[0][1];
    const float den = x1 * y2 - x2 * y1;
    const float dudx = (u1 * y2 - u2 * y1) / den, dudy = (u2 * x1 - u1 * x2) / den;
    const float dvdx = (v1 * y2 - v2 * y1) / den, dvdy = (v2 * x1 - v1 * x2) / den;
    const float ax = dudx * (float)t.w, bx = dvdx * (float)t.h, ay = dudy * (float)t.w, by = dvdy * (float)t.h;
    return maxsel(ax * ax + bx * bx, ay * ay + by * by);
}

// --------------------------------------------------------------------------------------------- shadow map
// LINEAR, CLAMP_TO_BORDER(1) (reference src/Application.cpp:45-53)
__device__ __forceinline__ float shadow_texel(const float* __restrict__ sm, int S, int x, int y) {
    return (x < 0 || y < 0 || x >= S || y >= S) ? 1.0f : __ldg(sm + (size_t)y * S + x);
}
__device__ __forceinline__ float shadow_linear(const float* __restrict__ sm, int S, float u, float v, int ox, int oy) {
    const float x = u * (float)S - 0.5f, y = v * (float)S - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f)) return 1.0f;
    const int x0 = (int)fx0 + ox, y0 = (int)fy0 + oy; const float fx = x - fx0, fy = y - fy0;
    const float top = shadow_texel(sm, S, x0, y0) * (1.0f - fx) + shadow_texel(sm, S, x0 + 1, y0) * fx;
    const float bot = shadow_texel(sm, S, x0, y0 + 1) * (1.0f - fx) + shadow_texel(sm, S, x0 + 1, y0 + 1) * fx;
    return top * (1.0f - fy) + bot * fy;
}
// voxelize.frag:160-184 == phong.frag:183-207
__device__ __forceinline__ float calc_shadow_factor(const float* __restrict__ sm, int S, V4 lsp) {
    const float sx = (lsp.x / lsp.w + 1.0f) * 0.5f, sy = (lsp.y / lsp.w + 1.0f) * 0.5f, sz = (lsp.z / lsp.w + 1.0f) * 0.5f;
    const float frag_depth = sz - 0.01f;
    if (frag_depth > 1.0f) return 0.0f;
    float f = 0.0f;
    if (frag_depth > shadow_linear(sm, S, sx, sy, 0, 0)) f += 1.0f;
    if (frag_depth > shadow_linear(sm, S, sx, sy, 1, 0)) f += 1.0f;
    if (frag_depth > shadow_linear(sm, S, sx, sy, 0, 1)) f += 1.0f;
    if (frag_depth > shadow_linear(sm, S, sx, sy, -1, 0)) f += 1.0f;
    if (frag_depth > shadow_linear(sm, S, sx, sy, 0, -1)) f += 1.0f;
    return f / 5.0f;
}

// ------------------------------------------------------------------------------ warpmap (exact software path)
// 32^3 RGBA16 unorm, LINEAR, CLAMP_TO_EDGE (reference src/Application.cpp:383-389)
__device__ __forceinline__ V3 warp_texel(const uint16_t* __restrict__ wm, int x, int y, int z) {
    const int n = VCT_WARP_DIM;
    x = min(max(x, 0), n - 1); y = min(max(y, 0), n - 1); z = min(max(z, 0), n - 1);
    const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(wm) + ((size_t)z * n + y) * n + x);
    return mk3((float)q.x / 65535.0f, (float)q.y / 65535.0f, (float)q.z / 65535.0f);
}
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, float t) { const float s = 1.0f - t; return mk3(a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t); }
__device__ __forceinline__ V3 warp_sample(const uint16_t* __restrict__ wm, V3 tc) {
    const float n = (float)VCT_WARP_DIM;
    const float x = tc.x * n - 0.5f, y = tc.y * n - 0.5f, z = tc.z * n - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f) || !(fabsf(fz0) < 1e9f)) return mk3(0.f, 0.f, 0.f);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    V3 c00 = lerp3(warp_texel(wm, x0, y0, z0), warp_texel(wm, x0 + 1, y0, z0), fx);
    V3 c10 = lerp3(warp_texel(wm, x0, y0 + 1, z0), warp_texel(wm, x0 + 1, y0 + 1, z0), fx);
    V3 c01 = lerp3(warp_texel(wm, x0, y0, z0 + 1), warp_texel(wm, x0 + 1, y0, z0 + 1), fx);
    V3 c11 = lerp3(warp_texel(wm, x0, y0 + 1, z0 + 1), warp_texel(wm, x0 + 1, y0 + 1, z0 + 1), fx);
    return lerp3(lerp3(c00, c10, fy), lerp3(c01, c11, fy), fz);
}

// ------------------------------------------------------------------------------------ common.glsl:6-64 (a8)
__device__ __forceinline__ V3 voxel_linear_position(V3 p, const vct_frame_params& fp) {
    return mk3((p.x - fp.voxel_center[0] - fp.voxel_min[0]) / (fp.voxel_max[0] - fp.voxel_min[0]),
               (p.y - fp.voxel_center[1] - fp.voxel_min[1]) / (fp.voxel_max[1] - fp.voxel_min[1]),
               (p.z - fp.voxel_center[2] - fp.voxel_min[2]) / (fp.voxel_max[2] - fp.voxel_min[2]));
}
__device__ __forceinline__ float voxel_warp_fn1(float x) {
    const float alpha = 0.25f;
    x = (alpha * x + (3.0f - 3.0f * alpha) * x * x) + (2.0f * alpha - 2.0f) * x * x * x;
    return clampf(x, 0.0f, 1.0f);
}
__device__ __forceinline__ V3 voxel_warp(V3 p, V3 c) {
    V3 o = p - c;
    o = mk3(0.5f * o.x + 0.5f, 0.5f * o.y + 0.5f, 0.5f * o.z + 0.5f);
    o = mk3(voxel_warp_fn1(o.x), voxel_warp_fn1(o.y), voxel_warp_fn1(o.z));
    o = mk3(2.0f * o.x - 1.0f, 2.0f * o.y - 1.0f, 2.0f * o.z - 1.0f);
    return c + o;
}
__device__ __forceinline__ V3 eye_of(const vct_frame_params& fp) { return mk3(fp.eye[0], fp.eye[1], fp.eye[2]); }
__device__ __forceinline__ V3 get_voxel_position(V3 pos, const vct_frame_params& fp, const uint16_t* __restrict__ warpmap) {
    if (fp.warp_voxels) return voxel_warp(voxel_linear_position(pos, fp), voxel_linear_position(eye_of(fp), fp));
    if (fp.warp_texture) return warp_sample(warpmap, voxel_linear_position(pos, fp));
    return voxel_linear_position(pos, fp);
}
// ivec3(vec3) truncation + image bounds (out-of-bounds image access is a no-op)
__device__ __forceinline__ bool to_voxel_index(V3 p, int D, int& ix, int& iy, int& iz) {
    const float fd = (float)D;
    if (!(p.x > -1.0f) || !(p.x < fd) || !(p.y > -1.0f) || !(p.y < fd) || !(p.z > -1.0f) || !(p.z < fd)) return false;
    ix = (int)p.x; iy = (int)p.y; iz = (int)p.z;
    return true;
}
