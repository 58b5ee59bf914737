// vct_oracle.cpp — CPU ORACLE (test infrastructure only).  Pinned bit for bit against the reference's own sources compiled in
// place (oracle/_ref, see vct_oracle.h): the host C++ of the warp tables/example, and the GLSL of transferVoxels,
// filterRadiance, voxelFillHoles, injectRadiance, the dead post-pass variants, voxelize.frag and phong.frag compiled as C++
// and generateWarpmap{,Weights}.frag (tests/test_glsl_ref.py).  Vertex/geometry stages: see vct_oracle.h.  Still unpinned: the two alpha tests.
//
// A restatement of the reference's GLSL passes in plain C++ with OpenGL's implementation-defined behaviour
// fixed to one explicit definition (DESIGN.md "Canonical GL semantics").  Compile with -ffp-contract=off:
// every float operation below is a separately rounded IEEE-754 binary32 op, in the order written.
//
// Each function cites the reference file:line it follows.
#include "vct_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------------ math
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
inline V3 v3(float x, float y, float z) { return {x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }   // GLSL clamp = min(max(x,lo),hi)
inline float maxf(float a, float b) { return a > b ? a : b; }   // GLSL max: returns b if a<b else a  (NaN a -> a)
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// mat4 (column-major) * vec4, accumulated left to right
inline V4 mul(const float* m, V4 v) {
    V4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}
// mat3(transpose(inverse(M))) = cofactor(M3)/det   (simple.vert:16, voxelize.vert:18, phong.vert:37)
inline void normal_matrix(const float* m, float n[9]) {
    const float a = m[0], b = m[4], c = m[8], d = m[1], e = m[5], f = m[9], g = m[2], h = m[6], i = m[10];
    // row-major view: [a b c; d e f; g h i]; cofactors
    const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    const float c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
    const float c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
    const float det = (a * c00 + b * c01) + c * c02;
    // n stored row-major: n[r*3+c] = cofactor(r,c)/det
    n[0] = c00 / det; n[1] = c01 / det; n[2] = c02 / det;
    n[3] = c10 / det; n[4] = c11 / det; n[5] = c12 / det;
    n[6] = c20 / det; n[7] = c21 / det; n[8] = c22 / det;
}
inline V3 mul3(const float n[9], V3 v) {
    return {(n[0] * v.x + n[1] * v.y) + n[2] * v.z, (n[3] * v.x + n[4] * v.y) + n[5] * v.z,
            (n[6] * v.x + n[7] * v.y) + n[8] * v.z};
}

// ------------------------------------------------------------------------------------ unorm conversions
inline uint32_t f2u_trunc(float v) {                   // GLSL uint(float): truncation; NaN/negative -> 0, saturating
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967040.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline uint32_t unorm8(float v) {                      // imageStore / packUnorm4x8: round(clamp(v,0,1)*255), ties to even
    if (!(v > 0.0f)) return 0u;                        // NaN -> 0
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)std::nearbyintf(v * 255.0f);
}
inline uint32_t pack_unorm(V4 c) { return unorm8(c.x) | unorm8(c.y) << 8 | unorm8(c.z) << 16 | unorm8(c.w) << 24; }
inline V4 unpack_unorm(uint32_t w) {                   // imageLoad rgba8
    return {(float)(w & 255u) / 255.0f, (float)((w >> 8) & 255u) / 255.0f, (float)((w >> 16) & 255u) / 255.0f,
            (float)(w >> 24) / 255.0f};
}

// IEEE half <-> float (RGBA16F render targets of the warp-weights pass)
inline uint16_t f2h(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | (x > 0x7F800000u ? 0x7E00u : 0x7C00u));
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);              // rounds to inf
    if (x < 0x33000001u) return (uint16_t)sign;                           // rounds to zero
    int e = (int)(x >> 23) - 127; uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift; uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7FFFFFu; }
    uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    uint32_t h = base + q;
    if (rem > half || (rem == half && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}
inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int k = 0; while (!(m & 1024u)) { m <<= 1; k++; } x = sign | (uint32_t)(113 - k) << 23 | (m & 1023u) << 13; }
    } else if (e == 31) x = sign | 0x7F800000u | m << 13;
    else x = sign | (e + 112u) << 23 | m << 13;
    float f; std::memcpy(&f, &x, 4); return f;
}

// ------------------------------------------------------------------------------------------- 2D textures
// Sampler state of every material texture: min LINEAR_MIPMAP_NEAREST, mag NEAREST, wrap REPEAT
// (GLHelper.cpp:180-183).  Anisotropy is pinned OFF.
struct Tex { int w, h, ch, levels; const uint8_t* lv[16]; };
inline int wrapi(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }
inline V4 texel2d(const Tex& t, int level, int x, int y) {
    const int w = std::max(1, t.w >> level), h = std::max(1, t.h >> level);
    const uint8_t* p = t.lv[level] + ((size_t)wrapi(y, h) * w + wrapi(x, w)) * t.ch;
    V4 r = {0, 0, 0, 1};
    r.x = (float)p[0] / 255.0f;
    if (t.ch >= 3) { r.y = (float)p[1] / 255.0f; r.z = (float)p[2] / 255.0f; }
    if (t.ch == 4) r.w = (float)p[3] / 255.0f;
    return r;
}
inline V4 lerp4(V4 a, V4 b, float t) {
    const float s = 1.0f - t;
    return {a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t, a.w * s + b.w * t};
}
// rho2 = squared scale factor rho^2 of GL 4.5 §8.14.1 (isotropic).  lambda = log2(rho) is never formed:
//   magnification  <=> lambda <= c = 0.5 <=> rho2 <= 2      -> NEAREST on level 0
//   else level d = ceil(lambda + 0.5) - 1 = smallest d >= 1 with rho2 <= 2^(2d+1), clamped to the last level
inline V4 sample2d(const Tex& t, float u, float v, float rho2) {
    if (!(rho2 > 2.0f)) {
        const int x = (int)std::floor(u * (float)t.w), y = (int)std::floor(v * (float)t.h);
        return texel2d(t, 0, x, y);
    }
    int d = 1; float lim = 8.0f;
    while (d < t.levels - 1 && rho2 > lim) { d++; lim *= 4.0f; }
    if (d > t.levels - 1) d = t.levels - 1;
    const int w = std::max(1, t.w >> d), h = std::max(1, t.h >> d);
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const int x0 = (int)fx0, y0 = (int)fy0; const float fx = x - fx0, fy = y - fy0;
    V4 top = lerp4(texel2d(t, d, x0, y0), texel2d(t, d, x0 + 1, y0), fx);
    V4 bot = lerp4(texel2d(t, d, x0, y0 + 1), texel2d(t, d, x0 + 1, y0 + 1), fx);
    return lerp4(top, bot, fy);
}

// shadow map: LINEAR, CLAMP_TO_BORDER border 1 (Application.cpp:45-53).  textureOffset adds integer texel offsets.
inline float shadow_texel(const float* sm, int S, int x, int y) {
    return (x < 0 || y < 0 || x >= S || y >= S) ? 1.0f : sm[(size_t)y * S + x];
}
inline float shadow_linear(const float* sm, int S, float u, float v, int ox, int oy) {
    const float x = u * (float)S - 0.5f, y = v * (float)S - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    if (!(std::fabs(fx0) < 1e9f) || !(std::fabs(fy0) < 1e9f)) return 1.0f;     // NaN / huge -> border
    const int x0 = (int)fx0 + ox, y0 = (int)fy0 + oy; const float fx = x - fx0, fy = y - fy0;
    const float top = shadow_texel(sm, S, x0, y0) * (1.0f - fx) + shadow_texel(sm, S, x0 + 1, y0) * fx;
    const float bot = shadow_texel(sm, S, x0, y0 + 1) * (1.0f - fx) + shadow_texel(sm, S, x0 + 1, y0 + 1) * fx;
    return top * (1.0f - fy) + bot * fy;
}
// voxelize.frag:160-184 == phong.frag:183-207
inline float calc_shadow_factor(const float* sm, int S, V4 lsp) {
    const float sx = (lsp.x / lsp.w + 1.0f) * 0.5f, sy = (lsp.y / lsp.w + 1.0f) * 0.5f, sz = (lsp.z / lsp.w + 1.0f) * 0.5f;
    const float frag_depth = sz - 0.01f;
    if (frag_depth > 1.0f) return 0.0f;
    static const int off[5][2] = {{0, 0}, {1, 0}, {0, 1}, {-1, 0}, {0, -1}};
    float f = 0.0f;
    for (int i = 0; i < 5; ++i)
        if (frag_depth > shadow_linear(sm, S, sx, sy, off[i][0], off[i][1])) f += 1.0f;
    return f / 5.0f;
}

// warpmap: 32^3 RGBA16 unorm, LINEAR, CLAMP_TO_EDGE (Application.cpp:383-389)
inline V3 warp_texel(const uint16_t* wm, int x, int y, int z) {
    const int n = VCT_WARP_DIM;
    x = std::min(std::max(x, 0), n - 1); y = std::min(std::max(y, 0), n - 1); z = std::min(std::max(z, 0), n - 1);
    const uint16_t* p = wm + (((size_t)z * n + y) * n + x) * 4;
    return {(float)p[0] / 65535.0f, (float)p[1] / 65535.0f, (float)p[2] / 65535.0f};
}
inline V3 lerp3(V3 a, V3 b, float t) { const float s = 1.0f - t; return {a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t}; }
inline V3 warp_sample(const uint16_t* wm, V3 tc) {
    const float n = (float)VCT_WARP_DIM;
    const float x = tc.x * n - 0.5f, y = tc.y * n - 0.5f, z = tc.z * n - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y), fz0 = std::floor(z);
    if (!(std::fabs(fx0) < 1e9f) || !(std::fabs(fy0) < 1e9f) || !(std::fabs(fz0) < 1e9f)) return {0, 0, 0};
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    V3 c00 = lerp3(warp_texel(wm, x0, y0, z0), warp_texel(wm, x0 + 1, y0, z0), fx);
    V3 c10 = lerp3(warp_texel(wm, x0, y0 + 1, z0), warp_texel(wm, x0 + 1, y0 + 1, z0), fx);
    V3 c01 = lerp3(warp_texel(wm, x0, y0, z0 + 1), warp_texel(wm, x0 + 1, y0, z0 + 1), fx);
    V3 c11 = lerp3(warp_texel(wm, x0, y0 + 1, z0 + 1), warp_texel(wm, x0 + 1, y0 + 1, z0 + 1), fx);
    return lerp3(lerp3(c00, c10, fy), lerp3(c01, c11, fy), fz);
}

// ------------------------------------------------------------------------------- common.glsl:6-64 (a8)
inline V3 voxel_linear_position(V3 p, const vct_frame_params* fp) {
    return {(p.x - fp->voxel_center[0] - fp->voxel_min[0]) / (fp->voxel_max[0] - fp->voxel_min[0]),
            (p.y - fp->voxel_center[1] - fp->voxel_min[1]) / (fp->voxel_max[1] - fp->voxel_min[1]),
            (p.z - fp->voxel_center[2] - fp->voxel_min[2]) / (fp->voxel_max[2] - fp->voxel_min[2])};
}
inline float voxel_warp_fn1(float x) {                // common.glsl:11-18
    const float alpha = 0.25f;
    x = (alpha * x + (3.0f - 3.0f * alpha) * x * x) + (2.0f * alpha - 2.0f) * x * x * x;
    return clampf(x, 0.0f, 1.0f);
}
inline V3 voxel_warp(V3 p, V3 c) {                    // common.glsl:20-27
    V3 o = p - c;
    o = {0.5f * o.x + 0.5f, 0.5f * o.y + 0.5f, 0.5f * o.z + 0.5f};
    o = {voxel_warp_fn1(o.x), voxel_warp_fn1(o.y), voxel_warp_fn1(o.z)};
    o = {2.0f * o.x - 1.0f, 2.0f * o.y - 1.0f, 2.0f * o.z - 1.0f};
    return c + o;
}
inline V3 eye3(const vct_frame_params* fp) { return {fp->eye[0], fp->eye[1], fp->eye[2]}; }
// common.glsl:37-42 (voxelizeTesselationWarp: the voxel grid is the camera frustum — NDC of pv mapped to [0,1]^3)
inline V3 tess_warp_position(V3 pos, const vct_frame_params* fp) {
    V4 q = mul(fp->pv, {pos.x, pos.y, pos.z, 1.0f});
    q.x /= q.w; q.y /= q.w; q.z /= q.w;
    return {q.x * 0.5f + 0.5f, q.y * 0.5f + 0.5f, q.z * 0.5f + 0.5f};
}
// common.glsl:44-60, priority warpVoxels > warpTexture > voxelizeTesselationWarp > linear
inline V3 get_voxel_position(V3 pos, const vct_frame_params* fp, const uint16_t* warpmap) {
    if (fp->warp_voxels) return voxel_warp(voxel_linear_position(pos, fp), voxel_linear_position(eye3(fp), fp));
    if (fp->warp_texture && warpmap) return warp_sample(warpmap, voxel_linear_position(pos, fp));
    if (fp->voxelize_tesselation_warp) return tess_warp_position(pos, fp);
    return voxel_linear_position(pos, fp);
}
// testTesselation.tese:134: voxelIndex(position.xyz, ..., false) inside the tessellation program — warpVoxels is passed as false
// and the host never sets that program's warpTexture uniform (Application.cpp:630-640), so only the tessellation warp or the
// linear mapping can apply, whatever the frame's other warp settings are
inline V3 tese_voxel_position(V3 pos, const vct_frame_params* fp) {
    return fp->voxelize_tesselation_warp ? tess_warp_position(pos, fp) : voxel_linear_position(pos, fp);
}

// --------------------------------------------------------------------------------- scene preprocessing
struct Prepared {
    std::vector<V3> wpos, wnrm, T, B;     // world position, normalMatrix*normal (NOT normalised), phong.vert T and B
    std::vector<Tex> tex;
};
Prepared prepare(const orc_scene* sc, bool tbn) {
    Prepared P;
    P.wpos.resize(sc->n_vertices); P.wnrm.resize(sc->n_vertices);
    if (tbn) { P.T.resize(sc->n_vertices); P.B.resize(sc->n_vertices); }
    std::vector<float> nm((size_t)sc->n_actors * 9);
    for (int a = 0; a < sc->n_actors; ++a) normal_matrix(sc->actor_model + 16 * a, &nm[9 * a]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < sc->n_vertices; ++i) {
        const float* v = sc->vertices + (size_t)14 * i;
        const int a = sc->vertex_actor[i];
        V4 w = mul(sc->actor_model + 16 * a, {v[0], v[1], v[2], 1.0f});
        P.wpos[i] = {w.x, w.y, w.z};
        const float* n = &nm[9 * a];
        V3 N = mul3(n, {v[3], v[4], v[5]});
        P.wnrm[i] = N;
        if (tbn) {                                        // phong.vert:49-54
            V3 T = mul3(n, {v[8], v[9], v[10]});
            T = normalize(T - N * dot(T, N));
            P.T[i] = T;
            P.B[i] = cross(N, T);
        }
    }
    P.tex.resize(sc->n_textures);
    for (int t = 0; t < sc->n_textures; ++t) {
        const orc_texture& s = sc->textures[t];
        Tex& d = P.tex[t]; d.w = s.width; d.h = s.height; d.ch = s.channels; d.levels = s.levels;
        for (int l = 0; l < s.levels && l < 16; ++l) d.lv[l] = sc->texels + s.offset[l];
    }
    return P;
}

// ---- clip transforms of the camera / light passes.  Default ("stepwise"): P * (V * world), ls * world — one world-space
// position shared by every pass (what the CUDA passes do).  Literal mode (orc_set_literal_vertex_transforms, tests only): the
// shaders' own association, simple.vert:21 / phong.vert:41,45: ((projection * view) * model) * vec4(position, 1) and
// (ls * model) * vec4(position, 1).  Same real-number result, different fp32 rounding.
static int g_literal_vertex_transforms = 0;
inline void matmul44(const float* a, const float* b, float* r) {            // r = a * b, column by column
    for (int j = 0; j < 4; ++j) { V4 c = mul(a, {b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]}); r[4 * j] = c.x; r[4 * j + 1] = c.y; r[4 * j + 2] = c.z; r[4 * j + 3] = c.w; }
}
struct ClipMats {                                                            // per actor: (A * B) * model
    std::vector<float> m; bool literal;
    ClipMats(const orc_scene* sc, const float* A, const float* B) : literal(g_literal_vertex_transforms != 0) {
        if (!literal) return;
        float ab[16];
        if (B) matmul44(A, B, ab); else std::memcpy(ab, A, sizeof ab);
        m.resize((size_t)sc->n_actors * 16);
        for (int a = 0; a < sc->n_actors; ++a) matmul44(ab, sc->actor_model + 16 * (size_t)a, &m[16 * (size_t)a]);
    }
    V4 apply(const orc_scene* sc, const float* A, const float* B, unsigned vi, V3 w) const {
        if (!literal) return B ? mul(A, mul(B, {w.x, w.y, w.z, 1.0f})) : mul(A, {w.x, w.y, w.z, 1.0f});
        const float* v = sc->vertices + 14 * (size_t)vi;
        return mul(&m[16 * (size_t)sc->vertex_actor[vi]], {v[0], v[1], v[2], 1.0f});
    }
};

// ------------------------------------------------------------------------------------------ rasteriser
// Canonical coverage: window = (ndc*0.5+0.5)*size, snapped to 8 sub-pixel bits with
// floor(x*256+0.5); 64-bit integer edge functions; sample at pixel centres; top-left rule in y-up window
// space: a pixel exactly on an edge a->b (interior on the left) is covered iff dy<0 || (dy==0 && dx<0).
struct RV { float x, y, z, w; };                        // clip-space vertex
inline int64_t snap(float ndc, int size) {
    float wv = (ndc * 0.5f + 0.5f) * (float)size;
    float s = std::floor(wv * 256.0f + 0.5f);
    if (!(s > -1073741824.0f)) s = -1073741824.0f;      // also catches NaN
    if (s > 1073741824.0f) s = 1073741824.0f;
    return (int64_t)s;
}
struct Setup {
    int64_t X[3], Y[3], area; int order[3];             // order[k] = source vertex placed at slot k (slots are CCW)
    int64_t bias[3];
    float zndc[3];
    int x0, x1, y0, y1; bool valid;
};
// Sample positions in 1/256 pixel from the pixel's lower-left corner.  Settings::conservativeRasterization == MSAA (reference default,
// Application.cpp:244-249, 673-678): GL_MULTISAMPLE on the 4-sample window framebuffer (main.cpp:256).  OpenGL 4.5 section 14.6.6: a fragment is
// produced for a pixel if ANY sample is covered (same point-sampling rule per sample, applied to the near/far-clipped triangle); without a
// `centroid` qualifier the fragment's inputs may be evaluated anywhere in the pixel — canonical choice: the pixel centre (what NVIDIA and AMD
// hardware do), extrapolated when the centre is outside.  The positions are the driver's (glGetMultisamplefv): a parameter, default the
// standard 4x pattern.
struct Samples { int n; int64_t x[4], y[4]; int64_t x_min, x_max, y_min, y_max; };
inline bool msaa_samples(const vct_frame_params* fp, Samples& ms) {
    if (fp->conservative_raster != VCT_RASTER_MSAA) return false;
    static const float standard4x[8] = {0.375f, 0.125f, 0.875f, 0.375f, 0.125f, 0.625f, 0.625f, 0.875f};
    bool given = false;
    for (int i = 0; i < 8; ++i) given = given || fp->msaa_samples[i] != 0.0f;
    const float* sp = given ? fp->msaa_samples : standard4x;
    ms.n = 4; ms.x_min = ms.y_min = 256; ms.x_max = ms.y_max = 0;
    for (int i = 0; i < 4; ++i) {
        ms.x[i] = std::min<int64_t>(255, std::max<int64_t>(0, std::lrint(sp[2 * i] * 256.0f))); ms.y[i] = std::min<int64_t>(255, std::max<int64_t>(0, std::lrint(sp[2 * i + 1] * 256.0f)));
        ms.x_min = std::min(ms.x_min, ms.x[i]); ms.x_max = std::max(ms.x_max, ms.x[i]); ms.y_min = std::min(ms.y_min, ms.y[i]); ms.y_max = std::max(ms.y_max, ms.y[i]);
    }
    return true;
}
inline Setup tri_setup(const RV v[3], int W, int H, bool cull_back, const Samples* ms = nullptr) {
    Setup s; s.valid = false;
    int64_t X[3], Y[3];
    for (int i = 0; i < 3; ++i) { X[i] = snap(v[i].x / v[i].w, W); Y[i] = snap(v[i].y / v[i].w, H); }
    int64_t area = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0]);
    if (area == 0) return s;
    s.order[0] = 0; s.order[1] = 1; s.order[2] = 2;
    if (area < 0) { if (cull_back) return s; s.order[1] = 2; s.order[2] = 1; area = -area; }
    for (int k = 0; k < 3; ++k) { s.X[k] = X[s.order[k]]; s.Y[k] = Y[s.order[k]]; s.zndc[k] = v[s.order[k]].z / v[s.order[k]].w; }
    s.area = area;
    for (int k = 0; k < 3; ++k) {                         // edge opposite slot k: a = slot k+1, b = slot k+2
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const int64_t dx = s.X[b] - s.X[a], dy = s.Y[b] - s.Y[a];
        s.bias[k] = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : -1;
    }
    const int64_t minx = std::min({s.X[0], s.X[1], s.X[2]}), maxx = std::max({s.X[0], s.X[1], s.X[2]});
    const int64_t miny = std::min({s.Y[0], s.Y[1], s.Y[2]}), maxy = std::max({s.Y[0], s.Y[1], s.Y[2]});
    // pixel i has centre 256 i + 128
    auto cdiv = [](int64_t a) { return (a >= 0) ? (a + 255) / 256 : -((-a) / 256); };      // ceil(a/256)
    auto fdiv = [](int64_t a) { return (a >= 0) ? a / 256 : -((-a + 255) / 256); };        // floor(a/256)
    // (multisampling: pixel i has samples at 256 i + offset; it can hold a covered one iff some offset lands inside [min, max])
    const int64_t ox_hi = ms ? ms->x_max : 128, ox_lo = ms ? ms->x_min : 128, oy_hi = ms ? ms->y_max : 128, oy_lo = ms ? ms->y_min : 128;
    s.x0 = (int)std::max<int64_t>(0, cdiv(minx - ox_hi)); s.x1 = (int)std::min<int64_t>(W - 1, fdiv(maxx - ox_lo));
    s.y0 = (int)std::max<int64_t>(0, cdiv(miny - oy_hi)); s.y1 = (int)std::min<int64_t>(H - 1, fdiv(maxy - oy_lo));
    s.valid = s.x0 <= s.x1 && s.y0 <= s.y1;
    return s;
}
// emit(px, py, l[3]) with l indexed by SOURCE vertex; raster-scan order (y outer, x inner)
template <class F>
inline void raster(const Setup& s, int ylo, int yhi, F&& emit) {
    const float fa = (float)s.area;
    for (int py = std::max(s.y0, ylo); py <= std::min(s.y1, yhi); ++py) {
        const int64_t Py = 256 * (int64_t)py + 128;
        for (int px = s.x0; px <= s.x1; ++px) {
            const int64_t Px = 256 * (int64_t)px + 128;
            int64_t E[3]; bool in = true;
            for (int k = 0; k < 3; ++k) {
                const int a = (k + 1) % 3, b = (k + 2) % 3;
                E[k] = (s.X[b] - s.X[a]) * (Py - s.Y[a]) - (s.Y[b] - s.Y[a]) * (Px - s.X[a]);
                if (E[k] + s.bias[k] < 0) { in = false; break; }
            }
            if (!in) continue;
            float l[3];
            for (int k = 0; k < 3; ++k) l[s.order[k]] = (float)E[k] / fa;
            emit(px, py, l);
        }
    }
}
inline float interp(const float l[3], float a0, float a1, float a2) { return (l[0] * a0 + l[1] * a1) + l[2] * a2; }
// Multisample variant: emit(px, py, l at the PIXEL CENTRE) for every pixel with at least one sample that is inside the triangle and inside
// -1 <= z <= 1 at that sample (z[] = ndc z per source vertex); raster-scan order.  The caller does not repeat the near/far test.
template <class F>
inline void raster_any_sample(const Setup& s, const Samples& ms, const float z[3], F&& emit) {
    const float fa = (float)s.area;
    auto edges = [&](int64_t Px, int64_t Py, int64_t E[3]) {
        bool in = true;
        for (int k = 0; k < 3; ++k) {
            const int a = (k + 1) % 3, b = (k + 2) % 3;
            E[k] = (s.X[b] - s.X[a]) * (Py - s.Y[a]) - (s.Y[b] - s.Y[a]) * (Px - s.X[a]);
            in = in && E[k] + s.bias[k] >= 0;
        }
        return in;
    };
    for (int py = s.y0; py <= s.y1; ++py)
        for (int px = s.x0; px <= s.x1; ++px) {
            bool any = false;
            for (int i = 0; i < ms.n && !any; ++i) {
                int64_t E[3];
                if (!edges(256 * (int64_t)px + ms.x[i], 256 * (int64_t)py + ms.y[i], E)) continue;
                float l[3];
                for (int k = 0; k < 3; ++k) l[s.order[k]] = (float)E[k] / fa;
                const float zs = interp(l, z[0], z[1], z[2]);
                any = !(zs < -1.0f || zs > 1.0f);
            }
            if (!any) continue;
            int64_t E[3]; float l[3];
            edges(256 * (int64_t)px + 128, 256 * (int64_t)py + 128, E);
            for (int k = 0; k < 3; ++k) l[s.order[k]] = (float)E[k] / fa;
            emit(px, py, l);
        }
}
inline V3 interp3(const float l[3], V3 a, V3 b, V3 c) {
    return {interp(l, a.x, b.x, c.x), interp(l, a.y, b.y, c.y), interp(l, a.z, b.z, c.z)};
}

// affine (orthographic) uv derivatives in window space -> rho^2 per triangle
inline float tri_rho2_affine(const RV v[3], const float uv[3][2], int W, int H, const Tex& t) {
    float x[3], y[3];
    for (int i = 0; i < 3; ++i) { x[i] = (v[i].x / v[i].w * 0.5f + 0.5f) * (float)W; y[i] = (v[i].y / v[i].w * 0.5f + 0.5f) * (float)H; }
    const float x1 = x[1] - x[0], x2 = x[2] - x[0], y1 = y[1] - y[0], y2 = y[2] - y[0];
    const float u1 = uv[1][0] - uv[0][0], u2 = uv[2][0] - uv[0][0], v1 = uv[1][1] - uv[0][1], v2 = uv[2][1] - uv[0][1];
    const float den = x1 * y2 - x2 * y1;
    const float dudx = (u1 * y2 - u2 * y1) / den, dudy = (u2 * x1 - u1 * x2) / den;
    const float dvdx = (v1 * y2 - v2 * y1) / den, dvdy = (v2 * x1 - v1 * x2) / den;
    const float ax = dudx * (float)t.w, bx = dvdx * (float)t.h, ay = dudy * (float)t.w, by = dvdy * (float)t.h;
    return maxf(ax * ax + bx * bx, ay * ay + by * by);
}

// homogeneous (perspective-correct) barycentrics of an unclipped clip-space triangle at an NDC point
struct Homog { float a[3], b[3], c[3]; };
inline Homog homog_setup(const RV v[3]) {
    Homog h;
    for (int i = 0; i < 3; ++i) {
        const RV &p = v[(i + 1) % 3], &q = v[(i + 2) % 3];
        h.a[i] = p.y * q.w - q.y * p.w;
        h.b[i] = q.x * p.w - p.x * q.w;
        h.c[i] = p.x * q.y - q.x * p.y;
    }
    return h;
}
inline void homog_eval(const Homog& h, float nx, float ny, float l[3]) {
    float b[3];
    for (int i = 0; i < 3; ++i) b[i] = (h.a[i] * nx + h.b[i] * ny) + h.c[i];
    const float s = (b[0] + b[1]) + b[2];
    for (int i = 0; i < 3; ++i) l[i] = b[i] / s;
}

inline bool has_alpha(const orc_scene* sc, int m) { return sc->materials[m].alpha_tex >= 0; }

// tests only (orc_alpha_trace): every alpha-tested fragment of the shadow-map and depth-prepass passes with what its fragment
// stage sees — 6 floats: pass (0 = reflectiveShadowMap.frag, 1 = dither.frag), material, u, v, rho^2 of the alpha map, kept (1)
// or discarded (0).  tests/test_glsl_ref.py replays the reference's two shaders on these records.  Each record carries its own
// decision, so the order in which the raster bands append does not matter.
struct AlphaTrace { float* rec = nullptr; long long cap = 0, n = 0; };
AlphaTrace g_alpha_trace;
inline void alpha_record(int pass, int material, float u, float v, float rho2, bool kept) {
    if (!g_alpha_trace.rec) return;
#pragma omp critical(orc_alpha_trace)
    {
        const long long i = g_alpha_trace.n++;
        if (i < g_alpha_trace.cap) {
            float* r = g_alpha_trace.rec + 6 * i;
            r[0] = (float)pass; r[1] = (float)material; r[2] = u; r[3] = v; r[4] = rho2; r[5] = kept ? 1.0f : 0.0f;
        }
    }
}

}  // namespace

// =================================================================================================== a0
// Application.cpp:212-233; simple.vert:15-22 (gl_Position = projection*view*model*pos, evaluated right to
// left as three mat*vec); reflectiveShadowMap.frag:34-38 (alpha discard); GL state: depth test LESS, back-face
// culling on (steady state after Application.cpp:285), depth = ndc.z*0.5+0.5 stored as float32, clear 1.
extern "C" void orc_set_literal_vertex_transforms(int on) { g_literal_vertex_transforms = on; }
extern "C" void orc_alpha_trace(float* rec, long long cap) { g_alpha_trace.rec = rec; g_alpha_trace.cap = cap; if (rec) g_alpha_trace.n = 0; }
extern "C" long long orc_alpha_trace_count(void) { return g_alpha_trace.n; }
extern "C" void orc_shadowmap(const orc_scene* sc, const vct_frame_params* fp, int S, float* depth) {
    Prepared P = prepare(sc, false);
    const ClipMats light_clip(sc, fp->lp, fp->lv);
    for (size_t i = 0; i < (size_t)S * S; ++i) depth[i] = 1.0f;
    int nb = 1;
#ifdef _OPENMP
    nb = omp_get_max_threads();
#endif
    const int band = (S + nb - 1) / nb;
#pragma omp parallel for schedule(static, 1)
    for (int bnd = 0; bnd < nb; ++bnd) {
        const int ylo = bnd * band, yhi = std::min(S, ylo + band) - 1;
        for (int t = 0; t < sc->n_tris; ++t) {
            RV cv[3]; float uv[3][2];
            for (int k = 0; k < 3; ++k) {
                const unsigned vi = sc->indices[3 * t + k];
                V3 w = P.wpos[vi];
                V4 c = light_clip.apply(sc, fp->lp, fp->lv, vi, w);
                cv[k] = {c.x, c.y, c.z, c.w};
                uv[k][0] = sc->vertices[14 * (size_t)vi + 6]; uv[k][1] = sc->vertices[14 * (size_t)vi + 7];
            }
            Setup s = tri_setup(cv, S, S, true);
            if (!s.valid || s.y1 < ylo || s.y0 > yhi) continue;
            const int m = sc->tri_material[t];
            const bool alpha = has_alpha(sc, m);
            float rho2 = 0.0f; const Tex* at = nullptr;
            if (alpha) { at = &P.tex[sc->materials[m].alpha_tex]; rho2 = tri_rho2_affine(cv, uv, S, S, *at); }
            const float z0 = cv[0].z / cv[0].w, z1 = cv[1].z / cv[1].w, z2 = cv[2].z / cv[2].w;
            raster(s, ylo, yhi, [&](int px, int py, const float l[3]) {
                const float z = interp(l, z0, z1, z2);
                if (z < -1.0f || z > 1.0f) return;                        // near/far clip
                if (alpha) {
                    const float u = interp(l, uv[0][0], uv[1][0], uv[2][0]), v = interp(l, uv[0][1], uv[1][1], uv[2][1]);
                    const bool kept = !(sample2d(*at, u, v, rho2).x < 0.1f);
                    alpha_record(0, m, u, v, rho2, kept);
                    if (!kept) return;
                }
                const float d = z * 0.5f + 0.5f;
                float& dst = depth[(size_t)py * S + px];
                if (d < dst) dst = d;
            });
        }
    }
}

// ============================================================================================== a1 + a2
namespace {
// voxelize.frag:111-139, sequential semantics of one imageAtomicRGBA8Avg call
inline uint32_t rgba8_avg_insert(uint32_t stored, float r, float g, float b) {
    const float vr = r * 255.0f, vg = g * 255.0f, vb = b * 255.0f, vw = 1.0f;
    if (stored == 0u)
        return (f2u_trunc(vw) & 255u) << 24 | (f2u_trunc(vb) & 255u) << 16 | (f2u_trunc(vg) & 255u) << 8 | (f2u_trunc(vr) & 255u);
    float rr = (float)(stored & 255u), rg = (float)((stored >> 8) & 255u), rb = (float)((stored >> 16) & 255u), rw = (float)(stored >> 24);
    rr *= rw; rg *= rw; rb *= rw;
    float cr = rr + vr, cg = rg + vg, cb = rb + vb; const float cw = rw + vw;
    cr /= cw; cg /= cw; cb /= cw;
    return (f2u_trunc(cw) & 255u) << 24 | (f2u_trunc(cb) & 255u) << 16 | (f2u_trunc(cg) & 255u) << 8 | (f2u_trunc(cr) & 255u);
}

struct VoxAxis { const float* mvp; int axis; };
// voxelize.geom:26-59
inline VoxAxis pick_axis(const vct_frame_params* fp, V3 n0, V3 n1, V3 n2) {
    V3 f = normalize((n0 + n1) + n2);
    const float ax = std::fabs(f.x), ay = std::fabs(f.y), az = std::fabs(f.z);
    int axis;
    if (ax > ay && ax > az) axis = 0; else if (ay > ax && ay > az) axis = 1; else axis = 2;
    if (fp->axis_override >= 0 && fp->axis_override <= 2) axis = fp->axis_override;
    return {axis == 0 ? fp->mvp_x : axis == 1 ? fp->mvp_y : fp->mvp_z, axis};
}
// voxelize.frag:79-108 (getVoxelPosition(ivec3 size)); returns size*unit
inline V3 frag_voxel_position(V3 ndc, int axis, int size, const vct_frame_params* fp, const uint16_t* warpmap, bool occupancy) {
    V3 u = {(ndc.x + 1.0f) * 0.5f, (ndc.y + 1.0f) * 0.5f, (ndc.z + 1.0f) * 0.5f};
    if (axis == 0) u = {1.0f - u.z, u.y, u.x};
    else if (axis == 1) u = {u.x, 1.0f - u.z, u.y};
    u.z = 1.0f - u.z;
    if (fp->warp_voxels) u = voxel_warp(u, voxel_linear_position(eye3(fp), fp));
    else if (fp->warp_texture && !occupancy && warpmap) u = warp_sample(warpmap, u);
    return {(float)size * u.x, (float)size * u.y, (float)size * u.z};
}
inline bool to_index(V3 p, int D, int idx[3]) {       // ivec3(vec3): trunc toward zero; OOB image access = no-op
    const float c[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; ++i) {
        if (!(c[i] > -1.0f) || !(c[i] < (float)D)) return false;     // also NaN
        idx[i] = (int)c[i];
    }
    return true;
}
}  // namespace

extern "C" unsigned orc_rgba8_avg(unsigned stored, float r, float g, float b) { return rgba8_avg_insert(stored, r, g, b); }
extern "C" unsigned orc_pack_unorm4x8(float r, float g, float b, float a) { return pack_unorm({r, g, b, a}); }

extern "C" void orc_occupancy(const orc_scene* sc, const vct_frame_params* fp, unsigned* occ) {
    const int D = VCT_WARP_DIM;
    Prepared P = prepare(sc, false);
    std::memset(occ, 0, sizeof(unsigned) * D * D * D);
    Samples ms; const bool msaa = msaa_samples(fp, ms);
    for (int t = 0; t < sc->n_tris; ++t) {
        const unsigned* ix = sc->indices + 3 * (size_t)t;
        VoxAxis va = pick_axis(fp, P.wnrm[ix[0]], P.wnrm[ix[1]], P.wnrm[ix[2]]);
        RV cv[3];
        for (int k = 0; k < 3; ++k) { V3 w = P.wpos[ix[k]]; V4 c = mul(va.mvp, {w.x, w.y, w.z, 1.0f}); cv[k] = {c.x, c.y, c.z, c.w}; }
        Setup s = tri_setup(cv, D, D, false, msaa ? &ms : nullptr);
        if (!s.valid) continue;
        auto frag = [&](int, int, const float l[3]) {
            V3 ndc = {interp(l, cv[0].x, cv[1].x, cv[2].x), interp(l, cv[0].y, cv[1].y, cv[2].y), interp(l, cv[0].z, cv[1].z, cv[2].z)};
            if (!msaa && (ndc.z < -1.0f || ndc.z > 1.0f)) return;
            int idx[3];
            if (!to_index(frag_voxel_position(ndc, va.axis, D, fp, nullptr, true), D, idx)) return;
            occ[((size_t)idx[2] * D + idx[1]) * D + idx[0]] |= 1u;
        };
        const float zs[3] = {cv[0].z, cv[1].z, cv[2].z};
        if (msaa) raster_any_sample(s, ms, zs, frag); else raster(s, 0, D - 1, frag);
    }
}

// Canonical fragment order: triangles in draw order, fragments of a triangle in raster-scan order.
static void voxelize_impl(const orc_scene* sc, const vct_frame_params* fp, int D, const float* shadow, int S,
                          const unsigned short* warpmap, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                          float* frag_rec, long long frag_cap, long long* frag_count) {
    Prepared P = prepare(sc, false);
    if (color) {
        std::memset(color, 0, sizeof(unsigned) * (size_t)D * D * D);   // glClearTexImage, Application.cpp:686-687
        std::memset(normal, 0, sizeof(unsigned) * (size_t)D * D * D);
    }
    unsigned total = 0;
    Samples ms; const bool msaa = msaa_samples(fp, ms);
    // Settings::voxelizeMultiplier: glViewport(0, 0, m * voxelDim, m * voxelDim) (Application.cpp:668; float -> GLsizei truncates); the voxel
    // index still comes from the interpolated position times imageSize (voxelize.frag:79-108)
    const int V = fp->voxelize_multiplier > 0.0f ? (int)(fp->voxelize_multiplier * (float)D) : D;
    for (int t = 0; t < sc->n_tris; ++t) {
        const unsigned* ix = sc->indices + 3 * (size_t)t;
        V3 n[3] = {P.wnrm[ix[0]], P.wnrm[ix[1]], P.wnrm[ix[2]]}, w[3] = {P.wpos[ix[0]], P.wpos[ix[1]], P.wpos[ix[2]]};
        VoxAxis va = pick_axis(fp, n[0], n[1], n[2]);
        RV cv[3]; float uv[3][2];
        for (int k = 0; k < 3; ++k) {
            V4 c = mul(va.mvp, {w[k].x, w[k].y, w[k].z, 1.0f}); cv[k] = {c.x, c.y, c.z, c.w};
            uv[k][0] = sc->vertices[14 * (size_t)ix[k] + 6]; uv[k][1] = sc->vertices[14 * (size_t)ix[k] + 7];
        }
        Setup s = tri_setup(cv, V, V, false, msaa ? &ms : nullptr);
        if (!s.valid) continue;
        const vct_material& mat = sc->materials[sc->tri_material[t]];
        const Tex* dt = mat.diffuse_tex >= 0 ? &P.tex[mat.diffuse_tex] : nullptr;
        const float rho2 = dt ? tri_rho2_affine(cv, uv, V, V, *dt) : 0.0f;
        auto frag = [&](int, int, const float l[3]) {
            V3 ndc = {interp(l, cv[0].x, cv[1].x, cv[2].x), interp(l, cv[0].y, cv[1].y, cv[2].y), interp(l, cv[0].z, cv[1].z, cv[2].z)};
            if (!msaa && (ndc.z < -1.0f || ndc.z > 1.0f)) return;       // near/far clip of the ortho volume (multisampling: per sample, in the rasteriser)
            total++;                                                      // voxelize.frag:195
            V3 wp = interp3(l, w[0], w[1], w[2]);
            V3 nn = interp3(l, n[0], n[1], n[2]);
            const float u = interp(l, uv[0][0], uv[1][0], uv[2][0]), v = interp(l, uv[0][1], uv[1][1], uv[2][1]);
            if (frag_rec && (long long)total <= frag_cap) {               // fragment-stage inputs (GS_OUT block + sampler state)
                float* r = frag_rec + 16 * (size_t)(total - 1);
                r[0] = ndc.x; r[1] = ndc.y; r[2] = ndc.z; r[3] = wp.x; r[4] = wp.y; r[5] = wp.z; r[6] = nn.x; r[7] = nn.y; r[8] = nn.z;
                r[9] = u; r[10] = v; r[11] = (float)va.axis; r[12] = rho2; r[13] = (float)mat.diffuse_tex; r[14] = r[15] = 0.0f;
            }
            if (!color) return;                                           // record only: rasteriser without the fragment stage
            V3 col = {0, 0, 0};
            if (dt) { V4 a = sample2d(*dt, u, v, rho2); col = {a.x, a.y, a.z}; }
            V3 N = normalize(nn);
            V3 nenc = {(N.x + 1.0f) * 0.5f, (N.y + 1.0f) * 0.5f, (N.z + 1.0f) * 0.5f};
            if (fp->voxelize_lighting) {                                  // voxelize.frag:200-226
                V3 fin = {0, 0, 0};
                for (int i = 0; i < sc->n_lights; ++i) {
                    const vct_light& L = sc->lights[i];
                    if (!L.enabled) continue;
                    V3 lc = {L.color[0], L.color[1], L.color[2]}, lp = {L.position[0], L.position[1], L.position[2]};
                    V3 lit = {0, 0, 0};
                    if (L.type == 0u) {
                        const float ndl = maxf(0.0f, dot(N, normalize(lp - wp)));
                        lit = ((col * L.intensity) * lc) * ndl;
                        const float e0 = 0.75f * L.range, e1 = L.range, dist = length(lp - wp);
                        const float tt = clampf((dist - e0) / (e1 - e0), 0.0f, 1.0f);
                        lit = lit * (1.0f - tt * tt * (3.0f - 2.0f * tt));
                    } else if (L.type == 1u) {
                        V3 nd = {-L.direction[0], -L.direction[1], -L.direction[2]};
                        const float ndl = maxf(0.0f, dot(N, normalize(nd)));
                        lit = ((col * lc) * L.intensity) * ndl;
                    }
                    if (L.shadow_caster) {
                        const float sf = 1.0f - calc_shadow_factor(shadow, S, mul(fp->ls, {wp.x, wp.y, wp.z, 1.0f}));
                        lit = lit * sf;
                    }
                    fin = fin + lit;
                }
                col = {clampf(fin.x, 0.0f, 1.0f), clampf(fin.y, 0.0f, 1.0f), clampf(fin.z, 0.0f, 1.0f)};
            }
            int idx[3];
            if (!to_index(frag_voxel_position(ndc, va.axis, D, fp, warpmap, false), D, idx)) return;
            const size_t o = ((size_t)idx[2] * D + idx[1]) * D + idx[0];
            if (fp->voxelize_atomic_max) {                                // voxelize.frag:271-274
                color[o] = std::max(color[o], pack_unorm({col.x, col.y, col.z, 1.0f}));
                normal[o] = std::max(normal[o], pack_unorm({nenc.x, nenc.y, nenc.z, 1.0f}));
            } else {                                                       // :275-278
                color[o] = rgba8_avg_insert(color[o], col.x, col.y, col.z);
                normal[o] = rgba8_avg_insert(normal[o], nenc.x, nenc.y, nenc.z);
            }
        };
        const float zs[3] = {cv[0].z, cv[1].z, cv[2].z};
        if (msaa) raster_any_sample(s, ms, zs, frag); else raster(s, 0, V - 1, frag);
    }
    if (info) { info->total_fragments = total; info->unique_voxels = 0; info->max_fragments_per_voxel = 0; }
    if (frag_count) *frag_count = total;
}
extern "C" void orc_voxelize(const orc_scene* sc, const vct_frame_params* fp, int D, const float* shadow, int S,
                             const unsigned short* warpmap, unsigned* color, unsigned* normal, vct_voxelize_info* info) {
    voxelize_impl(sc, fp, D, shadow, S, warpmap, color, normal, info, nullptr, 0, nullptr);
}
// Same pass, additionally recording what the rasteriser hands the fragment stage for every fragment, in canonical order:
// 16 floats = GS_OUT{position(ndc) 3, worldPosition 3, normal 3, texcoord 2, axis} + rho2 of the diffuse lookup + diffuse
// texture id (tests/test_glsl_ref.py replays them through the reference's voxelize.frag compiled as C++).
extern "C" void orc_voxelize_trace(const orc_scene* sc, const vct_frame_params* fp, int D, const float* shadow, int S,
                                   const unsigned short* warpmap, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                                   float* frag_rec, long long frag_cap, long long* frag_count) {
    voxelize_impl(sc, fp, D, shadow, S, warpmap, color, normal, info, frag_rec, frag_cap, frag_count);
}

// ============================================================================= a1'/a2' tessellation voxeliser (N4)
// The reference's DEFAULT voxeliser (Settings::voxelizeTesselation = true, src/Application.cpp:585-665): every triangle is a
// patch; testTesselation.tesc picks tessellation levels from its size in voxels; the fixed-function tessellator
// (triangles, equal_spacing, point_mode) emits one point per distinct vertex of the subdivision; testTesselation.tese stores
// the UNLIT diffuse texel at each point's voxel (atomicMax by default).  Nothing is rasterised.
// Canonical fixed function (OpenGL 4.5 §11.2.2, GL_MAX_TESS_GEN_LEVEL = 64):
//   * a patch with an outer level <= 0 (or NaN) is discarded;
//   * equal_spacing: every level is clamped to [1, 64] and rounded up to an integer; inner level 1 with an outer level > 1 counts
//     as 2; inner = outer = 1 yields the three corners only;
//   * concentric triangles: ring j >= 1 of inner level n is the triangle with corners (1-4j/(3n), 2j/(3n), 2j/(3n)) (and cyclic)
//     whose edges carry m = n-2j segments (m = 0: the single centre point); ring 0 is the patch itself with outer[0] segments on
//     the u = 0 edge, outer[1] on v = 0, outer[2] on w = 0;
//   * every barycentric coordinate is ONE fp32 division of two exactly representable integers (numerator/denominator below), so the
//     three coordinates of a point are symmetric under relabelling;
//   * emission order (only matters for the running average): ring 0 corners (1,0,0), (0,1,0), (0,0,1), then the interior points of
//     the edges w = 0 (from u=1 to v=1), u = 0 (v=1 to w=1), v = 0 (w=1 to u=1); then rings 1, 2, ... each walked A->B->C->A.
namespace {
inline float glsl_max(float x, float y) { return x < y ? y : x; }          // GLSL max(x, y): y if x < y, otherwise x
struct TessLevels { float inner, outer[3]; };
// testTesselation.tesc:32-78
inline TessLevels tess_control(const V3 w[3], const vct_frame_params* fp, float voxelDim) {
    TessLevels t = {1.0f, {1.0f, 1.0f, 1.0f}};
    int vox[3][3];
    for (int k = 0; k < 3; ++k) {
        V3 vp = voxel_linear_position(w[k], fp);                           // tcVoxelPosition
        const float c[3] = {voxelDim * vp.x, voxelDim * vp.y, voxelDim * vp.z};
        for (int a = 0; a < 3; ++a) vox[k][a] = (int)c[a];                 // ivec3(vec3): truncation (values are finite for finite input)
    }
    bool same = true;
    for (int a = 0; a < 3; ++a) same = same && vox[0][a] == vox[1][a] && vox[0][a] == vox[2][a];
    if (same) { t.inner = 0.0f; t.outer[0] = t.outer[1] = t.outer[2] = 0.0f; return t; }
    const V3 a = w[0], b = w[1], c = w[2];
    const V3 A = c - b, B = c - a, C = b - a;
    const float lx = length(A), ly = length(B), lz = length(C);
    const float s = ((lx + ly) + lz) * 0.5f;
    const float area = std::sqrt(((s * (s - lx)) * (s - ly)) * (s - lz));
    const float ax = (2.0f * area) / lx, ay = (2.0f * area) / ly, az = (2.0f * area) / lz;
    const float max_alt = glsl_max(ax, glsl_max(ay, az));
    const V3 vs = {(fp->voxel_max[0] - fp->voxel_min[0]) / voxelDim, (fp->voxel_max[1] - fp->voxel_min[1]) / voxelDim, (fp->voxel_max[2] - fp->voxel_min[2]) / voxelDim};
    const float dx = std::fabs(length(normalize(A) * vs)), dy = std::fabs(length(normalize(B) * vs)), dz = std::fabs(length(normalize(C) * vs));
    const float ox = glsl_max(1.0f, lx / dx), oy = glsl_max(1.0f, ly / dy), oz = glsl_max(1.0f, lz / dz);
    t.inner = glsl_max(1.0f, max_alt / vs.x);
    t.outer[0] = oz; t.outer[1] = ox; t.outer[2] = oy;
    return t;
}
inline int tess_round(float level) {                                        // equal_spacing: clamp to [1, 64], round up
    if (!(level > 1.0f)) return 1;
    if (level >= 64.0f) return 64;
    return (int)std::ceil(level);
}
// calls emit(u, v, w) for every distinct vertex of the subdivision, in the canonical order
template <class F>
inline void tess_points(const TessLevels& t, F&& emit) {
    for (int k = 0; k < 3; ++k) if (!(t.outer[k] > 0.0f)) return;          // discarded patch (also NaN)
    int n = tess_round(t.inner);
    const int o[3] = {tess_round(t.outer[0]), tess_round(t.outer[1]), tess_round(t.outer[2])};
    emit(1.0f, 0.0f, 0.0f); emit(0.0f, 1.0f, 0.0f); emit(0.0f, 0.0f, 1.0f);
    if (n == 1 && o[0] == 1 && o[1] == 1 && o[2] == 1) return;
    if (n == 1) n = 2;
    for (int i = 1; i < o[2]; ++i) { const float q = (float)i / (float)o[2], r = (float)(o[2] - i) / (float)o[2]; emit(r, q, 0.0f); }   // w = 0: u=1 -> v=1
    for (int i = 1; i < o[0]; ++i) { const float q = (float)i / (float)o[0], r = (float)(o[0] - i) / (float)o[0]; emit(0.0f, r, q); }   // u = 0: v=1 -> w=1
    for (int i = 1; i < o[1]; ++i) { const float q = (float)i / (float)o[1], r = (float)(o[1] - i) / (float)o[1]; emit(q, 0.0f, r); }   // v = 0: w=1 -> u=1
    for (int j = 1; n - 2 * j >= 0; ++j) {
        const int m = n - 2 * j;
        if (m == 0) { const float third = 1.0f / 3.0f; emit(third, third, third); break; }
        const float den = (float)(3 * n * m), small = (float)(2 * j) / (float)(3 * n);
        for (int e = 0; e < 3; ++e)
            for (int i = 0; i < m; ++i) {
                const float big = (float)((3 * n - 4 * j) * (m - i) + 2 * j * i) / den;      // coordinate falling from the ring corner
                const float rise = (float)(2 * j * (m - i) + (3 * n - 4 * j) * i) / den;     // coordinate rising towards the next corner
                if (e == 0) emit(big, rise, small); else if (e == 1) emit(small, big, rise); else emit(rise, small, big);
            }
    }
}
}  // namespace

// testTesselation.tese:82-147 per point; voxelStore (:60-80).  frag_rec (optional): 8 floats per point = triangle id, u, v, w, pad.
static void voxelize_tess_impl(const orc_scene* sc, const vct_frame_params* fp, int D, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                               float* rec, long long rec_cap, long long* rec_count) {
    Prepared P = prepare(sc, false);
    std::memset(color, 0, sizeof(unsigned) * (size_t)D * D * D);           // glClearTexImage, Application.cpp:626-627
    std::memset(normal, 0, sizeof(unsigned) * (size_t)D * D * D);
    const float voxelDim = (float)D;                                       // uniform float voxelDim
    long long count = 0;
    for (int t = 0; t < sc->n_tris; ++t) {
        const unsigned* ix = sc->indices + 3 * (size_t)t;
        const V3 w[3] = {P.wpos[ix[0]], P.wpos[ix[1]], P.wpos[ix[2]]}, n[3] = {P.wnrm[ix[0]], P.wnrm[ix[1]], P.wnrm[ix[2]]};
        float uv[3][2];
        for (int k = 0; k < 3; ++k) { uv[k][0] = sc->vertices[14 * (size_t)ix[k] + 6]; uv[k][1] = sc->vertices[14 * (size_t)ix[k] + 7]; }
        const TessLevels tl = tess_control(w, fp, voxelDim);
        const vct_material& mat = sc->materials[sc->tri_material[t]];
        const Tex* dt = mat.diffuse_tex >= 0 ? &P.tex[mat.diffuse_tex] : nullptr;
        tess_points(tl, [&](float u, float v, float ww) {
            if (rec && count < rec_cap) { float* r = rec + 8 * (size_t)count; r[0] = (float)t; r[1] = u; r[2] = v; r[3] = ww; r[4] = r[5] = r[6] = r[7] = 0.0f; }
            ++count;
            if (!color) return;
            const V3 pos = (w[0] * u + w[1] * v) + w[2] * ww;               // gl_TessCoord.x * tcPosition[0] + .y * [1] + .z * [2]
            const V3 nn = (n[0] * u + n[1] * v) + n[2] * ww;
            const float tu = (u * uv[0][0] + v * uv[1][0]) + ww * uv[2][0], tv = (u * uv[0][1] + v * uv[1][1]) + ww * uv[2][1];
            V4 col = {0, 0, 0, 1};
            if (dt) col = sample2d(*dt, tu, tv, 0.0f);                     // no derivatives in a TES: base level, magnification -> NEAREST
            // (the branch for patches inside one voxel, :128-131, is unreachable: those patches have tessellation level 0)
            V3 vp = tese_voxel_position(pos, fp);                          // voxelIndex(position, int(voxelDim), ..., false)
            vp = {voxelDim * vp.x, voxelDim * vp.y, voxelDim * vp.z};      // (int(voxelDim) * pos with an integral voxelDim)
            int idx[3];
            if (!to_index(vp, D, idx)) return;
            const size_t o = ((size_t)idx[2] * D + idx[1]) * D + idx[0];
            const V3 N = normalize(nn);
            const V3 nenc = {N.x * 0.5f + 0.5f, N.y * 0.5f + 0.5f, N.z * 0.5f + 0.5f};
            if (fp->voxelize_atomic_max) {                                  // :69-74 (texture alpha is part of the packed word)
                color[o] = std::max(color[o], pack_unorm(col));
                normal[o] = std::max(normal[o], pack_unorm({nenc.x, nenc.y, nenc.z, 1.0f}));
            } else {                                                        // :75-78
                color[o] = rgba8_avg_insert(color[o], col.x, col.y, col.z);
                normal[o] = rgba8_avg_insert(normal[o], nenc.x, nenc.y, nenc.z);
            }
        });
    }
    if (info) { info->total_fragments = 0; info->unique_voxels = 0; info->max_fragments_per_voxel = 0; }   // the TES counts nothing
    if (rec_count) *rec_count = count;
}
// vertex stage of simpleTesselated.vert / voxelize.vert for every vertex: world position and normalMatrix * normal (3 floats each)
extern "C" void orc_world_vertices(const orc_scene* sc, float* wpos, float* wnrm) {
    Prepared P = prepare(sc, false);
    for (int i = 0; i < sc->n_vertices; ++i) {
        wpos[3 * i] = P.wpos[i].x; wpos[3 * i + 1] = P.wpos[i].y; wpos[3 * i + 2] = P.wpos[i].z;
        wnrm[3 * i] = P.wnrm[i].x; wnrm[3 * i + 1] = P.wnrm[i].y; wnrm[3 * i + 2] = P.wnrm[i].z;
    }
}
// The vertex / geometry stages as the passes above evaluate them, for tests/test_glsl_ref.py (any output may be NULL):
//   voxel  n_tris x 13     : dominant axis, then gl_Position (x, y, z, w) of the three vertices      voxelize.vert + voxelize.geom
//   light  n_vertices x 4  : gl_Position of the shadow-map pass                                      simple.vert with lp, lv
//   cam    n_vertices x 4  : gl_Position of the depth prepass / phong pass                           simple.vert / phong.vert
//   phong  n_vertices x 16 : fragPosition 3, fragNormal 3, lightFragPos 4, TBN columns T 3 and B 3   phong.vert:36-58
extern "C" void orc_vertex_stage(const orc_scene* sc, const vct_frame_params* fp, float* voxel, float* light, float* cam, float* phong) {
    Prepared P = prepare(sc, true);
    const ClipMats light_clip(sc, fp->lp, fp->lv), cam_clip(sc, fp->projection, fp->view), ls_clip(sc, fp->ls, nullptr);
    if (voxel)
        for (int t = 0; t < sc->n_tris; ++t) {
            const unsigned* ix = sc->indices + 3 * (size_t)t;
            VoxAxis va = pick_axis(fp, P.wnrm[ix[0]], P.wnrm[ix[1]], P.wnrm[ix[2]]);
            float* o = voxel + 13 * (size_t)t;
            o[0] = (float)va.axis;
            for (int k = 0; k < 3; ++k) { V3 w = P.wpos[ix[k]]; V4 c = mul(va.mvp, {w.x, w.y, w.z, 1.0f}); o[1 + 4 * k] = c.x; o[2 + 4 * k] = c.y; o[3 + 4 * k] = c.z; o[4 + 4 * k] = c.w; }
        }
    for (int i = 0; i < sc->n_vertices; ++i) {
        const V3 w = P.wpos[i];
        if (light) { V4 c = light_clip.apply(sc, fp->lp, fp->lv, (unsigned)i, w); float* o = light + 4 * (size_t)i; o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
        if (cam) { V4 c = cam_clip.apply(sc, fp->projection, fp->view, (unsigned)i, w); float* o = cam + 4 * (size_t)i; o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; }
        if (phong) {
            V4 l = ls_clip.apply(sc, fp->ls, nullptr, (unsigned)i, w);
            const float v[16] = {w.x, w.y, w.z, P.wnrm[i].x, P.wnrm[i].y, P.wnrm[i].z, l.x, l.y, l.z, l.w, P.T[i].x, P.T[i].y, P.T[i].z, P.B[i].x, P.B[i].y, P.B[i].z};
            std::memcpy(phong + 16 * (size_t)i, v, sizeof v);
        }
    }
}
extern "C" void orc_voxelize_tess(const orc_scene* sc, const vct_frame_params* fp, int D, unsigned* color, unsigned* normal, vct_voxelize_info* info) {
    voxelize_tess_impl(sc, fp, D, color, normal, info, nullptr, 0, nullptr);
}
// the same pass, additionally recording (triangle, u, v, w) of every emitted point; with color == NULL only the points are produced
extern "C" void orc_voxelize_tess_trace(const orc_scene* sc, const vct_frame_params* fp, int D, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                                        float* rec, long long rec_cap, long long* rec_count) {
    voxelize_tess_impl(sc, fp, D, color, normal, info, rec, rec_cap, rec_count);
}
// fixed function + TCS of one triangle given by world positions (tests): levels[4] = inner, outer[0..2]; returns the point count
extern "C" long long orc_tess_patch(const float* wpos9, const vct_frame_params* fp, int D, float* levels4, float* uvw, long long cap) {
    const V3 w[3] = {{wpos9[0], wpos9[1], wpos9[2]}, {wpos9[3], wpos9[4], wpos9[5]}, {wpos9[6], wpos9[7], wpos9[8]}};
    const TessLevels tl = tess_control(w, fp, (float)D);
    if (levels4) { levels4[0] = tl.inner; levels4[1] = tl.outer[0]; levels4[2] = tl.outer[1]; levels4[3] = tl.outer[2]; }
    long long n = 0;
    tess_points(tl, [&](float u, float v, float ww) { if (uvw && n < cap) { uvw[3 * n] = u; uvw[3 * n + 1] = v; uvw[3 * n + 2] = ww; } ++n; });
    return n;
}

// =================================================================================================== a3
// transferVoxels.comp:29-70 (RGBA8 build) + radiance clear unless temporal (Application.cpp:762-764)
extern "C" void orc_transfer(const vct_frame_params* fp, int D, unsigned* color, unsigned* radiance, vct_voxelize_info* info) {
    const size_t n = (size_t)D * D * D;
    const bool temporal = fp->temporal_filter_radiance != 0;
    if (!temporal) std::memset(radiance, 0, n * 4);
    unsigned uniq = 0, maxf_ = 0;
#pragma omp parallel for schedule(static) reduction(+ : uniq) reduction(max : maxf_)
    for (long long i = 0; i < (long long)n; ++i) {
        V4 c = unpack_unorm(color[i]);
        if (c.w > 0.0f) {
            uniq++;
            maxf_ = std::max(maxf_, f2u_trunc(255.0f * c.w));
            if (fp->voxel_set_opacity > 0.0f) c.w = fp->voxel_set_opacity;
            color[i] = pack_unorm(c);
        }
        c.x = c.y = c.z = 0.0f;
        if (temporal) {                                   // mix(prev, c, 1 - decay)
            V4 p = unpack_unorm(radiance[i]);
            const float a = 1.0f - fp->temporal_decay;
            c = {mixf(p.x, c.x, a), mixf(p.y, c.y, a), mixf(p.z, c.z, a), mixf(p.w, c.w, a)};
        }
        if (temporal || c.w > 0.0f) radiance[i] = pack_unorm(c);
    }
    if (info) { info->unique_voxels = uniq; info->max_fragments_per_voxel = maxf_; }
}

// setVoxelOpacity.comp:18-35 (dead variant)
extern "C" void orc_set_voxel_opacity(int D, float opacity, unsigned* color, unsigned* radiance, vct_voxelize_info* info) {
    const size_t n = (size_t)D * D * D; unsigned uniq = 0, mx = 0;
    for (size_t i = 0; i < n; ++i) {
        V4 c = unpack_unorm(color[i]);
        if (c.w > 0.0f) {
            uniq++; mx = std::max(mx, f2u_trunc(255.0f * c.w));
            if (opacity > 0.0f) c.w = opacity;
            color[i] = pack_unorm(c);
            radiance[i] = pack_unorm({0, 0, 0, c.w});
        }
    }
    if (info) { info->unique_voxels = uniq; info->max_fragments_per_voxel = mx; }
}
// temporalRadianceFilter.comp:9-19 (dead variant)
extern "C" void orc_temporal_radiance_filter(int D, float decay, unsigned* vol) {
    const size_t n = (size_t)D * D * D;
    for (size_t i = 0; i < n; ++i) { V4 c = unpack_unorm(vol[i]); vol[i] = pack_unorm({c.x * decay, c.y * decay, c.z * decay, c.w * decay}); }
}
// normalizeVoxels.comp:19-42 (dead variant, RGBA16F volumes)
extern "C" void orc_normalize_voxels_f16(int D, float opacity, unsigned short* col, unsigned short* nrm, unsigned* radiance, vct_voxelize_info* info) {
    const size_t n = (size_t)D * D * D; unsigned uniq = 0, mx = 0;
    for (size_t i = 0; i < n; ++i) {
        float c[4]; for (int k = 0; k < 4; ++k) c[k] = h2f(col[4 * i + k]);
        if (c[3] > 0.0f) {
            uniq++; mx = std::max(mx, f2u_trunc(c[3]));
            const float a = c[3]; for (int k = 0; k < 4; ++k) c[k] = c[k] / a;
            if (opacity > 0.0f) c[3] = opacity;
            for (int k = 0; k < 4; ++k) col[4 * i + k] = f2h(c[k]);
            radiance[i] = pack_unorm({0, 0, 0, c[3]});
        }
        float m[4]; for (int k = 0; k < 4; ++k) m[k] = h2f(nrm[4 * i + k]);
        if (m[3] > 0.0f) { const float a = m[3]; for (int k = 0; k < 4; ++k) nrm[4 * i + k] = f2h(m[k] / a); }
    }
    if (info) { info->unique_voxels = uniq; info->max_fragments_per_voxel = mx; }
}

// =================================================================================================== a5
// injectRadiance.comp:40-103.  texture() in a compute stage has no derivatives -> base level, LINEAR.
extern "C" void orc_inject(const vct_frame_params* fp, int D, const unsigned* color, const unsigned* normal, const float* shadow,
                           int S, const unsigned short* warpmap, const float light_pos[3], const float light_int[3], unsigned* radiance) {
    // Several texels may hit one voxel; every writer stores the same word in the supported modes, so any
    // order gives the same result.
#pragma omp parallel for schedule(static)
    for (int y = 0; y < S; ++y)
        for (int x = 0; x < S; ++x) {
            const float tu = (float)x / (float)S, tv = (float)y / (float)S;
            const float d = shadow_linear(shadow, S, tu, tv, 0, 0);
            const float nx = tu * 2.0f - 1.0f, ny = tv * 2.0f - 1.0f, nz = d * 2.0f - 1.0f;
            V4 w = mul(fp->ls_inverse, {nx, ny, nz, 1.0f});
            V3 vp = get_voxel_position({w.x, w.y, w.z}, fp, warpmap);
            vp = {(float)D * vp.x, (float)D * vp.y, (float)D * vp.z};
            int idx[3];
            if (!to_index(vp, D, idx)) continue;
            const size_t o = ((size_t)idx[2] * D + idx[1]) * D + idx[0];
            V4 c = unpack_unorm(color[o]);
            if (fp->radiance_lighting) {                                // injectRadiance.comp:77-97 (temporal branch excluded)
                V4 n4 = unpack_unorm(normal[o]);
                V3 n = {2.0f * n4.x - 1.0f, 2.0f * n4.y - 1.0f, 2.0f * n4.z - 1.0f};
                V3 lpv = get_voxel_position({light_pos[0], light_pos[1], light_pos[2]}, fp, warpmap);
                lpv = {(float)D * lpv.x, (float)D * lpv.y, (float)D * lpv.z};
                V3 lv = normalize(lpv - v3((float)idx[0], (float)idx[1], (float)idx[2]));
                const float diff = maxf(dot(n, lv), 0.0f);
                c.x = (diff * light_int[0]) * c.x; c.y = (diff * light_int[1]) * c.y; c.z = (diff * light_int[2]) * c.z;
            }
            radiance[o] = pack_unorm(c);
        }
}

// =================================================================================================== a4
// voxelFillHoles.comp:8-36, then glCopyImageSubData back (Application.cpp:867-872)
extern "C" void orc_fill_holes(int D, unsigned* radiance) {
    const size_t n = (size_t)D * D * D;
    std::vector<unsigned> tmp(n);
#pragma omp parallel for schedule(static)
    for (int z = 0; z < D; ++z)
        for (int y = 0; y < D; ++y)
            for (int x = 0; x < D; ++x) {
                const size_t o = ((size_t)z * D + y) * D + x;
                V4 cur = unpack_unorm(radiance[o]);
                if (cur.w == 0.0f) {
                    float count = 0.0f;
                    for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) for (int k = -1; k <= 1; ++k) {
                        const int xx = x + i, yy = y + j, zz = z + k;          // threadId + ivec3(i,j,k)
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= D || yy >= D || zz >= D) continue;   // OOB load = 0
                        V4 v = unpack_unorm(radiance[((size_t)zz * D + yy) * D + xx]);
                        if (v.w > 0.0f) { cur = {cur.x + v.x, cur.y + v.y, cur.z + v.z, cur.w + v.w}; count += 1.0f; }
                    }
                    if (count > 0.0f) cur = {cur.x / count, cur.y / count, cur.z / count, cur.w / count};
                }
                tmp[o] = pack_unorm(cur);
            }
    std::memcpy(radiance, tmp.data(), n * 4);
}

// =================================================================================================== a6
// filterRadiance.comp:14-65
extern "C" void orc_mip(int Ds, const unsigned* src, unsigned* dst, int mode) {
    const int Dd = Ds / 2;
    auto load = [&](int x, int y, int z) -> V4 {
        if (x < 0 || y < 0 || z < 0 || x >= Ds || y >= Ds || z >= Ds) return {0, 0, 0, 0};
        return unpack_unorm(src[((size_t)z * Ds + y) * Ds + x]);
    };
    auto add = [](V4 a, V4 b) -> V4 { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; };
#pragma omp parallel for schedule(static)
    for (int z = 0; z < Dd; ++z)
        for (int y = 0; y < Dd; ++y)
            for (int x = 0; x < Dd; ++x) {
                const int sx = 2 * x, sy = 2 * y, sz = 2 * z;
                V4 v = {0, 0, 0, 0}; float k = 0.125f;
                if (mode == 0) {
                    static const int o[8][3] = {{0,0,0},{0,0,1},{0,1,0},{0,1,1},{1,0,0},{1,0,1},{1,1,0},{1,1,1}};
                    for (int i = 0; i < 8; ++i) v = add(v, load(sx + o[i][0], sy + o[i][1], sz + o[i][2]));
                } else if (mode == 1) {
                    for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) for (int q = -1; q <= 1; ++q) v = add(v, load(sx + i, sy + j, sz + q));
                    k = 0.037f;
                } else {
                    static const int o[7][3] = {{0,0,0},{0,0,1},{0,1,0},{1,0,0},{0,0,-1},{0,-1,0},{-1,0,0}};
                    for (int i = 0; i < 7; ++i) v = add(v, load(sx + o[i][0], sy + o[i][1], sz + o[i][2]));
                    k = 0.143f;
                }
                dst[((size_t)z * Dd + y) * Dd + x] = pack_unorm({v.x * k, v.y * k, v.z * k, v.w * k});
            }
}
// filter3d.comp:16-47 (dead variant): textureLodOffset on texel centres of the LINEAR_MIPMAP_LINEAR/NEAREST
// sampler with explicit lod 0 -> lambda <= 0.5 -> mag NEAREST -> exact texel fetches; border colour 0.
extern "C" void orc_filter3d(int Ds, const unsigned* src, unsigned* dst) {
    const int Dd = Ds / 2;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < Dd; ++z)
        for (int y = 0; y < Dd; ++y)
            for (int x = 0; x < Dd; ++x) {
                const float tc[3] = {(float)x / (float)Dd + 0.5f * (1.0f / (float)Ds), (float)y / (float)Dd + 0.5f * (1.0f / (float)Ds),
                                     (float)z / (float)Dd + 0.5f * (1.0f / (float)Ds)};
                int b[3]; for (int i = 0; i < 3; ++i) b[i] = (int)std::floor(tc[i] * (float)Ds);
                static const int o[8][3] = {{0,0,0},{0,0,1},{0,1,0},{0,1,1},{1,0,0},{1,0,1},{1,1,0},{1,1,1}};
                V4 v = {0, 0, 0, 0};
                for (int i = 0; i < 8; ++i) {
                    const int xx = b[0] + o[i][0], yy = b[1] + o[i][1], zz = b[2] + o[i][2];
                    if (xx < 0 || yy < 0 || zz < 0 || xx >= Ds || yy >= Ds || zz >= Ds) continue;
                    V4 t = unpack_unorm(src[((size_t)zz * Ds + yy) * Ds + xx]);
                    v = {v.x + t.x, v.y + t.y, v.z + t.z, v.w + t.w};
                }
                dst[((size_t)z * Dd + y) * Dd + x] = pack_unorm({v.x * 0.125f, v.y * 0.125f, v.z * 0.125f, v.w * 0.125f});
            }
}

// =================================================================================================== a9
namespace {
// One axis of calculateWarpPosition (src/main.cpp:101-123) == generateWarpmap.frag:77-90:
//   previousPartial = occupied ? partial-1 : partial ; offset = l*(index-previousPartial) + h*previousPartial ;
//   inner = fract*(occupied ? h : l) ; warped = (offset+inner)/dim
inline float warp_axis(float lo, float hi, bool occd, int part, int cid, float fr, int n) {
    const float res = occd ? hi : lo;
    const float prev = occd ? (float)part - 1.0f : (float)part;
    const float off = lo * ((float)cid - prev) + hi * prev;
    const float inner = fr * res;
    return (off + inner) / (float)n;
}
}  // namespace

// The reference's only worked example of the warp algorithm: the `#if 0` 2D rig of src/main.cpp:20-127 (same
// arithmetic as shaders/testWarpTexture.frag:21-56).  cells = n x n row-major [y][x]; fixed_low > 0 reproduces
// the rig's constant l (main.cpp:77-81), otherwise the production table (Application.cpp:346-370) is used.
extern "C" void orc_warp_rig(int n, const float* cells, float fixed_low, float high, float low, const float* tc, int npts,
                             float* out, int* part_x, int* part_y, float* wl, float* wh) {
    for (int row = 0; row < n; ++row) { int s = 0; for (int x = 0; x < n; ++x) { s += cells[row * n + x] > 0.5f ? 1 : 0; part_x[row * n + x] = s; } }
    for (int col = 0; col < n; ++col) { int s = 0; for (int y = 0; y < n; ++y) { s += cells[y * n + col] > 0.5f ? 1 : 0; part_y[y * n + col] = s; } }
    if (fixed_low > 0.0f) {
        for (int occ = 0; occ <= n; ++occ) {
            if (occ == 0 || occ == n) { wl[occ] = wh[occ] = 1.0f; continue; }
            const int empty = n - occ;
            wl[occ] = fixed_low; wh[occ] = ((float)n - fixed_low * (float)empty) / (float)occ;
        }
    } else orc_warp_weight_table(n, high, low, wl, wh);
    for (int i = 0; i < npts; ++i) {
        float idx[2], fr[2];
        for (int k = 0; k < 2; ++k) { const float lt = tc[2 * i + k] * (float)n; idx[k] = std::trunc(lt); fr[k] = lt - idx[k]; }
        const int x = (int)idx[0], y = (int)idx[1];
        const bool occd = cells[y * n + x] > 0.5f;
        const int tot[2] = {part_x[y * n + n - 1], part_y[(n - 1) * n + x]};
        const int part[2] = {part_x[y * n + x], part_y[y * n + x]};
        const int cid[2] = {x, y};
        for (int k = 0; k < 2; ++k) out[2 * i + k] = warp_axis(wl[tot[k]], wh[tot[k]], occd, part[k], cid[k], fr[k], n);
    }
}

extern "C" void orc_warp_weight_table(int dim, float high, float low, float* lo, float* hi) {     // Application.cpp:346-370
    for (int occ = 0; occ <= dim; ++occ) {
        if (occ == 0 || occ == dim) { lo[occ] = hi[occ] = 1.0f; continue; }
        const int empty = dim - occ;
        float h = high, l = ((float)dim - h * (float)occ) / (float)empty;
        if (l < low) { l = low; h = ((float)dim - l * (float)empty) / (float)occ; }
        lo[occ] = l; hi[occ] = h;
    }
}
// inclusive per-axis prefix counts of occupied cells, Application.cpp:311-343
static void warp_partials(const unsigned* occ, int* px, int* py, int* pz) {
    const int n = VCT_WARP_DIM;
    auto at = [&](int x, int y, int z) { return ((size_t)z * n + y) * n + x; };
    for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) { int s = 0; for (int x = 0; x < n; ++x) { s += occ[at(x, y, z)] > 0 ? 1 : 0; px[at(x, y, z)] = s; } }
    for (int z = 0; z < n; ++z) for (int x = 0; x < n; ++x) { int s = 0; for (int y = 0; y < n; ++y) { s += occ[at(x, y, z)] > 0 ? 1 : 0; py[at(x, y, z)] = s; } }
    for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) { int s = 0; for (int z = 0; z < n; ++z) { s += occ[at(x, y, z)] > 0 ? 1 : 0; pz[at(x, y, z)] = s; } }
}
extern "C" void orc_warp_partials(const unsigned* occ, int* xyz /* 32^3 x 3, interleaved like glm::ivec3 */) {
    const int n = VCT_WARP_DIM, N = n * n * n;
    std::vector<int> px(N), py(N), pz(N);
    warp_partials(occ, px.data(), py.data(), pz.data());
    for (int i = 0; i < N; ++i) { xyz[3 * i] = px[i]; xyz[3 * i + 1] = py[i]; xyz[3 * i + 2] = pz[i]; }
}
extern "C" void orc_warpmap(const unsigned* occ, const vct_frame_params* fp, unsigned short* warpmap, unsigned short* wlo16, unsigned short* whi16) {
    const int n = VCT_WARP_DIM;
    auto at = [&](int x, int y, int z) { return ((size_t)z * n + y) * n + x; };
    std::vector<int> px(n * n * n), py(n * n * n), pz(n * n * n);
    warp_partials(occ, px.data(), py.data(), pz.data());
    float wl[VCT_WARP_DIM + 1], wh[VCT_WARP_DIM + 1];
    orc_warp_weight_table(n, fp->warp_texture_high_resolution, fp->warp_texture_low_resolution, wl, wh);
    std::memset(warpmap, 0, sizeof(unsigned short) * 4 * n * n * n);    // unwritten texels: defined as 0 (quirk a9)
    std::memset(wlo16, 0, sizeof(unsigned short) * 4 * n * n * n);
    std::memset(whi16, 0, sizeof(unsigned short) * 4 * n * n * n);
    // The quad is drawn at 0.8 scale (quad.vert:12): pixel i is covered iff its centre lies in [3.2, 28.8],
    // and the interpolated tc there is ((i+.5)/32 - .5)/.8 + .5.
    auto covered = [&](int i) { const float c = (float)i + 0.5f; return c >= 3.2f && c <= 28.8f; };
    auto tc_of = [&](int i) { return (((float)i + 0.5f) / (float)n - 0.5f) / 0.8f + 0.5f; };
    struct Cell { int x, y, z; bool occd; int tot[3]; };
    auto cell_of = [&](const float tc[3], float frac[3]) {
        Cell c; int id[3];
        for (int k = 0; k < 3; ++k) { const float lt = tc[k] * (float)n; const float fl = std::trunc(lt); frac[k] = lt - fl; id[k] = (int)fl; }
        c.x = id[0]; c.y = id[1]; c.z = id[2];
        c.occd = occ[at(c.x, c.y, c.z)] > 0;
        c.tot[0] = px[at(n - 1, c.y, c.z)]; c.tot[1] = py[at(c.x, n - 1, c.z)]; c.tot[2] = pz[at(c.x, c.y, n - 1)];
        return c;
    };
    // pass 1: generateWarpmapWeights.frag:37-74 -> two RGBA16F targets
    if (fp->use_warpmap_weights_texture)
        for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) {
            if (!covered(x) || !covered(y)) continue;
            const float tc[3] = {tc_of(x), tc_of(y), ((float)z + 0.5f) / (float)n}; float fr[3];
            Cell c = cell_of(tc, fr);
            for (int k = 0; k < 3; ++k) { wlo16[at(x, y, z) * 4 + k] = f2h(wl[c.tot[k]]); whi16[at(x, y, z) * 4 + k] = f2h(wh[c.tot[k]]); }
            wlo16[at(x, y, z) * 4 + 3] = f2h(c.occd ? 0.0f : 1.0f); whi16[at(x, y, z) * 4 + 3] = f2h(c.occd ? 1.0f : 0.0f);
        }
    // pass 2: generateWarpmap.frag:44-100
    for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) {
        if (!covered(x) || !covered(y)) continue;
        const float tc[3] = {tc_of(x), tc_of(y), ((float)z + 0.5f) / (float)n}; float fr[3];
        Cell c = cell_of(tc, fr);
        const int part[3] = {px[at(c.x, c.y, c.z)], py[at(c.x, c.y, c.z)], pz[at(c.x, c.y, c.z)]};
        float lo[3], hi[3];
        if (fp->use_warpmap_weights_texture) {            // NEAREST, CLAMP_TO_EDGE fetch at tc
            int t[3]; for (int k = 0; k < 3; ++k) t[k] = std::min(std::max((int)std::floor(tc[k] * (float)n), 0), n - 1);
            for (int k = 0; k < 3; ++k) { lo[k] = h2f(wlo16[at(t[0], t[1], t[2]) * 4 + k]); hi[k] = h2f(whi16[at(t[0], t[1], t[2]) * 4 + k]); }
        } else for (int k = 0; k < 3; ++k) { lo[k] = wl[c.tot[k]]; hi[k] = wh[c.tot[k]]; }
        const int cid[3] = {c.x, c.y, c.z};
        float out[4];
        for (int k = 0; k < 3; ++k) {
            const float warped = warp_axis(lo[k], hi[k], c.occd, part[k], cid[k], fr[k], n);
            const bool use = fp->warp_texture_linear ? false : fp->warp_texture_axes[k] != 0;
            out[k] = use ? warped : tc[k];
        }
        const unsigned bits = ((unsigned)c.tot[0] & 31u) | ((unsigned)c.tot[1] & 31u) << 5 | ((unsigned)c.tot[2] & 31u) << 10 | (c.occd ? 1u : 0u) << 15;
        out[3] = (float)bits / 65535.0f;
        for (int k = 0; k < 4; ++k) {                     // RGBA16 unorm store
            float v = out[k]; if (!(v > 0.0f)) v = 0.0f; if (v > 1.0f) v = 1.0f;
            warpmap[at(x, y, z) * 4 + k] = (unsigned short)std::nearbyintf(v * 65535.0f);
        }
    }
}

// =================================================================================================== a0'
// Depth prepass (dither.frag, depth LESS, cull back) followed by GL_EQUAL shading (Application.cpp:936-977):
// the shaded surface at a pixel is the nearest alpha-tested fragment; among equal depths the LAST drawn wins.
// vis = depthbits << 32 | (0xFFFFFFFF - drawIndex); empty = all ones.
namespace {
struct ClipTri { RV v[3]; };
// near-plane (z >= -w) Sutherland-Hodgman; intersections always computed from the inside vertex
inline int clip_near(const RV in[3], ClipTri out[2]) {
    float d[3]; for (int i = 0; i < 3; ++i) d[i] = in[i].z + in[i].w;
    RV poly[4]; int n = 0;
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        const bool ii = d[i] >= 0.0f, jj = d[j] >= 0.0f;
        if (ii) poly[n++] = in[i];
        if (ii != jj) {
            const RV& a = ii ? in[i] : in[j]; const RV& b = ii ? in[j] : in[i];
            const float da = ii ? d[i] : d[j], db = ii ? d[j] : d[i];
            const float t = da / (da - db);
            poly[n++] = {(b.x - a.x) * t + a.x, (b.y - a.y) * t + a.y, (b.z - a.z) * t + a.z, (b.w - a.w) * t + a.w};
        }
    }
    if (n < 3) return 0;
    out[0] = {{poly[0], poly[1], poly[2]}};
    if (n == 4) { out[1] = {{poly[0], poly[2], poly[3]}}; return 2; }
    return 1;
}
inline float pixel_rho2(const Homog& h, const float uv[3][2], float nx, float ny, float dx, float dy, const Tex& t, float l_out[3]) {
    float l[3], lx[3], ly[3];
    homog_eval(h, nx, ny, l); homog_eval(h, nx + dx, ny, lx); homog_eval(h, nx, ny + dy, ly);
    const float u = interp(l, uv[0][0], uv[1][0], uv[2][0]), v = interp(l, uv[0][1], uv[1][1], uv[2][1]);
    const float ux = interp(lx, uv[0][0], uv[1][0], uv[2][0]) - u, vx = interp(lx, uv[0][1], uv[1][1], uv[2][1]) - v;
    const float uy = interp(ly, uv[0][0], uv[1][0], uv[2][0]) - u, vy = interp(ly, uv[0][1], uv[1][1], uv[2][1]) - v;
    const float ax = ux * (float)t.w, bx = vx * (float)t.h, ay = uy * (float)t.w, by = vy * (float)t.h;
    for (int i = 0; i < 3; ++i) l_out[i] = l[i];
    return maxf(ax * ax + bx * bx, ay * ay + by * by);
}
}  // namespace

extern "C" void orc_visibility(const orc_scene* sc, const vct_frame_params* fp, int W, int H, unsigned long long* vis) {
    Prepared P = prepare(sc, false);
    const ClipMats cam_clip(sc, fp->projection, fp->view);
    for (size_t i = 0; i < (size_t)W * H; ++i) vis[i] = ~0ull;
    int nb = 1;
#ifdef _OPENMP
    nb = omp_get_max_threads();
#endif
    const int band = (H + nb - 1) / nb;
#pragma omp parallel for schedule(static, 1)
    for (int bnd = 0; bnd < nb; ++bnd) {
        const int ylo = bnd * band, yhi = std::min(H, ylo + band) - 1;
        for (int t = 0; t < sc->n_tris; ++t) {
            RV cv[3]; float uv[3][2];
            for (int k = 0; k < 3; ++k) {
                const unsigned vi = sc->indices[3 * t + k];
                V3 w = P.wpos[vi];
                V4 c = cam_clip.apply(sc, fp->projection, fp->view, vi, w);
                cv[k] = {c.x, c.y, c.z, c.w};
                uv[k][0] = sc->vertices[14 * (size_t)vi + 6]; uv[k][1] = sc->vertices[14 * (size_t)vi + 7];
            }
            ClipTri ct[2]; const int nct = clip_near(cv, ct);
            if (!nct) continue;
            const int m = sc->tri_material[t];
            const bool alpha = has_alpha(sc, m);
            const Tex* at = alpha ? &P.tex[sc->materials[m].alpha_tex] : nullptr;
            Homog hg; if (alpha) hg = homog_setup(cv);
            for (int q = 0; q < nct; ++q) {
                Setup s = tri_setup(ct[q].v, W, H, true);
                if (!s.valid || s.y1 < ylo || s.y0 > yhi) continue;
                const float z0 = ct[q].v[0].z / ct[q].v[0].w, z1 = ct[q].v[1].z / ct[q].v[1].w, z2 = ct[q].v[2].z / ct[q].v[2].w;
                raster(s, ylo, yhi, [&](int px, int py, const float l[3]) {
                    const float z = interp(l, z0, z1, z2);
                    if (z < -1.0f || z > 1.0f) return;
                    if (alpha) {
                        const float nx = ((float)px + 0.5f) / (float)W * 2.0f - 1.0f, ny = ((float)py + 0.5f) / (float)H * 2.0f - 1.0f;
                        float lb[3];
                        const float rho2 = pixel_rho2(hg, uv, nx, ny, 2.0f / (float)W, 2.0f / (float)H, *at, lb);
                        const float u = interp(lb, uv[0][0], uv[1][0], uv[2][0]), v = interp(lb, uv[0][1], uv[1][1], uv[2][1]);
                        const bool kept = !(sample2d(*at, u, v, rho2).x < 0.1f);
                        alpha_record(1, m, u, v, rho2, kept);
                        if (!kept) return;
                    }
                    const float d = z * 0.5f + 0.5f;
                    uint32_t db; std::memcpy(&db, &d, 4);
                    const unsigned long long key = (unsigned long long)db << 32 | (0xFFFFFFFFu - (unsigned)t);
                    unsigned long long& dst = vis[(size_t)py * W + px];
                    if (key < dst) dst = key;
                });
            }
        }
    }
}

// =================================================================================================== a7
namespace {
const float PI_REF = 3.1415982f;                        // common.glsl:1 [sic]

struct Vol { int D, L; const unsigned* lv[VCT_MAX_LEVELS]; };
inline V4 vol_texel(const Vol& v, int l, int x, int y, int z) {           // CLAMP_TO_BORDER, border 0
    const int d = std::max(1, v.D >> l);
    if (x < 0 || y < 0 || z < 0 || x >= d || y >= d || z >= d) return {0, 0, 0, 0};
    return unpack_unorm(v.lv[l][((size_t)z * d + y) * d + x]);
}
inline V4 vol_linear(const Vol& v, int l, V3 s) {
    const float d = (float)std::max(1, v.D >> l);
    const float x = s.x * d - 0.5f, y = s.y * d - 0.5f, z = s.z * d - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y), fz0 = std::floor(z);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    V4 c00 = lerp4(vol_texel(v, l, x0, y0, z0), vol_texel(v, l, x0 + 1, y0, z0), fx);
    V4 c10 = lerp4(vol_texel(v, l, x0, y0 + 1, z0), vol_texel(v, l, x0 + 1, y0 + 1, z0), fx);
    V4 c01 = lerp4(vol_texel(v, l, x0, y0, z0 + 1), vol_texel(v, l, x0 + 1, y0, z0 + 1), fx);
    V4 c11 = lerp4(vol_texel(v, l, x0, y0 + 1, z0 + 1), vol_texel(v, l, x0 + 1, y0 + 1, z0 + 1), fx);
    return lerp4(lerp4(c00, c10, fy), lerp4(c01, c11, fy), fz);
}
// textureLod on sampler3D with min LINEAR_MIPMAP_LINEAR, mag NEAREST, CLAMP_TO_BORDER(0)
// (Application.cpp:1094-1098, Application.h:146).  GL 4.5 §8.14: magnification iff lambda <= c, c = 0.5 for
// this filter pair -> NEAREST on the base level; otherwise linear blend of LINEAR samples of levels
// floor(lambda), floor(lambda)+1, lambda clamped to the last level.
inline V4 vol_sample(const Vol& v, V3 s, float lambda) {
    if (!(lambda > 0.5f)) {
        const float d = (float)v.D;
        return vol_texel(v, 0, (int)std::floor(s.x * d), (int)std::floor(s.y * d), (int)std::floor(s.z * d));
    }
    const float q = (float)(v.L - 1);
    if (lambda >= q) return vol_linear(v, v.L - 1, s);
    const float fl = std::floor(lambda);
    const int l0 = (int)fl;
    return lerp4(vol_linear(v, l0, s), vol_linear(v, l0 + 1, s), lambda - fl);
}

// phong.frag:135-180
inline V4 trace_cone(const Vol& vol, const vct_frame_params* fp, const uint16_t* warpmap, V3 position, V3 normal, V3 direction,
                     int steps, float bias, float cone_angle, float cone_height, float lod_offset, unsigned long long& fetches) {
    direction = normalize(direction);
    V3 color = {0, 0, 0}; float alpha = 0.0f;
    const float scale = 1.0f / (float)vol.D;
    V3 start = position + (normal * bias) * scale;
    for (int i = 0; i < steps && alpha < 0.95f; ++i) {
        const float cone_radius = cone_height * std::tan(cone_angle / 2.0f);
        const float lod = std::log2(maxf(1.0f, 2.0f * cone_radius));
        V3 sp = start + (direction * cone_height) * scale;
        if (!(sp.x == clampf(sp.x, 0.0f, 1.0f)) || !(sp.y == clampf(sp.y, 0.0f, 1.0f)) || !(sp.z == clampf(sp.z, 0.0f, 1.0f))) break;
        if (fp->warp_texture && warpmap) sp = warp_sample(warpmap, sp);
        else if (fp->warp_voxels) sp = voxel_warp(sp, voxel_linear_position(eye3(fp), fp));
        else if (fp->voxelize_tesselation_warp) {                          // :158-162: back to world space, then through pv
            const V3 world = {(sp.x * (fp->voxel_max[0] - fp->voxel_min[0]) + fp->voxel_center[0]) + fp->voxel_min[0],
                              (sp.y * (fp->voxel_max[1] - fp->voxel_min[1]) + fp->voxel_center[1]) + fp->voxel_min[1],
                              (sp.z * (fp->voxel_max[2] - fp->voxel_min[2]) + fp->voxel_center[2]) + fp->voxel_min[2]};
            sp = tess_warp_position(world, fp);                            // getVoxelPosition(world, ..., false): warpTexture is off on this branch
        }
        V4 sc = vol_sample(vol, sp, lod + lod_offset);
        fetches++;
        const float a = 1.0f - alpha;
        color = color + v3(sc.x, sc.y, sc.z) * a;
        alpha += a * sc.w;
        cone_height += cone_radius;
    }
    return {color.x, color.y, color.z, alpha};
}

inline float pow2(float x) { return x * x; }
// phong.frag:524-549
inline float D_ggxtr(V3 N, V3 Hh, float rough) {
    const float ndh = maxf(0.0f, dot(N, Hh)), a2 = pow2(rough);
    return a2 / (PI_REF * pow2(pow2(ndh) * (a2 - 1.0f) + 1.0f));
}
inline float G1(V3 N, V3 V, float rough) {
    const float k = pow2(rough + 1.0f) / 8.0f, ndv = maxf(0.0f, dot(N, V));
    return ndv / (ndv * (1.0f - k) + k);
}
struct LR { V3 diffuse, specular; };
// phong.frag:230-258
inline LR cook_torrance(V3 dc, V3 lc, V3 N, V3 V, V3 L, V3 Hh, float rough, float metal) {
    const float ndl = maxf(0.0f, dot(N, L)), ndv = maxf(0.0f, dot(N, V)), hdv = maxf(0.0f, dot(Hh, V));
    V3 lambert = dc / PI_REF;
    const float Dg = D_ggxtr(N, Hh, rough), G = G1(N, V, rough) * G1(N, L, rough);
    V3 F0 = {mixf(0.04f, dc.x, metal), mixf(0.04f, dc.y, metal), mixf(0.04f, dc.z, metal)};
    const float p5 = std::pow(1.0f - hdv, 5.0f);
    V3 F = {F0.x + (1.0f - F0.x) * p5, F0.y + (1.0f - F0.y) * p5, F0.z + (1.0f - F0.z) * p5};
    const float den = maxf(4.0f * ndl * ndv, 0.001f);
    V3 fct = (F * (Dg * G)) / den;
    V3 kd = {(1.0f - F.x) * (1.0f - metal), (1.0f - F.y) * (1.0f - metal), (1.0f - F.z) * (1.0f - metal)};
    LR r; r.diffuse = ((lc * ndl) * kd) * lambert; r.specular = (lc * ndl) * fct;
    return r;
}
inline V3 postprocess(V3 c) {                            // phong.frag:210-218
    c = {c.x / (c.x + 1.0f), c.y / (c.y + 1.0f), c.z / (c.z + 1.0f)};
    const float g = 1.0f / 2.2f;
    return {std::pow(c.x, g), std::pow(c.y, g), std::pow(c.z, g)};
}
}  // namespace

extern "C" void orc_shade_rows(const orc_scene* sc, const vct_frame_params* fp, int W, int H, int y_lo, int y_hi, int y_stride,
                               const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                               const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps);
// voxelNormal for VCT_VIEW_VOXEL_NORMALS (phong.frag:350-353; texture unit 1, Application.cpp:1046): set before a shade call, D^3 words
static const unsigned* g_normal_volume = nullptr;
extern "C" void orc_set_normal_volume(const unsigned* normal) { g_normal_volume = normal; }
extern "C" void orc_shade(const orc_scene* sc, const vct_frame_params* fp, int W, int H, const unsigned long long* vis, int D, int L,
                          const unsigned* radiance_pyr, const unsigned* color_pyr, const float* shadow, int S,
                          const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps) {
    orc_shade_rows(sc, fp, W, H, 0, H, 1, vis, D, L, radiance_pyr, color_pyr, shadow, S, warpmap, image, cone_steps);
}
// rows y_lo, y_lo + y_stride, ... < y_hi only (bench.py's bounded CPU sample; other rows of `image` are untouched)
static void shade_impl(const orc_scene* sc, const vct_frame_params* fp, int W, int H, int y_lo, int y_hi, int y_stride,
                       const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                       const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps, float* frag_rec) {
    Prepared P = prepare(sc, true);
    const ClipMats cam_clip(sc, fp->projection, fp->view), ls_clip(sc, fp->ls, nullptr);
    Vol rad, colv; rad.D = colv.D = D; rad.L = colv.L = L;
    { size_t off = 0; for (int l = 0; l < L; ++l) { rad.lv[l] = radiance_pyr + off; colv.lv[l] = color_pyr ? color_pyr + off : nullptr; const size_t d = std::max(1, D >> l); off += d * d * d; } }
    const Vol& vol = (fp->draw_radiance || !color_pyr) ? rad : colv;      // radiance ? voxelRadiance : voxelColor
    const unsigned clear = pack_unorm({fp->clear_color[0], fp->clear_color[1], fp->clear_color[2], 1.0f});
    unsigned long long fetch_total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : fetch_total)
    for (int py = y_lo; py < y_hi; py += y_stride)
        for (int px = 0; px < W; ++px) {
            const unsigned long long key = vis[(size_t)py * W + px];
            unsigned scratch = 0;
            unsigned& out = image ? image[(size_t)py * W + px] : scratch;     // image == NULL: record only
            if (key == ~0ull) { out = clear; continue; }
            const int t = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
            const unsigned* ix = sc->indices + 3 * (size_t)t;
            RV cv[3]; float uv[3][2]; V4 lfp[3];
            for (int k = 0; k < 3; ++k) {
                V3 w = P.wpos[ix[k]];
                V4 c = cam_clip.apply(sc, fp->projection, fp->view, ix[k], w); cv[k] = {c.x, c.y, c.z, c.w};
                lfp[k] = ls_clip.apply(sc, fp->ls, nullptr, ix[k], w);         // phong.vert:45
                uv[k][0] = sc->vertices[14 * (size_t)ix[k] + 6]; uv[k][1] = sc->vertices[14 * (size_t)ix[k] + 7];
            }
            const vct_material& mat = sc->materials[sc->tri_material[t]];
            Homog hg = homog_setup(cv);
            const float nx = ((float)px + 0.5f) / (float)W * 2.0f - 1.0f, ny = ((float)py + 0.5f) / (float)H * 2.0f - 1.0f;
            float l[3], lx[3], ly[3];
            homog_eval(hg, nx, ny, l); homog_eval(hg, nx + 2.0f / (float)W, ny, lx); homog_eval(hg, nx, ny + 2.0f / (float)H, ly);
            const float u = interp(l, uv[0][0], uv[1][0], uv[2][0]), v = interp(l, uv[0][1], uv[1][1], uv[2][1]);
            const float ux = interp(lx, uv[0][0], uv[1][0], uv[2][0]) - u, vx = interp(lx, uv[0][1], uv[1][1], uv[2][1]) - v;
            const float uy = interp(ly, uv[0][0], uv[1][0], uv[2][0]) - u, vy = interp(ly, uv[0][1], uv[1][1], uv[2][1]) - v;
            auto fetch = [&](int tex) -> V4 {
                const Tex& T = P.tex[tex];
                const float ax = ux * (float)T.w, bx = vx * (float)T.h, ay = uy * (float)T.w, by = vy * (float)T.h;
                return sample2d(T, u, v, maxf(ax * ax + bx * bx, ay * ay + by * by));
            };
            V3 Pw = interp3(l, P.wpos[ix[0]], P.wpos[ix[1]], P.wpos[ix[2]]);
            V3 fn = interp3(l, P.wnrm[ix[0]], P.wnrm[ix[1]], P.wnrm[ix[2]]);
            V3 Tt = interp3(l, P.T[ix[0]], P.T[ix[1]], P.T[ix[2]]), Bt = interp3(l, P.B[ix[0]], P.B[ix[1]], P.B[ix[2]]);
            V4 lsp = {interp(l, lfp[0].x, lfp[1].x, lfp[2].x), interp(l, lfp[0].y, lfp[1].y, lfp[2].y), interp(l, lfp[0].z, lfp[1].z, lfp[2].z), interp(l, lfp[0].w, lfp[1].w, lfp[2].w)};
            if (frag_rec) {                                                   // VS_OUT block + material + uv footprint of this pixel
                float* r = frag_rec + 28 * ((size_t)py * W + px);
                const float rec[28] = {Pw.x, Pw.y, Pw.z, fn.x, fn.y, fn.z, u, v, lsp.x, lsp.y, lsp.z, lsp.w, Tt.x, Tt.y, Tt.z, Bt.x, Bt.y, Bt.z,
                                     ux, vx, uy, vy, (float)sc->tri_material[t], 1.0f, 0, 0, 0, 0};
                std::memcpy(r, rec, sizeof rec);
                if (!image) continue;                                         // rasteriser + interpolation without the fragment stage
            }
            auto tbn = [&](V3 d) -> V3 { return (Tt * d.x + Bt * d.y) + fn * d.z; };   // mat3(T,B,N) * d
            const int view = fp->debug_view;
            if (view == VCT_VIEW_WARP_TEXTURE || view == VCT_VIEW_WARP_TEXTURE_TC) {   // :354-357 (voxelize && debugWarpTexture; `toggle` shows the linear position)
                const V3 tc = voxel_linear_position(Pw, fp);
                const V3 wv = view == VCT_VIEW_WARP_TEXTURE_TC ? tc : (warpmap ? warp_sample(warpmap, tc) : V3{0.0f, 0.0f, 0.0f});   // no warp map generated yet: an empty texture
                out = pack_unorm({wv.x, wv.y, wv.z, 1.0f});
                continue;
            }
            if (view == VCT_VIEW_VOXELS || view == VCT_VIEW_VOXEL_NORMALS) {   // phong.frag:347-404 (`voxelize`; the warp-slope sub-views index out of bounds: not built)
                V3 gp = get_voxel_position(Pw, fp, warpmap);
                const float Df = (float)D;
                V3 vi = {(Df * gp.x) / Df, (Df * gp.y) / Df, (Df * gp.z) / Df};   // voxelIndex(...) / voxelDim
                if (view == VCT_VIEW_VOXEL_NORMALS) {                         // :350-353: one level, NEAREST min and mag, CLAMP_TO_BORDER 0
                    const float c3[3] = {std::floor(vi.x * Df), std::floor(vi.y * Df), std::floor(vi.z * Df)};
                    V4 t4 = {0, 0, 0, 0};
                    if (g_normal_volume && c3[0] >= 0 && c3[1] >= 0 && c3[2] >= 0 && c3[0] < Df && c3[1] < Df && c3[2] < Df)
                        t4 = unpack_unorm(g_normal_volume[((size_t)(int)c3[2] * D + (int)c3[1]) * D + (int)c3[0]]);
                    out = pack_unorm({t4.x, t4.y, t4.z, 1.0f});
                    continue;
                }
                V4 c = vol_sample(vol, vi, fp->miplevel);
                fetch_total += 1;                                             // a volume fetch like any cone step
                out = pack_unorm({c.x, c.y, c.z, 1.0f});
                continue;
            }
            if (view == VCT_VIEW_MATERIAL_DIFFUSE || view == VCT_VIEW_MATERIAL_ROUGHNESS || view == VCT_VIEW_MATERIAL_METALLIC) {   // :405-425
                V3 c = {0.5f, 0.0f, 0.5f};
                if (view == VCT_VIEW_MATERIAL_DIFFUSE && mat.diffuse_tex >= 0) { V4 t4 = fetch(mat.diffuse_tex); c = {t4.x, t4.y, t4.z}; }
                if (view == VCT_VIEW_MATERIAL_ROUGHNESS && mat.roughness_tex >= 0) { const float r = fetch(mat.roughness_tex).x; c = {r, r, r}; }
                if (view == VCT_VIEW_MATERIAL_METALLIC && mat.metallic_tex >= 0) { const float r = fetch(mat.metallic_tex).x; c = {r, r, r}; }
                out = pack_unorm({c.x, c.y, c.z, 1.0f});
                continue;
            }
            // phong.frag:427-439
            V3 N;
            if (fp->enable_normal_map && mat.normal_tex >= 0) {
                V4 nm = fetch(mat.normal_tex);
                V3 n = normalize(v3(nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f));
                N = normalize(tbn(n));
            } else N = normalize(fn);
            if (view == VCT_VIEW_NORMALS) { out = pack_unorm({N.x, N.y, N.z, 1.0f}); continue; }   // :441-443
            if (view == VCT_VIEW_DOMINANT_AXIS) {                            // :444-447  step(vec3(max component), |n|)
                const float ax = std::fabs(N.x), ay = std::fabs(N.y), az = std::fabs(N.z), m = maxf(maxf(ax, ay), az);
                out = pack_unorm({ax < m ? 0.0f : 1.0f, ay < m ? 0.0f : 1.0f, az < m ? 0.0f : 1.0f, 1.0f});
                continue;
            }
            V4 dc4 = mat.diffuse_tex >= 0 ? fetch(mat.diffuse_tex) : V4{mat.diffuse[0], mat.diffuse[1], mat.diffuse[2], 1.0f};
            V3 dc = {dc4.x, dc4.y, dc4.z};
            // calculateDirectLighting, phong.frag:305-344
            V3 eye = eye3(fp);
            V3 Vv = normalize(eye - Pw);
            float rough = 0.5f; if (mat.roughness_tex >= 0) rough = fetch(mat.roughness_tex).x;
            float metal = 0.0f; if (mat.metallic_tex >= 0) metal = fetch(mat.metallic_tex).x;
            V3 dsum = {0, 0, 0}, ssum = {0, 0, 0};
            for (int i = 0; i < sc->n_lights; ++i) {
                const vct_light& Lt = sc->lights[i];
                if (!Lt.enabled) continue;
                V3 lc = {Lt.color[0], Lt.color[1], Lt.color[2]}, lpos = {Lt.position[0], Lt.position[1], Lt.position[2]};
                LR r = {{0, 0, 0}, {0, 0, 0}};
                if (Lt.type == 0u) {                                          // :267-287
                    const float dist = length(lpos - Pw);
                    if (!(dist > Lt.range)) {
                        const float e0 = 0.75f * Lt.range, tt = clampf((dist - e0) / (Lt.range - e0), 0.0f, 1.0f);
                        const float att = 1.0f - tt * tt * (3.0f - 2.0f * tt);
                        V3 Ld = normalize(lpos - Pw);
                        if (fp->cooktorrance) { r = cook_torrance(dc, lc, N, Vv, Ld, normalize(Vv + Ld), rough, metal); r.diffuse = r.diffuse * att; r.specular = r.specular * att; }
                        else {
                            const float df = maxf(dot(N, Ld), 0.0f), sp = std::pow(maxf(dot(N, normalize(Vv + Ld)), 0.0f), mat.shininess);
                            r.diffuse = ((lc * (df * att)) * Lt.intensity) * dc; r.specular = ((lc * (sp * att)) * Lt.intensity) * dc;
                        }
                    }
                } else if (Lt.type == 1u) {                                   // :289-301
                    V3 Ld = normalize(v3(-Lt.direction[0], -Lt.direction[1], -Lt.direction[2]));
                    if (fp->cooktorrance) r = cook_torrance(dc, lc, N, Vv, Ld, normalize(Vv + Ld), rough, metal);
                    else {
                        const float df = maxf(dot(N, Ld), 0.0f), sp = std::pow(maxf(dot(N, normalize(Vv + Ld)), 0.0f), mat.shininess);
                        r.diffuse = ((lc * df) * Lt.intensity) * dc; r.specular = ((lc * sp) * Lt.intensity) * dc;
                    }
                }
                if (Lt.shadow_caster) { const float sf = 1.0f - calc_shadow_factor(shadow, S, lsp); r.diffuse = r.diffuse * sf; r.specular = r.specular * sf; }
                dsum = dsum + r.diffuse; ssum = ssum + r.specular;
            }
            if (!fp->enable_diffuse) dsum = {0, 0, 0};
            if (!fp->enable_specular) ssum = {0, 0, 0};
            V3 col;
            if (fp->enable_indirect) {                                        // phong.frag:455-512
                V3 vp = voxel_linear_position(Pw, fp);
                static const float dirs[6][3] = {{0, 1, 0}, {0, 0.5f, 0.866025f}, {0.823639f, 0.5f, 0.267617f}, {0.509037f, 0.5f, -0.700629f},
                                                 {-0.5909037f, 0.5f, -0.700629f}, {-0.823639f, 0.5f, 0.267617f}};
                static const float wts[6] = {0.25f, 0.15f, 0.15f, 0.15f, 0.15f, 0.15f};
                V4 ind = {0, 0, 0, 0}; unsigned long long f = 0;
                for (int i = 0; i < 6; ++i) {
                    V3 dir = normalize(tbn(v3(dirs[i][0], dirs[i][1], dirs[i][2])));
                    V4 c = trace_cone(vol, fp, warpmap, vp, N, dir, fp->diffuse_cone.steps, fp->diffuse_cone.bias, fp->diffuse_cone.cone_angle,
                                      fp->diffuse_cone.cone_initial_height, fp->diffuse_cone.lod_offset, f);
                    ind = {ind.x + wts[i] * c.x, ind.y + wts[i] * c.y, ind.z + wts[i] * c.z, ind.w + wts[i] * c.w};
                }
                const float occl = 1.0f - clampf(ind.w, 0.0f, 1.0f);
                if (view == VCT_VIEW_INDIRECT) {                              // :489  (returns before reflections and post-processing)
                    const float k = fp->draw_occlusion ? occl : 1.0f;
                    fetch_total += f; out = pack_unorm({ind.x * k, ind.y * k, ind.z * k, 1.0f}); continue;
                }
                if (view == VCT_VIEW_OCCLUSION) { fetch_total += f; out = pack_unorm({occl, occl, occl, 1.0f}); continue; }   // :490
                if (fp->enable_reflections) {
                    float ang = fp->specular_cone.cone_angle;
                    if (fp->specular_cone_angle_from_roughness && mat.roughness_tex >= 0) ang = (fetch(mat.roughness_tex).x * PI_REF) * 0.1f;
                    V3 I = Pw - eye;
                    V3 R = I - N * (2.0f * dot(N, I));                        // reflect(I, N)
                    V4 rc = trace_cone(vol, fp, warpmap, vp, N, R, fp->specular_cone.steps, fp->specular_cone.bias, ang,
                                       fp->specular_cone.cone_initial_height, fp->specular_cone.lod_offset, f);
                    ind.x += rc.x * fp->reflect_scale; ind.y += rc.y * fp->reflect_scale; ind.z += rc.z * fp->reflect_scale;
                    if (view == VCT_VIEW_REFLECTIONS) { fetch_total += f; out = pack_unorm({rc.x, rc.y, rc.z, 1.0f}); continue; }   // :505
                }
                fetch_total += f;
                V3 indc = v3(ind.x, ind.y, ind.z) * (dc * fp->ambient_scale);  // indirect.rgb *= ambientScale * diffuseColor.rgb
                col = (indc + dsum) + ssum;
                if (fp->draw_occlusion) col = col * occl;
            } else col = (dc * fp->ambient_scale + dsum) + ssum;
            if (fp->enable_postprocess) col = postprocess(col);
            out = pack_unorm({col.x, col.y, col.z, 1.0f});
        }
    if (cone_steps) *cone_steps = fetch_total;
}
extern "C" void orc_shade_rows(const orc_scene* sc, const vct_frame_params* fp, int W, int H, int y_lo, int y_hi, int y_stride,
                               const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                               const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps) {
    shade_impl(sc, fp, W, H, y_lo, y_hi, y_stride, vis, D, L, radiance_pyr, color_pyr, shadow, S, warpmap, image, cone_steps, nullptr);
}
// Same pass, additionally recording the fragment-stage inputs of every covered pixel: 28 floats per pixel =
// VS_OUT{fragPosition 3, fragNormal 3, fragTexcoord 2, lightFragPos 4, TBN columns T 3 and B 3 (the third is fragNormal)},
// the uv differences to the right / upper neighbour pixel (du_x, dv_x, du_y, dv_y), material id, covered flag
// (tests/test_glsl_ref.py replays them through the reference's phong.frag compiled as C++).
extern "C" void orc_shade_trace(const orc_scene* sc, const vct_frame_params* fp, int W, int H, const unsigned long long* vis, int D, int L,
                                const unsigned* radiance_pyr, const unsigned* color_pyr, const float* shadow, int S,
                                const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps, float* frag_rec) {
    shade_impl(sc, fp, W, H, 0, H, 1, vis, D, L, radiance_pyr, color_pyr, shadow, S, warpmap, image, cone_steps, frag_rec);
}
// rows y_lo, y_lo + y_stride, ... only; image may be NULL (record only)
extern "C" void orc_shade_trace_rows(const orc_scene* sc, const vct_frame_params* fp, int W, int H, int y_lo, int y_hi, int y_stride,
                                     const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                                     const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps, float* frag_rec) {
    shade_impl(sc, fp, W, H, y_lo, y_hi, y_stride, vis, D, L, radiance_pyr, color_pyr, shadow, S, warpmap, image, cone_steps, frag_rec);
}

// closed-form KAT helper: one cone marched through a volume whose every level holds the same word
extern "C" float orc_cone_trace_const(int D, int L, unsigned word, const vct_cone_settings* cs, int* steps_out) {
    std::vector<std::vector<unsigned>> lv(L);
    Vol v; v.D = D; v.L = L;
    for (int l = 0; l < L; ++l) { const size_t d = std::max(1, D >> l); lv[l].assign(d * d * d, word); v.lv[l] = lv[l].data(); }
    vct_frame_params fp{}; unsigned long long f = 0;
    V4 r = trace_cone(v, &fp, nullptr, {0.5f, 0.5f, 0.5f}, {0, 0, 1}, {0, 0, 1}, cs->steps, cs->bias, cs->cone_angle, cs->cone_initial_height, cs->lod_offset, f);
    if (steps_out) *steps_out = (int)f;
    return r.w;
}

extern "C" int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ============================================================================= N3: Application::debugVoxels (src/Application.cpp:1222-1275)
// glDrawArraysInstanced(GL_POINTS, 0, 1, D^3) through debugVoxels.vert / .geom / .frag: every instance is a voxel of the base grid (its colour the
// pyramid sampled at the voxel's centre with lod = Settings::miplevel); a voxel with alpha > 0 becomes a cube of one voxel's size — a 21-vertex
// triangle strip, i.e. 19 triangles, the repeated indices making degenerate stitches — drawn with depth test (GL_LESS), back-face culling (CCW front) and the
// voxel's colour.  Canonical fixed function as everywhere else in this file: the same fixed-point rasteriser, near-plane clipping in clip space,
// far plane per fragment, z_ndc interpolated with screen-space barycentrics, window depth z_ndc * 0.5 + 0.5 compared as fp32; among equal depths
// the FIRST drawn fragment stays (GL_LESS), instances in order, triangles of a strip in order (odd strip triangles have their winding reversed,
// OpenGL 4.5 section 10.1.8).  quirk kept: debugVoxels.vert derives the voxel from float(gl_InstanceID) with two modf calls — beyond 2^24 instances
// (512^3) consecutive ids collapse onto even multiples, so some voxels are drawn twice and others never.
namespace {
struct DbgVoxel { V3 tc; V3 world; };
inline DbgVoxel debug_voxel_instance(unsigned id, int D, const vct_frame_params* fp) {       // debugVoxels.vert:19-30
    const float dim = (float)D;
    float instance = (float)id;
    float t = instance / dim; instance = std::trunc(t); const float x = t - instance;       // modf(instance / voxelDim, instance)
    t = instance / dim; instance = std::trunc(t); const float y = t - instance;
    const float z = instance / dim;
    const float h = 0.5f / dim;
    DbgVoxel o; o.tc = {x + h, y + h, z + h};
    const float tcv[3] = {o.tc.x, o.tc.y, o.tc.z}; float w[3];
    for (int k = 0; k < 3; ++k)                                                               // position (0,0,0) + voxelCenter + mix(voxelMin, voxelMax, voxelPosition)
        w[k] = (0.0f + fp->voxel_center[k]) + (fp->voxel_min[k] * (1.0f - tcv[k]) + fp->voxel_max[k] * tcv[k]);
    o.world = {w[0], w[1], w[2]};
    return o;
}
const float kCubeVertices[8][3] = {{-0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, -0.5f}, {0.5f, 0.5f, 0.5f}, {-0.5f, 0.5f, 0.5f},
                                   {-0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, -0.5f}, {0.5f, -0.5f, 0.5f}, {-0.5f, -0.5f, 0.5f}};   // debugVoxels.geom:16-26
const int kCubeStrip[21] = {5, 4, 1, 0, 0, 0, 0, 3, 1, 2, 5, 6, 4, 7, 0, 3, 3, 3, 2, 7, 6};                                  // :27-33
}  // namespace

extern "C" void orc_debug_voxel_vertices(const vct_frame_params* fp, int D, unsigned id, float* world3, float* clip21x4) {
    float mvp[16];                                                                            // glm::mat4 mvp = projection * view (Application.cpp:926)
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            mvp[4 * c + r] = ((fp->projection[r] * fp->view[4 * c] + fp->projection[4 + r] * fp->view[4 * c + 1]) + fp->projection[8 + r] * fp->view[4 * c + 2]) + fp->projection[12 + r] * fp->view[4 * c + 3];
    const DbgVoxel v = debug_voxel_instance(id, D, fp);
    world3[0] = v.world.x; world3[1] = v.world.y; world3[2] = v.world.z;
    const float dim = (float)D;
    const float size[3] = {(fp->voxel_max[0] - fp->voxel_min[0]) / dim, (fp->voxel_max[1] - fp->voxel_min[1]) / dim, (fp->voxel_max[2] - fp->voxel_min[2]) / dim};
    for (int i = 0; i < 21; ++i) {
        const float* cvx = kCubeVertices[kCubeStrip[i]];
        const V4 p = mul(mvp, {size[0] * cvx[0] + v.world.x, size[1] * cvx[1] + v.world.y, size[2] * cvx[2] + v.world.z, 1.0f});
        clip21x4[4 * i] = p.x; clip21x4[4 * i + 1] = p.y; clip21x4[4 * i + 2] = p.z; clip21x4[4 * i + 3] = p.w;
    }
}

extern "C" void orc_debug_voxel_color(const vct_frame_params* fp, int D, int L, const unsigned* pyr, unsigned id, float* rgba) {   // debugVoxels.vert:25
    Vol vol; vol.D = D; vol.L = L;
    { size_t off = 0; for (int l = 0; l < L; ++l) { vol.lv[l] = pyr + off; const size_t d = std::max(1, D >> l); off += d * d * d; } }
    const V4 c = vol_sample(vol, debug_voxel_instance(id, D, fp).tc, fp->miplevel);
    rgba[0] = c.x; rgba[1] = c.y; rgba[2] = c.z; rgba[3] = c.w;
}
// pyr: the levels of the pyramid the reference binds (settings.drawRadiance ? voxelRadiance : voxelColor, Application.cpp:927), packed back to back
extern "C" void orc_debug_voxels(const vct_frame_params* fp, int W, int H, int D, int L, const unsigned* pyr, unsigned* image) {
    Vol vol; vol.D = D; vol.L = L;
    { size_t off = 0; for (int l = 0; l < L; ++l) { vol.lv[l] = pyr + off; const size_t d = std::max(1, D >> l); off += d * d * d; } }
    const unsigned clear = pack_unorm({fp->clear_color[0], fp->clear_color[1], fp->clear_color[2], 1.0f});
    std::vector<float> depth((size_t)W * H, 1.0f);                                            // glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
    for (size_t i = 0; i < (size_t)W * H; ++i) image[i] = clear;
    const unsigned n = (unsigned)D * D * D;
    for (unsigned id = 0; id < n; ++id) {
        const DbgVoxel v = debug_voxel_instance(id, D, fp);
        const V4 c = vol_sample(vol, v.tc, fp->miplevel);                                     // debugVoxels.vert:25 textureLod(voxels, tc, level)
        if (!(c.w > 0.0f)) continue;                                                          // debugVoxels.geom:46
        const unsigned word = pack_unorm({c.x, c.y, c.z, 1.0f});                              // debugVoxels.frag:10 (alpha 1 like every image of this library)
        float world[3], clip[21 * 4];
        orc_debug_voxel_vertices(fp, D, id, world, clip);
        for (int k = 0; k + 2 < 21; ++k) {
            RV cv[3];
            for (int j = 0; j < 3; ++j) { const float* q = clip + 4 * (k + j); cv[j] = {q[0], q[1], q[2], q[3]}; }
            if (k & 1) std::swap(cv[0], cv[1]);                                               // odd strip triangle: winding reversed
            ClipTri ct[2]; const int nct = clip_near(cv, ct);
            for (int q = 0; q < nct; ++q) {
                Setup s = tri_setup(ct[q].v, W, H, true);
                if (!s.valid) continue;
                const float z[3] = {ct[q].v[0].z / ct[q].v[0].w, ct[q].v[1].z / ct[q].v[1].w, ct[q].v[2].z / ct[q].v[2].w};
                raster(s, 0, H - 1, [&](int px, int py, const float l[3]) {
                    const float zn = interp(l, z[0], z[1], z[2]);
                    if (zn > 1.0f) return;                                                    // far plane
                    const float dw = zn * 0.5f + 0.5f;
                    float& dref = depth[(size_t)py * W + px];
                    if (dw < dref) { dref = dw; image[(size_t)py * W + px] = word; }          // GL_LESS
                });
            }
        }
    }
}
