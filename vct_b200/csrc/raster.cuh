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

// Sample positions of a raster pass in 1/256 pixel from the pixel's lower-left corner.  Pixel centres: n = 1 at (128, 128).  With
// GL_MULTISAMPLE (the reference's default "conservative" voxelisation, Application.cpp:673-678) a fragment exists where ANY sample is
// covered; x_min .. y_max bound the offsets for the pixel boxes.
struct SampleSet { int n; int x[4], y[4]; int x_min, x_max, y_min, y_max; };

__device__ __forceinline__ bool tri_setup(const RV v[3], int W, int H, bool cull_back, TriSetup& s, const SampleSet* ss = nullptr) {
    int X[3], Y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {                                          // x / 1 == x exactly: orthographic views (w == 1) skip the IEEE division
        const bool unit = v[i].w == 1.0f;
        X[i] = snap_coord(unit ? v[i].x : v[i].x / v[i].w, W); Y[i] = snap_coord(unit ? v[i].y : v[i].y / v[i].w, H);
    }
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
        s.z[k] = v[k].w == 1.0f ? v[k].z : v[k].z / v[k].w;
    }
    const int minx = min(s.X[0], min(s.X[1], s.X[2])), maxx = max(s.X[0], max(s.X[1], s.X[2]));
    const int miny = min(s.Y[0], min(s.Y[1], s.Y[2])), maxy = max(s.Y[0], max(s.Y[1], s.Y[2]));
    // pixel p holds a sample inside [min, max] iff 256 p + offset does for some sample: lower bound from the largest offset, upper from the smallest
    const int ox_hi = ss ? ss->x_max : 128, ox_lo = ss ? ss->x_min : 128, oy_hi = ss ? ss->y_max : 128, oy_lo = ss ? ss->y_min : 128;
    s.x0 = max(0, cdiv256((long long)minx - ox_hi)); s.x1 = min(W - 1, fdiv256((long long)maxx - ox_lo));
    s.y0 = max(0, cdiv256((long long)miny - oy_hi)); s.y1 = min(H - 1, fdiv256((long long)maxy - oy_lo));
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
// near-plane clip (z >= -w), intersections computed from the inside vertex (watertight across shared edges)
__device__ __forceinline__ int clip_near(const RV in[3], RV out[2][3]) {
    float d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = in[i].z + in[i].w;
    RV poly[4]; int n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        const bool ii = d[i] >= 0.0f, jj = d[j] >= 0.0f;
        if (ii) poly[n++] = in[i];
        if (ii != jj) {
            const RV a = ii ? in[i] : in[j], b = ii ? in[j] : in[i];
            const float da = ii ? d[i] : d[j], db = ii ? d[j] : d[i];
            const float t = da / (da - db);
            RV r; r.x = (b.x - a.x) * t + a.x; r.y = (b.y - a.y) * t + a.y; r.z = (b.z - a.z) * t + a.z; r.w = (b.w - a.w) * t + a.w;
            poly[n++] = r;
        }
    }
    if (n < 3) return 0;
    out[0][0] = poly[0]; out[0][1] = poly[1]; out[0][2] = poly[2];
    if (n == 4) { out[1][0] = poly[0]; out[1][1] = poly[2]; out[1][2] = poly[3]; return 2; }
    return 1;
}

// Multisample coverage (OpenGL 4.5 section 14.6.6): true if any sample of pixel (px, py) is covered by the triangle clipped to -1 <= z <= 1
// (the same top-left rule and the same barycentric formula per sample; z(sample) = sum l_i z_i like the centre's clip test).  l = the
// barycentrics AT THE PIXEL CENTRE, where the fragment's inputs are interpolated — extrapolated if the centre itself is outside.
__device__ __forceinline__ bool tri_cover_any(const TriSetup& s, int px, int py, const SampleSet& ss, float l[3]) {
    const float fa = (float)s.area;
    bool any = false;
    for (int i = 0; i < ss.n; ++i) {
        const long long Px = 256ll * px + ss.x[i], Py = 256ll * py + ss.y[i];
        long long E[3]; bool in = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int a = (k + 1) % 3, b = (k + 2) % 3;
            E[k] = (long long)(s.X[b] - s.X[a]) * (Py - s.Y[a]) - (long long)(s.Y[b] - s.Y[a]) * (Px - s.X[a]);
            in = in && E[k] + s.bias[k] >= 0;
        }
        if (!in) continue;
        const float e0 = (float)E[0] / fa, e1 = (float)E[1] / fa, e2 = (float)E[2] / fa;
        const float z = (e0 * s.z[0] + (s.swapped ? e2 : e1) * s.z[1]) + (s.swapped ? e1 : e2) * s.z[2];
        if (z < -1.0f || z > 1.0f) continue;
        any = true;
    }
    if (!any) return false;
    const long long Px = 256ll * px + 128, Py = 256ll * py + 128;
    float e[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        e[k] = (float)((long long)(s.X[b] - s.X[a]) * (Py - s.Y[a]) - (long long)(s.Y[b] - s.Y[a]) * (Px - s.X[a])) / fa;
    }
    l[0] = e[0]; l[1] = s.swapped ? e[2] : e[1]; l[2] = s.swapped ? e[1] : e[2];
    return true;
}
// ---------------------------------------------------------------------------------------- tile work queue
// Rasterisation work is cut into 8x4-pixel tiles = 32 pixels = one warp with ONE LANE PER PIXEL.  A tile item is
// 8 bytes: (setup slot, origin x | y << 16); origins are absolute pixels and need not be grid aligned — a
// triangle's tiles start at its bounding-box corner.  Three kernels keep every warp's serial work short:
//   bin     one thread per triangle: setup; single-tile triangles push their tile directly; multi-tile ones push
//           EXPAND items, each covering a band of tile rows with at most ~kExpandTiles tiles
//   expand  one warp per expand item: enumerates the band's tiles 32 at a time, drops trivially rejected ones,
//           pushes the rest (one atomic per 32 candidates)
//   tiles   one warp per tile
constexpr int kTileW = 8, kTileH = 4;
constexpr int kExpandTiles = 512;

struct TileQueues {
    uint2* tiles; unsigned tile_cap; unsigned* tile_count;
    uint2* expand; unsigned expand_cap; unsigned* expand_count;
    uint2* pixels; unsigned pixel_cap; unsigned* pixel_count;      // single pixels of tiny triangles: (slot, x | y << 16)
    Counters* counters;             // vct_flag_overflow
    // sharded frame: which pixels this rank rasterises.  Camera pass (own_band == 0): its 64x64 screen tiles (tile (tx, ty) belongs to rank
    // (tx + ty) mod N, like the cone trace).  Shadow pass (own_band > 0): its band of own_band rows.  own_world <= 1: no filter.
    int own_world, own_rank, own_band;
};
__device__ __forceinline__ bool pixel_owned(const TileQueues& q, int px, int py) {
    if (q.own_world <= 1) return true;
    if (q.own_band) return py / q.own_band == q.own_rank;
    return ((px >> 6) + (py >> 6)) % q.own_world == q.own_rank;
}
// may the 8x4 raster tile at (ox, oy), clipped to (x1, y1), hold a pixel this rank owns?  (it touches at most 2x2 screen tiles)
__device__ __forceinline__ bool tile_owned(const TileQueues& q, int ox, int oy, int x1, int y1) {
    if (q.own_world <= 1) return true;
    const int ex = min(ox + 7, x1), ey = min(oy + 3, y1);
    return pixel_owned(q, ox, oy) || pixel_owned(q, ex, oy) || pixel_owned(q, ox, ey) || pixel_owned(q, ex, ey);
}

// Is the box of sample points of pixels [bx0,bx1]x[by0,by1] (pixel centres without a sample set) entirely outside one of the edges?
__device__ __forceinline__ bool tile_rejected(const TriSetup& s, int bx0, int by0, int bx1, int by1, const SampleSet* ss = nullptr) {
    const int ox_hi = ss ? ss->x_max : 128, ox_lo = ss ? ss->x_min : 128, oy_hi = ss ? ss->y_max : 128, oy_lo = ss ? ss->y_min : 128;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const long long ex = s.X[b] - s.X[a], ey = s.Y[b] - s.Y[a];
        // E = ex*(Py - Ya) - ey*(Px - Xa) is maximised at Py = (ex>0 ? top : bottom), Px = (ey>0 ? left : right)
        const long long Py = ex > 0 ? 256ll * by1 + oy_hi : 256ll * by0 + oy_lo, Px = ey > 0 ? 256ll * bx0 + ox_lo : 256ll * bx1 + ox_hi;
        if (ex * (Py - s.Y[a]) - ey * (Px - s.X[a]) + s.bias[k] < 0) return true;
    }
    return false;
}

// Called by a CONVERGED warp.  Lanes with `queued` own a triangle setup `s` already published in slot `sslot`.
__device__ __forceinline__ void enqueue_tiles(bool queued, const TriSetup& s, uint32_t sslot, const TileQueues& q, int max_tiles = kExpandTiles) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int bw = queued ? s.x1 - s.x0 + 1 : 0, bh = queued ? s.y1 - s.y0 + 1 : 0;
    const int ntx = (bw + kTileW - 1) / kTileW, nty = (bh + kTileH - 1) / kTileH;
    const bool single = queued && ntx * nty == 1 && tile_owned(q, s.x0, s.y0, s.x1, s.y1);
    const unsigned sm = __ballot_sync(0xffffffffu, single);
    if (sm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(q.tile_count, (unsigned)__popc(sm));
        const uint32_t pos = __shfl_sync(0xffffffffu, base, 0) + __popc(sm & lt_mask);
        if (single) { if (pos < q.tile_cap) q.tiles[pos] = make_uint2(sslot, (unsigned)s.x0 | (unsigned)s.y0 << 16); else vct_flag_overflow(q.counters); }
    }
    // multi-tile: bands of tile rows, <= max_tiles tiles each (at least one row)
    const bool multi = queued && ntx * nty > 1;
    const int rows = multi ? max(1, max_tiles / ntx) : 1;
    int nitems = multi ? (nty + rows - 1) / rows : 0;
    int inc = nitems;                                                       // warp prefix sum -> one atomic per warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (!total) return;
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(q.expand_count, (unsigned)total);
    uint32_t pos = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(inc - nitems);
    for (int r = 0; r < nty && nitems; r += rows, ++pos) {
        if (pos < q.expand_cap) q.expand[pos] = make_uint2(sslot, (unsigned)r | (unsigned)min(rows, nty - r) << 16); else vct_flag_overflow(q.counters);
    }
}
// expand: one warp per item.  `setups` is an array of records of `stride` bytes that begin with a TriSetup.
__device__ __forceinline__ void expand_items(const unsigned char* __restrict__ setups, size_t stride, const TileQueues& q) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned n_items = min(*q.expand_count, q.expand_cap);
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < n_items; item += warps) {
        const uint2 it = __ldg(q.expand + item);
        const TriSetup s = *reinterpret_cast<const TriSetup*>(setups + (size_t)it.x * stride);
        const int ntx = (s.x1 - s.x0 + kTileW) / kTileW;
        const int r0 = (int)(it.y & 0xFFFFu), nr = (int)(it.y >> 16), nt = ntx * nr;
        int ty = 0, tx = lane;                                              // tile (tx, ty) of this lane, advanced without divisions
        while (tx >= ntx) { tx -= ntx; ty++; }
        for (int tb = 0; tb < nt; tb += 32) {
            bool keep = false; int ox = 0, oy = 0;
            if (tb + lane < nt) {
                ox = s.x0 + tx * kTileW; oy = s.y0 + (r0 + ty) * kTileH;
                keep = tile_owned(q, ox, oy, s.x1, s.y1) && !tile_rejected(s, ox, oy, min(ox + kTileW - 1, s.x1), min(oy + kTileH - 1, s.y1));
            }
            tx += 32; while (tx >= ntx) { tx -= ntx; ty++; }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (!m) continue;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(q.tile_count, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const uint32_t pos = base + __popc(m & lt_mask);
                if (pos < q.tile_cap) q.tiles[pos] = make_uint2(it.x, (unsigned)ox | (unsigned)oy << 16); else vct_flag_overflow(q.counters);
            }
        }
    }
}
static inline TileQueues vctk_tile_queues(vct_ctx* c) {
    TileQueues q;
    q.tiles = reinterpret_cast<uint2*>(c->d_tile_queue); q.tile_cap = (unsigned)c->tile_queue_cap; q.tile_count = &c->d_counters->tile_queue_count;
    q.expand = reinterpret_cast<uint2*>(c->d_expand_queue); q.expand_cap = (unsigned)c->expand_cap; q.expand_count = &c->d_counters->expand_count;
    q.pixels = reinterpret_cast<uint2*>(c->d_pixel_queue); q.pixel_cap = (unsigned)c->pixel_cap; q.pixel_count = &c->d_counters->pixel_count;
    q.counters = c->d_counters;
    q.own_world = 0; q.own_rank = 0; q.own_band = 0;
    return q;
}
// reserve one setup slot per lane with `want` (one atomic per warp); returns the lane's slot
__device__ __forceinline__ uint32_t reserve_slots(bool want, unsigned* __restrict__ counter) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return 0u;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned)__popc(m));
    return __shfl_sync(0xffffffffu, base, 0) + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ float interp1(const float l[3], float a0, float a1, float a2) { return (l[0] * a0 + l[1] * a1) + l[2] * a2; }
__device__ __forceinline__ V3 interp3(const float l[3], V3 a, V3 b, V3 c) {
    return mk3(interp1(l, a.x, b.x, c.x), interp1(l, a.y, b.y, c.y), interp1(l, a.z, b.z, c.z));
}

// ------------------------------------------------------------------------------------ vertex transform
// voxelize.vert:15-23 / phong.vert:37-54: world position, normalMatrix*normal, Gram-Schmidt tangent, bitangent.
// A device function so that the frame-begin launch (volume_passes.cu) can run it next to the sparse clear.
__device__ __forceinline__ void transform_vertices_part(const float* __restrict__ verts, const int32_t* __restrict__ vactor, const Mat4* __restrict__ models,
                                                        const float* __restrict__ nmats, size_t n, float4* __restrict__ wpos, float4* __restrict__ wnrm,
                                                        float4* __restrict__ wT, float4* __restrict__ wB, unsigned block, unsigned n_blocks) {
    for (size_t i = block * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)n_blocks * blockDim.x) {
        const float* v = verts + 14 * i;
        const int a = vactor[i];
        const V4 w = mul44(models[a], mk4(v[0], v[1], v[2], 1.0f));
        const float* m = nmats + 9 * a;
        auto mul3 = [&](V3 q) { return mk3((m[0] * q.x + m[1] * q.y) + m[2] * q.z, (m[3] * q.x + m[4] * q.y) + m[5] * q.z, (m[6] * q.x + m[7] * q.y) + m[8] * q.z); };
        const V3 N = mul3(mk3(v[3], v[4], v[5]));
        V3 T = mul3(mk3(v[8], v[9], v[10]));
        T = normalize3(T - N * dot3(T, N));                               // phong.vert:52
        const V3 B = cross3(N, T);                                        // phong.vert:53
        wpos[i] = make_float4(w.x, w.y, w.z, 1.0f);
        wnrm[i] = make_float4(N.x, N.y, N.z, 0.0f);
        wT[i] = make_float4(T.x, T.y, T.z, 0.0f);
        wB[i] = make_float4(B.x, B.y, B.z, 0.0f);
    }
}

// ---------------------------------------------------------------------------------- 2D material textures
// min LINEAR_MIPMAP_NEAREST, mag NEAREST, REPEAT (reference src/Graphics/GLHelper.cpp:180-183); rho2 is the
// squared GL scale factor; level selection by comparisons only (no log2).
__device__ __forceinline__ int wrapi(int i, int n) {
    if ((n & (n - 1)) == 0) return i & (n - 1);                            // power-of-two sizes: REPEAT is a mask (same result)
    int r = i % n; return r < 0 ? r + n : r;
}
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
// voxelize.frag:160-184 == phong.frag:183-207.  The five LINEAR taps (offsets (0,0),(1,0),(0,1),(-1,0),(0,-1)) share a
// 4x4 texel neighbourhood: its 12 distinct texels are fetched once; every tap is then evaluated with exactly the
// arithmetic of shadow_linear().
__device__ __forceinline__ float calc_shadow_factor(const float* __restrict__ sm, int S, V4 lsp) {
    const float sx = (lsp.x / lsp.w + 1.0f) * 0.5f, sy = (lsp.y / lsp.w + 1.0f) * 0.5f, sz = (lsp.z / lsp.w + 1.0f) * 0.5f;
    const float frag_depth = sz - 0.01f;
    if (frag_depth > 1.0f) return 0.0f;
    const float x = sx * (float)S - 0.5f, y = sy * (float)S - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f)) return 0.0f;          // every tap reads the border (1.0): frag_depth <= 1 is never greater
    const int x0 = (int)fx0, y0 = (int)fy0; const float fx = x - fx0, fy = y - fy0;
    float t[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) t[j][i] = ((i == 0 || i == 3) && (j == 0 || j == 3)) ? 0.0f : shadow_texel(sm, S, x0 - 1 + i, y0 - 1 + j);
    auto tap = [&](int ox, int oy) {
        const int i = 1 + ox, j = 1 + oy;
        const float top = t[j][i] * (1.0f - fx) + t[j][i + 1] * fx;
        const float bot = t[j + 1][i] * (1.0f - fx) + t[j + 1][i + 1] * fx;
        return top * (1.0f - fy) + bot * fy;
    };
    float f = 0.0f;
    if (frag_depth > tap(0, 0)) f += 1.0f;
    if (frag_depth > tap(1, 0)) f += 1.0f;
    if (frag_depth > tap(0, 1)) f += 1.0f;
    if (frag_depth > tap(-1, 0)) f += 1.0f;
    if (frag_depth > tap(0, -1)) f += 1.0f;
    return f / 5.0f;
}

// ------------------------------------------------------------------------------ warpmap (exact software path)
// 32^3 RGBA16 unorm, LINEAR, CLAMP_TO_EDGE (reference src/Application.cpp:383-389)
// The float copy behind the unorm16 texels (k_warpmap_floats): q / 65535 per channel, the IEEE division done once per texel.  One entry = texel
// (x, y, z) and its +x neighbour (CLAMP_TO_EDGE), 32 bytes: one 256-bit load (LDG.E.256, sm_100) fetches both x corners of a trilinear cell.
// x, y, z already clamped.
__device__ __forceinline__ void warp_texel_pair(const uint16_t* __restrict__ wm, int x, int y, int z, V3& a, V3& b) {
    const int n = VCT_WARP_DIM;
    const float4* p = reinterpret_cast<const float4*>(wm + 4 * (size_t)n * n * n) + 2 * (((size_t)z * n + y) * n + x);
    float ax, ay, az, aw, bx, by, bz, bw;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(ax), "=f"(ay), "=f"(az), "=f"(aw), "=f"(bx), "=f"(by), "=f"(bz), "=f"(bw) : "l"(p));
    a = mk3(ax, ay, az); b = mk3(bx, by, bz);
}
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, float t) { const float s = 1.0f - t; return mk3(a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t); }
__device__ __forceinline__ V3 warp_sample(const uint16_t* __restrict__ wm, V3 tc) {
    const float n = (float)VCT_WARP_DIM;
    const float x = tc.x * n - 0.5f, y = tc.y * n - 0.5f, z = tc.z * n - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    if (!(fabsf(fx0) < 1e9f) || !(fabsf(fy0) < 1e9f) || !(fabsf(fz0) < 1e9f)) return mk3(0.f, 0.f, 0.f);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0; const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
    const int n1 = VCT_WARP_DIM - 1;
    const int xc = min(max(x0, 0), n1), ya = min(max(y0, 0), n1), yb = min(max(y0 + 1, 0), n1), za = min(max(z0, 0), n1), zb = min(max(z0 + 1, 0), n1);
    V3 a00, b00, a10, b10, a01, b01, a11, b11;
    warp_texel_pair(wm, xc, ya, za, a00, b00); warp_texel_pair(wm, xc, yb, za, a10, b10);
    warp_texel_pair(wm, xc, ya, zb, a01, b01); warp_texel_pair(wm, xc, yb, zb, a11, b11);
    if (x0 < 0) { b00 = a00; b10 = a10; b01 = a01; b11 = a11; }                // both x corners clamp to texel 0
    else if (x0 > n1) { a00 = b00; a10 = b10; a01 = b01; a11 = b11; }          // (x0 >= n: both clamp to the last texel, whose entry holds it twice)
    V3 c00 = lerp3(a00, b00, fx), c10 = lerp3(a10, b10, fx), c01 = lerp3(a01, b01, fx), c11 = lerp3(a11, b11, fx);
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
// common.glsl:37-42 (voxelizeTesselationWarp): the camera frustum as voxel grid, (pv * P).xyz / w * 0.5 + 0.5
__device__ __forceinline__ V3 tess_warp_position(V3 pos, const vct_frame_params& fp) {
    const float* m = fp.pv;
    const float qx = ((m[0] * pos.x + m[4] * pos.y) + m[8] * pos.z) + m[12];
    const float qy = ((m[1] * pos.x + m[5] * pos.y) + m[9] * pos.z) + m[13];
    const float qz = ((m[2] * pos.x + m[6] * pos.y) + m[10] * pos.z) + m[14];
    const float qw = ((m[3] * pos.x + m[7] * pos.y) + m[11] * pos.z) + m[15];
    return mk3((qx / qw) * 0.5f + 0.5f, (qy / qw) * 0.5f + 0.5f, (qz / qw) * 0.5f + 0.5f);
}
// common.glsl:44-60, priority warpVoxels > warpTexture > voxelizeTesselationWarp > linear
__device__ __forceinline__ V3 get_voxel_position(V3 pos, const vct_frame_params& fp, const uint16_t* __restrict__ warpmap) {
    if (fp.warp_voxels) return voxel_warp(voxel_linear_position(pos, fp), voxel_linear_position(eye_of(fp), fp));
    if (fp.warp_texture) return warp_sample(warpmap, voxel_linear_position(pos, fp));
    if (fp.voxelize_tesselation_warp) return tess_warp_position(pos, fp);
    return voxel_linear_position(pos, fp);
}
// ivec3(vec3) truncation + image bounds (out-of-bounds image access is a no-op)
__device__ __forceinline__ bool to_voxel_index(V3 p, int D, int& ix, int& iy, int& iz) {
    const float fd = (float)D;
    if (!(p.x > -1.0f) || !(p.x < fd) || !(p.y > -1.0f) || !(p.y < fd) || !(p.z > -1.0f) || !(p.z < fd)) return false;
    ix = (int)p.x; iy = (int)p.y; iz = (int)p.z;
    return true;
}
