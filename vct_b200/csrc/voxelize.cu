// voxelize.cu — dominant-axis triangle voxelisation with running-average RGBA8 accumulation
// (compiled with -fmad=false: coverage, voxel index and colours are bit-exact vs the oracle).
//
// Replaces the GL draw of voxelize.vert/.geom/.frag (reference src/Application.cpp:666-754) and the 32^3
// occupancy draw (:235-301).  Stages, all on the context stream, no host synchronisation:
//   k_transform_vertices  voxelize.vert:15-23 / phong.vert:37-54 : world position, normalMatrix*normal, T, B
//   k_voxel_bin           one thread per triangle: voxelize.geom:26-73 (axis pick, ortho projection) + fixed-point
//                         setup.  ~87 % of Sponza's triangles cover no pixel centre at 256^3 and end here.
//                         Triangles whose bounding box is <= kInlineArea pixels are finished in place; the others
//                         are cut into 8x4-pixel tiles (trivially rejected tiles skipped) pushed to a work queue.
//   k_voxel_tiles         persistent grid, one warp per tile, ONE LANE PER PIXEL: coverage, voxel index,
//                         voxelize.frag:186-280 shading, then the image atomic of the selected mode
//   k_voxel_resolve       deterministic running average only (see below)
//
// Deterministic running average.  imageAtomicRGBA8Avg (voxelize.frag:111-139) truncates on every insertion, so
// the result depends on insertion order; the oracle fixes the canonical order "triangles in draw order,
// fragments of a triangle in raster-scan order".  Instead of sorting all fragments globally, every fragment is
// appended to a per-voxel linked list (head pointers live in the cleared voxelColor volume, one atomicExch per
// fragment); k_voxel_resolve lets the thread that owns a voxel's head walk its list (Sponza: <= 10 entries), order
// it by (triangle, raster rank) in registers, replay the insertions sequentially and store the final words.
#include "raster.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kInlineArea = 4;                   // bounding boxes up to this many pixels are rasterised by the bin thread
enum { MODE_SORTED = 0, MODE_CAS = 1, MODE_MAX = 2, MODE_OCC = 3 };

// ------------------------------------------------------------------------------------ vertex transform
__global__ void __launch_bounds__(kThreads) k_transform_vertices(const float* __restrict__ verts, const int32_t* __restrict__ vactor,
                                                                 const Mat4* __restrict__ models, const float* __restrict__ nmats, size_t n,
                                                                 float4* __restrict__ wpos, float4* __restrict__ wnrm, float4* __restrict__ wT, float4* __restrict__ wB) {
    transform_vertices_part(verts, vactor, models, nmats, n, wpos, wnrm, wT, wB, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------- per-triangle
// What a tile needs to know about its triangle (written once by k_voxel_bin, broadcast-loaded by the tile warp).
struct ShadeIn { V3 w[3], n[3]; float uv[3][2]; };          // world position, normalMatrix*normal, uv per source vertex (96 B)
struct __align__(16) VoxHead {  // what coverage and the voxel index need: 128 B, broadcast-loaded as 8 x 16 bytes
    TriSetup s;               // 80 B
    float cvx[3], cvy[3];     // clip-space x,y per source vertex (w == 1 for the orthographic views; z is s.z)
    uint32_t tri; int axis; int material; float rho2;
    uint32_t pad[2];
};
struct __align__(16) VoxSetup : VoxHead {
    ShadeIn in;               // shading inputs, gathered once by the bin thread: tile warps see no dependent vertex loads
};                            // and fetch them (6 x 16 bytes) only when the tile has a fragment
static_assert(sizeof(VoxHead) == 128 && sizeof(VoxSetup) == 224, "VoxSetup layout");

struct __align__(16) Frag {   // 48 B = three 16-byte words; k_voxel_resolve reads only the first for a voxel with one fragment
    uint32_t key;             // voxel index
    uint32_t next;            // index+1 of the next fragment of the same voxel, 0 = end
    uint32_t cw1, nw1;        // the packed words this fragment yields if it stays alone in its voxel: insert(0, colour / normal)
    uint32_t tri; float cr, cg, cb;     // canonical order key 1 (draw index) | shaded colour
    uint32_t rank; float nr, ng, nb;    // canonical order key 2 (py*D+px)    | encoded normal
};
static_assert(sizeof(Frag) == 48, "Frag layout");

__device__ __forceinline__ V3 ld3(const float4* __restrict__ p, uint32_t i) { const float4 q = __ldg(p + i); return mk3(q.x, q.y, q.z); }

struct VoxArgs {
    const FrameConst* fc; int D;
    int V;                    // viewport of the raster pass in pixels: (int)(voxelizeMultiplier * D) for the voxelise pass (Application.cpp:668), D otherwise
    const uint32_t* indices; const int32_t* trimat; const float* verts; uint32_t n_tris;
    const float4 *wpos, *wnrm;
    const DevTexture* tex; const DevMaterial* mats; const float* shadow; const uint16_t* warpmap;
    VoxSetup* setups; unsigned setup_cap; TileQueues q;
    Frag* frags; unsigned frag_cap; uint8_t* displaced;       // displaced[slot] = 1: a later fragment took over the head of that voxel's list
    uint32_t *color, *normal, *occ;
    uint8_t* seg;             // segment mask of this frame (common.cuh): one byte per 8 voxels of an x-row
    int msaa; SampleSet ms;   // Settings::conservativeRasterization == MSAA: any-sample coverage (raster.cuh tri_cover_any)
    int slab_cull;            // multi-GPU, linear mapping: triangles that cannot touch this rank's z layers stop at the bin kernel
    Counters* counters;
};

// ---- long per-voxel lists (deterministic mode).  Normally a voxel holds 1-10 fragments.  A dense mesh on a coarse grid reaches a few
// hundred, and a warp mode can pile a large part of the scene into ONE voxel (the reference's warp map is zero outside the 0.8-scaled
// quad of quad.vert:12, so everything in the outer tenth of the volume is sampled towards voxel 0: 751 k of config 4's 2.0 M fragments).
// What keeps this cheap and exact: imageAtomicRGBA8Avg forgets its history whenever the 8-bit count wraps to 0 (the next insertion
// multiplies the stored colour by a weight of 0, voxelize.frag:127-133), so a voxel's final words depend only on the LAST
// r = ((N - 1) mod 256) + 1 of its N fragments in canonical order.  k_voxel_tiles counts N per voxel in the (still unused) normal
// volume; the resolve kernel handles N <= kSortMax itself and queues the rest: a warp per voxel up to kMedMax fragments (list walk,
// bitonic sort in shared memory, replay of the last r), and for the few voxels beyond that a scan of ALL fragment records instead of a
// walk of a 751 k-entry list: compact the voxel's (order key, slot) pairs, radix-select the r-th largest key, sort and replay those r.
constexpr int kMedMax = 1024;
constexpr int kHugeMax = 16;                                         // (api.cu reserves 16 entries behind the long queue)
constexpr uint32_t kWalkMax = 1u << 16;                             // longest list a warp still walks when the huge table has no room for it
struct LongEntry { uint32_t key, head, count, pad; };               // voxel, head slot of its list, fragment count
struct __align__(16) HugeItem { unsigned long long k; uint32_t slot, h; };   // order key + 1, fragment slot, index into the huge table
constexpr int kHugeBins = 4096, kHugeSmall = 16384;                 // coarse histogram over the triangle index; per-voxel short list of the gather pass
struct HugeAux { unsigned hist[kHugeMax][kHugeBins]; unsigned long long thresh[kHugeMax]; unsigned small_n[kHugeMax]; unsigned pad[8]; };
struct LongArgs { LongEntry* queue; unsigned cap; LongEntry* huge; HugeItem* items; unsigned item_cap; int inline_long; HugeAux* aux; HugeItem* small; int tri_shift; };   // inline_long: the long-list kernels do not run this frame

// voxelize.geom:26-73: axis from the summed vertex normals, projection through that axis' ortho view
__device__ __forceinline__ bool make_setup(const VoxArgs& a, const FrameConst& fc, uint32_t t, VoxSetup& S) {
    const uint32_t i0 = __ldg(a.indices + 3 * (size_t)t), i1 = __ldg(a.indices + 3 * (size_t)t + 1), i2 = __ldg(a.indices + 3 * (size_t)t + 2);
    const V3 w[3] = {ld3(a.wpos, i0), ld3(a.wpos, i1), ld3(a.wpos, i2)};
    if (a.slab_cull) {
        // Multi-GPU: a triangle whose voxel-space box (1.5 voxels of slack) lies inside the volume but misses this rank's
        // z layers can produce no fragment here — neither an owned one nor one outside the volume (those are counted by the
        // rank that owns z = 0) — so it skips setup, binning and tiles.  Linear mapping only (no warp mode).
        const float fd = (float)a.D;
        float lo[3], hi[3];
        const float wc[3][3] = {{w[0].x, w[1].x, w[2].x}, {w[0].y, w[1].y, w[2].y}, {w[0].z, w[1].z, w[2].z}};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sc = fd / (fc.p.voxel_max[k] - fc.p.voxel_min[k]), of = fc.p.voxel_center[k] + fc.p.voxel_min[k];
            const float v0 = (wc[k][0] - of) * sc, v1 = (wc[k][1] - of) * sc, v2 = (wc[k][2] - of) * sc;
            lo[k] = fminf(v0, fminf(v1, v2)) - 1.5f; hi[k] = fmaxf(v0, fmaxf(v1, v2)) + 1.5f;
        }
        const bool inside = lo[0] >= 0.0f && lo[1] >= 0.0f && lo[2] >= 0.0f && hi[0] <= fd && hi[1] <= fd && hi[2] <= fd;   // false for NaN
        if (inside && !owns_any_z(fc.st, max((int)lo[2], 0), min((int)hi[2], a.D - 1))) return false;
    }
    const V3 n0 = ld3(a.wnrm, i0), n1 = ld3(a.wnrm, i1), n2 = ld3(a.wnrm, i2);
    S.in.w[0] = w[0]; S.in.w[1] = w[1]; S.in.w[2] = w[2]; S.in.n[0] = n0; S.in.n[1] = n1; S.in.n[2] = n2;
    const V3 f = normalize3((n0 + n1) + n2);
    const float ax = fabsf(f.x), ay = fabsf(f.y), az = fabsf(f.z);
    int axis;
    if (ax > ay && ax > az) axis = 0; else if (ay > ax && ay > az) axis = 1; else axis = 2;
    if (fc.p.axis_override >= 0 && fc.p.axis_override <= 2) axis = fc.p.axis_override;
    const Mat4& mvp = axis == 0 ? fc.mvp_x : axis == 1 ? fc.mvp_y : fc.mvp_z;
    RV cv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { const V4 c = mul44(mvp, mk4(w[k].x, w[k].y, w[k].z, 1.0f)); cv[k].x = c.x; cv[k].y = c.y; cv[k].z = c.z; cv[k].w = c.w; }
    if (cv[0].w == 1.0f && cv[1].w == 1.0f && cv[2].w == 1.0f) {
        // Early out for the 80-90 % of triangles whose pixel box holds no pixel centre.  The snapped window coordinate differs from
        // (ndc * 0.5 + 0.5) * D by at most 1/512 pixel, so with a margin of 1/128 pixel an empty float box proves the exact box
        // [ceil(min - 0.5), floor(max - 0.5)] of tri_setup empty; everything else (and every NaN) takes the exact path.
        // (multisampling: the box of sample points instead — smallest offset at the upper end, largest at the lower)
        const float fd = (float)a.V, e = 0.0078125f;
        const float ox_lo = a.msaa ? (float)a.ms.x_min * 0.00390625f : 0.5f, ox_hi = a.msaa ? (float)a.ms.x_max * 0.00390625f : 0.5f;
        const float oy_lo = a.msaa ? (float)a.ms.y_min * 0.00390625f : 0.5f, oy_hi = a.msaa ? (float)a.ms.y_max * 0.00390625f : 0.5f;
        const float x0 = (cv[0].x * 0.5f + 0.5f) * fd, x1 = (cv[1].x * 0.5f + 0.5f) * fd, x2 = (cv[2].x * 0.5f + 0.5f) * fd;
        const float y0 = (cv[0].y * 0.5f + 0.5f) * fd, y1 = (cv[1].y * 0.5f + 0.5f) * fd, y2 = (cv[2].y * 0.5f + 0.5f) * fd;
        const float sum = ((x0 + x1) + x2) + ((y0 + y1) + y2);             // NaN or infinite anywhere: fminf / fmaxf would hide it, take the exact path
        if (fabsf(sum) < 1e30f) {
            if (floorf(fmaxf(x0, fmaxf(x1, x2)) - ox_lo + e) < ceilf(fminf(x0, fminf(x1, x2)) - ox_hi - e)) return false;
            if (floorf(fmaxf(y0, fmaxf(y1, y2)) - oy_lo + e) < ceilf(fminf(y0, fminf(y1, y2)) - oy_hi - e)) return false;
        }
    }
    if (!tri_setup(cv, a.V, a.V, false, S.s, a.msaa ? &a.ms : nullptr)) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) { S.cvx[k] = cv[k].x; S.cvy[k] = cv[k].y; S.s.z[k] = cv[k].z; }   // the shader interpolates gl_Position.xyz (w == 1)
    S.tri = t; S.axis = axis; S.material = 0; S.rho2 = 0.0f;
    return true;
}
// texture LOD input of the triangle (voxelize.frag:197 implicit derivatives; affine in an orthographic view)
__device__ __forceinline__ void make_shading_setup(const VoxArgs& a, VoxSetup& S) {
    S.material = __ldg(a.trimat + S.tri);
    RV cv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t vi = __ldg(a.indices + 3 * (size_t)S.tri + k);
        S.in.uv[k][0] = __ldg(a.verts + 14 * (size_t)vi + 6); S.in.uv[k][1] = __ldg(a.verts + 14 * (size_t)vi + 7);
        cv[k].x = S.cvx[k]; cv[k].y = S.cvy[k]; cv[k].z = S.s.z[k]; cv[k].w = 1.0f;
    }
    const int dt = a.mats[S.material].diffuse_tex;
    if (dt < 0) return;
    S.rho2 = tri_rho2_affine(cv, S.in.uv, a.V, a.V, a.tex[dt]);
}

// voxelize.frag:79-108
__device__ __forceinline__ bool frag_voxel(const FrameConst& fc, const VoxHead& S, const float l[3], int D, const uint16_t* __restrict__ warpmap, bool occupancy,
                                           int& ix, int& iy, int& iz) {
    const V3 ndc = mk3(interp1(l, S.cvx[0], S.cvx[1], S.cvx[2]), interp1(l, S.cvy[0], S.cvy[1], S.cvy[2]), interp1(l, S.s.z[0], S.s.z[1], S.s.z[2]));
    V3 u = mk3((ndc.x + 1.0f) * 0.5f, (ndc.y + 1.0f) * 0.5f, (ndc.z + 1.0f) * 0.5f);
    if (S.axis == 0) u = mk3(1.0f - u.z, u.y, u.x);
    else if (S.axis == 1) u = mk3(u.x, 1.0f - u.z, u.y);
    u.z = 1.0f - u.z;
    if (fc.p.warp_voxels) u = voxel_warp(u, voxel_linear_position(eye_of(fc.p), fc.p));
    else if (fc.p.warp_texture && !occupancy) u = warp_sample(warpmap, u);
    return to_voxel_index(mk3((float)D * u.x, (float)D * u.y, (float)D * u.z), D, ix, iy, iz);
}
// fragment exists (coverage + near/far clip) and belongs to this rank's slab.  MS: any-sample coverage, near/far clip per sample
template <bool MS>
__device__ __forceinline__ bool frag_test(const FrameConst& fc, const VoxHead& S, int px, int py, int D, const uint16_t* __restrict__ warpmap, bool occupancy,
                                          const SampleSet& ms, float l[3], bool& oob, int& ix, int& iy, int& iz) {
    if (MS) { if (!tri_cover_any(S.s, px, py, ms, l)) return false; }
    else {
        if (!tri_cover(S.s, px, py, l)) return false;
        const float z = interp1(l, S.s.z[0], S.s.z[1], S.s.z[2]);
        if (z < -1.0f || z > 1.0f) return false;
    }
    oob = !frag_voxel(fc, S, l, D, warpmap, occupancy, ix, iy, iz);
    if (occupancy) return true;                                           // the 32^3 occupancy grid is not sharded: every rank builds all of it
    if (oob) return fc.st.rank == 0;                                      // counted once, by the rank owning z = 0
    return owns_z(fc.st, iz);
}

struct Shaded { V3 color, nenc; };
// voxelize.frag:195-228
__device__ __forceinline__ Shaded shade_fragment(const FrameConst& fc, int material, float rho2, const ShadeIn& I, const float l[3], const DevTexture* __restrict__ tex,
                                                 const DevMaterial* __restrict__ mats, const float* __restrict__ shadow) {
    const V3 wp = interp3(l, I.w[0], I.w[1], I.w[2]);
    const V3 nn = interp3(l, I.n[0], I.n[1], I.n[2]);
    const float u = interp1(l, I.uv[0][0], I.uv[1][0], I.uv[2][0]), v = interp1(l, I.uv[0][1], I.uv[1][1], I.uv[2][1]);
    V3 col = mk3(0.f, 0.f, 0.f);
    const int dt = mats[material].diffuse_tex;
    if (dt >= 0) { const V4 a = sample2d(tex[dt], u, v, rho2); col = mk3(a.x, a.y, a.z); }
    const V3 N = normalize3(nn);
    Shaded out; out.nenc = mk3((N.x + 1.0f) * 0.5f, (N.y + 1.0f) * 0.5f, (N.z + 1.0f) * 0.5f);
    if (fc.p.voxelize_lighting) {
        V3 fin = mk3(0.f, 0.f, 0.f);
        for (int i = 0; i < fc.n_lights; ++i) {
            const vct_light& L = fc.lights[i];
            if (!L.enabled) continue;
            const V3 lc = mk3(L.color[0], L.color[1], L.color[2]), lp = mk3(L.position[0], L.position[1], L.position[2]);
            V3 lit = mk3(0.f, 0.f, 0.f);
            if (L.type == 0u) {
                const float ndl = maxsel(0.0f, dot3(N, normalize3(lp - wp)));
                lit = ((col * L.intensity) * lc) * ndl;
                const float e0 = 0.75f * L.range, e1 = L.range, dist = length3(lp - wp);
                const float tt = clampf((dist - e0) / (e1 - e0), 0.0f, 1.0f);
                lit = lit * (1.0f - tt * tt * (3.0f - 2.0f * tt));
            } else if (L.type == 1u) {
                const float ndl = maxsel(0.0f, dot3(N, normalize3(mk3(-L.direction[0], -L.direction[1], -L.direction[2]))));
                lit = ((col * lc) * L.intensity) * ndl;
            }
            if (L.shadow_caster) {
                const float sf = 1.0f - calc_shadow_factor(shadow, fc.S, mul44(fc.ls, mk4(wp.x, wp.y, wp.z, 1.0f)));
                lit = lit * sf;
            }
            fin = fin + lit;
        }
        col = mk3(clampf(fin.x, 0.0f, 1.0f), clampf(fin.y, 0.0f, 1.0f), clampf(fin.z, 0.0f, 1.0f));
    }
    out.color = col;
    return out;
}

// voxelize.frag:111-139 — one sequential insertion
__device__ __forceinline__ uint32_t rgba8_avg_insert(uint32_t stored, float r, float g, float b) {
    const float vr = r * 255.0f, vg = g * 255.0f, vb = b * 255.0f;
    if (stored == 0u) return 1u << 24 | (f2u_trunc(vb) & 255u) << 16 | (f2u_trunc(vg) & 255u) << 8 | (f2u_trunc(vr) & 255u);
    float rr = (float)(stored & 255u), rg = (float)((stored >> 8) & 255u), rb = (float)((stored >> 16) & 255u);
    const float rw = (float)(stored >> 24);
    rr *= rw; rg *= rw; rb *= rw;
    float cr = rr + vr, cg = rg + vg, cb = rb + vb; const float cw = rw + 1.0f;
    cr /= cw; cg /= cw; cb /= cw;
    return (f2u_trunc(cw) & 255u) << 24 | (f2u_trunc(cb) & 255u) << 16 | (f2u_trunc(cg) & 255u) << 8 | (f2u_trunc(cr) & 255u);
}
// free-running mode: the GLSL CAS loop as written
__device__ __forceinline__ void rgba8_avg_atomic(uint32_t* addr, float r, float g, float b) {
    uint32_t prev = 0u, nv = rgba8_avg_insert(0u, r, g, b), cur;
    while ((cur = atomicCAS(addr, prev, nv)) != prev) { prev = cur; nv = rgba8_avg_insert(cur, r, g, b); }
}

// the image atomic of the selected mode; `slot` is the fragment record reserved for MODE_SORTED
template <int MODE>
__device__ __forceinline__ void store_fragment(const VoxArgs& a, uint32_t tri, const Shaded& sh, int D, int px, int py, int ix, int iy, int iz, uint32_t slot, unsigned batch_mask) {
    const uint32_t o = (uint32_t)(((size_t)iz * D + iy) * D + ix);
    a.seg[o >> 3] = 1;                                                      // every writer stores the same byte
    if (MODE == MODE_SORTED) {
        Frag f;
        f.key = o; f.tri = tri; f.rank = (uint32_t)(py * a.V + px);
        // Warp-aggregated push: the lanes of this batch that hit the same voxel chain their records among themselves (slots are
        // consecutive: base + lane) and only the first of them touches the voxel — one atomicExch on the list head and one
        // atomicAdd on the fragment count (kept in voxelNormal until the resolve overwrites it) per voxel per warp.
        const unsigned grp = __match_any_sync(batch_mask, o);
        const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
        const unsigned above = grp & ~((2u << lane) - 1u);                 // group members in higher lanes
        uint32_t prev = 0u;
        if (lane == leader) { prev = atomicExch(a.color + o, slot + 1u); atomicAdd(a.normal + o, (unsigned)__popc(grp)); }
        prev = __shfl_sync(batch_mask, prev, leader);
        f.next = above ? slot + (uint32_t)(__ffs(above) - 1 - lane) + 1u : prev;
        if (lane != leader) a.displaced[slot] = 1;                         // only the leader's record can be a list head
        else if (prev) a.displaced[prev - 1u] = 1;                         // the previous head is no head any more
        f.cr = sh.color.x; f.cg = sh.color.y; f.cb = sh.color.z; f.cw1 = rgba8_avg_insert(0u, sh.color.x, sh.color.y, sh.color.z);
        f.nr = sh.nenc.x; f.ng = sh.nenc.y; f.nb = sh.nenc.z; f.nw1 = rgba8_avg_insert(0u, sh.nenc.x, sh.nenc.y, sh.nenc.z);
        {   // streaming store of the 48-byte record (read once by k_voxel_resolve)
            const uint4* src = reinterpret_cast<const uint4*>(&f); uint4* dst = reinterpret_cast<uint4*>(a.frags + slot);
            __stcs(dst, src[0]); __stcs(dst + 1, src[1]); __stcs(dst + 2, src[2]);
        }
    } else if (MODE == MODE_CAS) {
        rgba8_avg_atomic(a.color + o, sh.color.x, sh.color.y, sh.color.z);
        rgba8_avg_atomic(a.normal + o, sh.nenc.x, sh.nenc.y, sh.nenc.z);
    } else if (MODE == MODE_MAX) {                                         // voxelize.frag:271-274
        atomicMax(a.color + o, pack_unorm(mk4(sh.color.x, sh.color.y, sh.color.z, 1.0f)));
        atomicMax(a.normal + o, pack_unorm(mk4(sh.nenc.x, sh.nenc.y, sh.nenc.z, 1.0f)));
    }
}

// ------------------------------------------------------------------------------------------------- bin
// One thread per triangle: set-up, then the triangle's 8x4 tiles go to the tile queue.  A triangle whose pixel box fits one tile
// pushes that tile itself (Sponza at 256^3: 50 k of the 56 k triangles that cover a pixel centre).  The others are listed in shared
// memory and the whole CTA enumerates their tiles together — tile index -> triangle by binary search over a prefix sum, 32 tiles per
// warp step, trivially rejected tiles dropped — so one large triangle does not serialise in its thread and no separate expand launch
// (and grid drain) is needed.  Nothing is rasterised here: coverage belongs to k_voxel_tiles.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 3) k_voxel_bin(VoxArgs a) {
    const FrameConst& fc = *a.fc;
    __shared__ uint32_t s_slot[kThreads];                   // multi-tile triangles of this round: setup slot
    __shared__ uint32_t s_pref[kThreads + 1];               // exclusive prefix sum of their tile counts
    __shared__ uint32_t s_wsum[kThreads / 32];
    __shared__ unsigned s_n[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x < 2) s_n[threadIdx.x] = 0u;
    __syncthreads();
    // Triangle t = (round * kThreads + thread) * gridDim + block: CONSECUTIVE triangles go to consecutive CTAs.  The large triangles of
    // a scene come in runs (a wall, a floor: neighbours in the index buffer); spread over the grid their tiles are enumerated by many
    // CTAs at once instead of queueing up in one (measured with block-contiguous triangles: 102 us for Sponza, the CTA that owned the
    // atrium floor finished last; with warp-contiguous groups of 32: 85 us, config 4 256 us against 66 us).  The price is uncoalesced
    // index loads (12 bytes per thread, a grid apart), 3 MB in all.
    const uint32_t per_round = gridDim.x * kThreads;
    const uint32_t n_rounds = (a.n_tris + per_round - 1u) / per_round;             // the same for every thread: the barriers below stay converged
    int round = 0;
    for (uint32_t it = 0; it < n_rounds; ++it, round ^= 1) {
        const uint32_t t = (it * kThreads + threadIdx.x) * gridDim.x + blockIdx.x;
        VoxSetup S;
        const bool valid = t < a.n_tris && make_setup(a, fc, t, S);
        const uint32_t sslot = reserve_slots(valid, &a.counters->setup_count);
        bool stored = false;
        if (valid) {
            if (sslot < a.setup_cap) { if (MODE != MODE_OCC) make_shading_setup(a, S); a.setups[sslot] = S; stored = true; }
            else vct_flag_overflow(a.counters);
        }
        const int bw = stored ? S.s.x1 - S.s.x0 + 1 : 0, bh = stored ? S.s.y1 - S.s.y0 + 1 : 0;
        const int ntx = (bw + kTileW - 1) / kTileW, nty = (bh + kTileH - 1) / kTileH;
        // ---- single-tile triangles: the small-item queue (a handful of candidate pixels each), one push per warp
        const bool single = stored && ntx * nty == 1;
        const unsigned sm = __ballot_sync(0xffffffffu, single);
        if (sm) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.q.pixel_count, (unsigned)__popc(sm));
            const uint32_t pos = __shfl_sync(0xffffffffu, base, 0) + __popc(sm & lt_mask);
            if (single) { if (pos < a.q.pixel_cap) a.q.pixels[pos] = make_uint2(sslot, (unsigned)S.s.x0 | (unsigned)S.s.y0 << 16); else vct_flag_overflow(a.counters); }
        }
        // ---- multi-tile triangles: listed, then expanded by the whole CTA
        const bool multi = stored && ntx * nty > 1;
        uint32_t li = 0;
        if (multi) { li = atomicAdd(&s_n[round], 1u); s_slot[li] = sslot; s_pref[li] = (uint32_t)(ntx * nty); }
        __syncthreads();                                     // list complete; the setups written above are visible to the CTA
        const unsigned n_multi = s_n[round];
        if (threadIdx.x == 0) s_n[round ^ 1] = 0u;           // next round's counter (its pushes come after the barrier at the end)
        if (n_multi) {                                       // uniform over the CTA
            // exclusive prefix sum of the tile counts (n_multi <= kThreads)
            const uint32_t mine = threadIdx.x < n_multi ? s_pref[threadIdx.x] : 0u;
            uint32_t inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            uint32_t woff = 0;
            for (int k = 0; k < wid; ++k) woff += s_wsum[k];
            uint32_t total = 0;
            for (int k = 0; k < kThreads / 32; ++k) total += s_wsum[k];
            __syncthreads();                                 // every thread has read its own count
            if (threadIdx.x < n_multi) s_pref[threadIdx.x] = woff + inc - mine;
            if (threadIdx.x == 0) s_pref[n_multi] = total;
            __syncthreads();
            for (uint32_t c0 = wid * 32u; c0 < total; c0 += kThreads) {
                const uint32_t c = c0 + lane;
                bool keep = false; int ox = 0, oy = 0; uint32_t slot = 0;
                if (c < total) {
                    int lo = 0, hi = (int)n_multi;           // largest j with s_pref[j] <= c
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_pref[mid] <= c) lo = mid; else hi = mid; }
                    slot = s_slot[lo];
                    const TriSetup ts = a.setups[slot].s;
                    const int tntx = (ts.x1 - ts.x0 + kTileW) / kTileW;
                    const int local = (int)(c - s_pref[lo]), ty = local / tntx, tx = local - ty * tntx;
                    ox = ts.x0 + tx * kTileW; oy = ts.y0 + ty * kTileH;
                    keep = !tile_rejected(ts, ox, oy, min(ox + kTileW - 1, ts.x1), min(oy + kTileH - 1, ts.y1), a.msaa ? &a.ms : nullptr);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (!m) continue;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(a.q.tile_count, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    const uint32_t pos = base + __popc(m & lt_mask);
                    if (pos < a.q.tile_cap) a.q.tiles[pos] = make_uint2(slot, (unsigned)ox | (unsigned)oy << 16); else vct_flag_overflow(a.counters);
                }
            }
        }
        __syncthreads();                                     // list and prefix are free for the next round
    }
}

// ----------------------------------------------------------------------------------------------- tiles
// Coverage and shading with every lane busy.  A tile item is (setup slot, origin); its candidate pixels are the part of the 8x4
// tile inside the triangle's pixel box — 2 on average for the many small triangles, 32 for tiles inside large ones.  A warp takes
// 32 items, one per lane, and walks the concatenation of their candidates 32 at a time (candidate -> item by a 5-step search over the
// prefix sum held in the lanes): coverage, near/far clip and the voxel index of voxelize.frag:79-108 run on full warps whatever the
// triangle sizes.  Fragments are collected in a per-warp ring in shared memory and shaded (voxelize.frag:195-228: albedo fetch,
// lighting, 5-tap PCF) 32 at a time, so the expensive part runs on full warps too; the image atomic of the selected mode follows.
struct Hit { uint32_t slot, pxy, vox; float l0, l1, l2; };          // vox = ix | iy << 10 | iz << 20 (dim <= 1024)
constexpr int kRing = 64;
// (not inlined: one copy of the shading code per kernel — with three inlined copies the tile kernel stalled on instruction fetch,
// `no_instruction` 6.1 warps per issue in profiles/r02g)
template <int MODE>
__device__ __noinline__ void shade_batch(const VoxArgs& a, const FrameConst& fc, const Hit* __restrict__ ring, int head, int n) {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (MODE == MODE_SORTED) {
        if (lane == 0) base = atomicAdd(&a.counters->n_frag_slots, (unsigned)n);
        base = __shfl_sync(0xffffffffu, base, 0);
    }
    if (MODE == MODE_SORTED && base + (uint32_t)n > a.frag_cap) { if (lane == 0) vct_flag_overflow(a.counters); return; }   // the fragment buffer is full: drop the batch
    if (lane >= n) return;
    const Hit h = ring[(head + lane) & (kRing - 1)];
    const VoxSetup* __restrict__ sp = a.setups + h.slot;
    const uint32_t tri = sp->tri; const int material = sp->material; const float rho2 = sp->rho2;
    const ShadeIn I = sp->in;
    const float l[3] = {h.l0, h.l1, h.l2};
    const Shaded sh = shade_fragment(fc, material, rho2, I, l, a.tex, a.mats, a.shadow);
    store_fragment<MODE>(a, tri, sh, a.D, (int)(h.pxy & 0xFFFFu), (int)(h.pxy >> 16), (int)(h.vox & 1023u), (int)((h.vox >> 10) & 1023u), (int)(h.vox >> 20), base + (uint32_t)lane,
                         n >= 32 ? 0xffffffffu : (1u << n) - 1u);
}
template <int MODE, bool MS>
__global__ void __launch_bounds__(kThreads, 4) k_voxel_tiles(VoxArgs a) {
    const FrameConst& fc = *a.fc;
    __shared__ Hit s_ring[kThreads / 32][kRing];
    const int D = a.D, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool occupancy = MODE == MODE_OCC;
    Hit* ring = s_ring[wid];
    const unsigned warps = gridDim.x * (kThreads / 32), gw = blockIdx.x * (kThreads / 32) + wid;
    unsigned counted = 0; int head = 0, tail = 0;
    // one candidate pixel per lane: coverage, clip, voxel index; fragments go to the ring, a full ring batch is shaded
    auto candidate = [&](bool cv, uint32_t slot, int px, int py) {
        float l[3]; bool oob = false; int ix = 0, iy = 0, iz = 0;
        bool frag = false;
        if (cv) {
            const VoxHead S = a.setups[slot];
            frag = px <= S.s.x1 && py <= S.s.y1 && frag_test<MS>(fc, S, px, py, D, a.warpmap, occupancy, a.ms, l, oob, ix, iy, iz);
        }
        const bool hit = frag && !oob;
        if (frag) counted++;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (!m) return;
        if (MODE == MODE_OCC) { if (hit) atomicOr(a.occ + ((size_t)iz * D + iy) * D + ix, 1u); return; }
        if (hit) {
            Hit h; h.slot = slot; h.pxy = (uint32_t)px | (uint32_t)py << 16; h.vox = (uint32_t)ix | (uint32_t)iy << 10 | (uint32_t)iz << 20;
            h.l0 = l[0]; h.l1 = l[1]; h.l2 = l[2];
            ring[(tail + __popc(m & lt_mask)) & (kRing - 1)] = h;
        }
        tail += __popc(m);
        __syncwarp();
        if (tail - head >= 32) { shade_batch<MODE>(a, fc, ring, head, 32); head += 32; __syncwarp(); }
    };
    // One loop, one call site of candidate() (code size).  A warp first takes the tiles of multi-tile triangles, one 8x4 tile per step
    // (one lane per pixel), round-robin over all warps; then the single-tile triangles, 32 per chunk, their candidate pixels concatenated.
    const unsigned n_tiles = min(*a.q.tile_count, a.q.tile_cap), n_items = min(*a.q.pixel_count, a.q.pixel_cap);
    // small items per chunk: 32 when there is enough work for every warp, fewer otherwise — a warp works through its chunk's candidates 32 at
    // a time with a dependent fetch chain per step, so spreading a small scene's items over all warps shortens the kernel
    const unsigned per = min(32u, max(1u, (n_items + warps - 1u) / warps));
    unsigned item = gw, base = gw * per;
    uint32_t cslot = 0; int cox = 0, coy = 0, ciw = 1, cexcl = 0, total = 0, c0 = 0;        // the current chunk of small items (one per lane)
    for (;;) {
        bool cv; uint32_t slot; int px, py;
        if (item < n_tiles) {
            const uint2 it = __ldg(a.q.tiles + item);
            item += warps;
            cv = true; slot = it.x; px = (int)(it.y & 0xFFFFu) + (lane & (kTileW - 1)); py = (int)(it.y >> 16) + (lane >> 3);
        } else {
            if (c0 >= total) {                               // next chunk: one item per lane, prefix sum of the candidate counts
                if (base >= n_items) break;
                int cnt = 0; cslot = 0; cox = 0; coy = 0; ciw = 1;
                if ((unsigned)lane < per && base + lane < n_items) {
                    const uint2 it = __ldg(a.q.pixels + base + lane);
                    cslot = it.x; cox = (int)(it.y & 0xFFFFu); coy = (int)(it.y >> 16);
                    const int4 bb = *reinterpret_cast<const int4*>(&a.setups[cslot].s.x0);      // x0, x1, y0, y1
                    ciw = min(kTileW, bb.y - cox + 1);
                    cnt = ciw * min(kTileH, bb.w - coy + 1);
                }
                base += warps * per;
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                total = __shfl_sync(0xffffffffu, inc, 31); cexcl = inc - cnt; c0 = 0;
                if (total == 0) continue;
            }
            const int c = c0 + lane;
            c0 += 32;
            int j = 0;                                       // largest lane whose exclusive prefix is <= c
#pragma unroll
            for (int st = 16; st; st >>= 1) { const int e = __shfl_sync(0xffffffffu, cexcl, j + st); if (e <= c) j += st; }
            slot = __shfl_sync(0xffffffffu, cslot, j);
            const int jox = __shfl_sync(0xffffffffu, cox, j), joy = __shfl_sync(0xffffffffu, coy, j), jw = __shfl_sync(0xffffffffu, ciw, j);
            const int local = c - __shfl_sync(0xffffffffu, cexcl, j);
            const int ly = local / jw;
            cv = c < total; px = jox + (local - ly * jw); py = joy + ly;
        }
        candidate(cv, slot, px, py);
    }
    if (MODE != MODE_OCC && tail > head) shade_batch<MODE>(a, fc, ring, head, tail - head);
    if (MODE != MODE_OCC) {
#pragma unroll
        for (int o = 16; o; o >>= 1) counted += __shfl_xor_sync(0xffffffffu, counted, o);
        if (lane == 0 && counted) atomicAdd(&a.counters->total_fragments, counted);
    }
}

// --------------------------------------------------------------------------------------------- resolve
// One thread per fragment; the thread whose fragment is the head of its voxel's list (no later push displaced it: `displaced` is set
// by the pusher and cleared again here, so it needs no per-frame memset) produces the voxel's final words.  One fragment in the
// voxel — the common case (Sponza: 85 %) — and the words are the ones the fragment precomputed: 16 of the record's 48 bytes are
// read.  Otherwise the list (Sponza: <= 10 entries) is ordered by (triangle, raster rank) in registers and the truncating average of
// voxelize.frag:111-139 is replayed in that order; longer lists take the O(n^2) selection path.
// TRANSFER: on a sparse frame without the temporal filter the head also does transferVoxels.comp:39-62 for its voxel (alpha ->
// voxelSetOpacity, radiance <- (0,0,0,alpha), VoxelizeInfo counters) — every occupied voxel has exactly one head, the masked clear
// has zeroed radiance in every flagged segment, and an unoccupied voxel keeps its zeros — so no separate transfer pass runs.
constexpr int kSortMax = 24;
__device__ __forceinline__ unsigned long long order_key(const Frag& f) { return ((unsigned long long)f.tri << 32 | f.rank) + 1ull; }   // canonical order; 0 = padding
// final words of one voxel -> volumes; TRANSFER: transferVoxels.comp:39-62 for this voxel (temporal filter off)
template <bool TRANSFER>
__device__ __forceinline__ void finish_voxel(uint32_t key, uint32_t cw, uint32_t nw, uint32_t* __restrict__ color, uint32_t* __restrict__ normal, uint32_t* __restrict__ radiance,
                                             float opacity, unsigned& uniq, unsigned& maxfrag) {
    if (TRANSFER) {
        const uint32_t a8 = cw >> 24;
        if (a8) {
            uniq++;
            float aw = (float)a8 / 255.0f;
            maxfrag = max(maxfrag, f2u_trunc(255.0f * aw));
            if (opacity > 0.0f) aw = opacity;
            const uint32_t ab = unorm8(aw) << 24;
            cw = (cw & 0x00FFFFFFu) | ab;
            radiance[key] = ab;
        }
    }
    color[key] = cw; normal[key] = nw;
}
template <bool TRANSFER>
__global__ void __launch_bounds__(kThreads) k_voxel_resolve(const Frag* __restrict__ frags, Counters* __restrict__ counters, unsigned frag_cap,
                                                            uint8_t* __restrict__ displaced, uint32_t* __restrict__ color, uint32_t* __restrict__ normal,
                                                            uint32_t* __restrict__ radiance, float opacity, LongArgs lq) {
    const unsigned n = min(counters->n_frag_slots, frag_cap);
    unsigned uniq = 0, maxfrag = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 r0 = __ldcs(reinterpret_cast<const uint4*>(frags + i));      // key, next, cw1, nw1 (flag and record loads in flight together)
        if (displaced[i]) { displaced[i] = 0; continue; }
        const uint32_t key = r0.x;
        uint32_t cw = r0.z, nw = r0.w;                                          // next == 0: a single fragment
        if (r0.y != 0u) {
            const uint32_t count = normal[key];                                 // fragments of this voxel (counted by k_voxel_tiles)
            if (count > (uint32_t)kSortMax && !lq.inline_long) {                // long list: a warp (or the scan kernels) takes it
                const unsigned pos = atomicAdd(&counters->long_count, 1u);
                if (pos < lq.cap) { LongEntry e; e.key = key; e.head = i; e.count = count; e.pad = 0u; lq.queue[pos] = e; }
                else vct_flag_overflow(counters);
                continue;
            }
            if (count > (uint32_t)kSortMax) {
                // The long-list kernels were not launched (no warp mode, and no long list seen so far in this context): tell the host —
                // it launches them from the next frame on — and resolve this one here by selection: find the r-th largest key top
                // down, then replay upwards from it.  2 r walks of the list: fine for the dense-mesh case this is for (N ~ 50-400, once).
                if (counters->overflow_host) counters->overflow_host[1] = 1u;
                if (count > (uint32_t)kMedMax) {                                // a pile-up without the scan kernels: reported (value 2), not resolved
                    counters->overflow = 1u; if (counters->overflow_host) *counters->overflow_host = 2u;
                    continue;
                }
                const int r = (int)((count - 1u) & 255u) + 1;
                unsigned long long bound = ~0ull;                               // descend r times: bound = r-th largest key
                for (int q = 0; q < r; ++q) {
                    unsigned long long best = 0ull;
                    for (uint32_t j = i + 1u; j; j = frags[j - 1u].next) { const unsigned long long k = order_key(frags[j - 1u]); if (k < bound && k > best) best = k; }
                    bound = best;
                }
                cw = 0u; nw = 0u;
                unsigned long long last = bound - 1ull;                         // ascend from it
                for (int q = 0; q < r; ++q) {
                    unsigned long long best = ~0ull; uint32_t bi = 0u;
                    for (uint32_t j = i + 1u; j; j = frags[j - 1u].next) { const unsigned long long k = order_key(frags[j - 1u]); if (k > last && k < best) { best = k; bi = j - 1u; } }
                    const Frag& f = frags[bi];
                    cw = rgba8_avg_insert(cw, f.cr, f.cg, f.cb); nw = rgba8_avg_insert(nw, f.nr, f.ng, f.nb);
                    last = best;
                }
                finish_voxel<TRANSFER>(key, cw, nw, color, normal, radiance, opacity, uniq, maxfrag);
                continue;
            }
            cw = 0u; nw = 0u;
            unsigned long long ord[kSortMax]; uint32_t idx[kSortMax];
            int cnt = 0;
            for (uint32_t j = i + 1u; j && cnt < kSortMax; j = frags[j - 1u].next) {
                const unsigned long long k = order_key(frags[j - 1u]);
                int p = cnt++;
                while (p > 0 && ord[p - 1] > k) { ord[p] = ord[p - 1]; idx[p] = idx[p - 1]; --p; }
                ord[p] = k; idx[p] = j - 1u;
            }
            for (int q = 0; q < cnt; ++q) {
                const Frag& f = frags[idx[q]];
                cw = rgba8_avg_insert(cw, f.cr, f.cg, f.cb); nw = rgba8_avg_insert(nw, f.nr, f.ng, f.nb);
            }
        }
        finish_voxel<TRANSFER>(key, cw, nw, color, normal, radiance, opacity, uniq, maxfrag);
    }
    if (TRANSFER) {
#pragma unroll
        for (int o = 16; o; o >>= 1) { uniq += __shfl_xor_sync(0xffffffffu, uniq, o); maxfrag = max(maxfrag, __shfl_xor_sync(0xffffffffu, maxfrag, o)); }
        if ((threadIdx.x & 31) == 0) { if (uniq) atomicAdd(&counters->unique_voxels, uniq); if (maxfrag) atomicMax(&counters->max_fragments_per_voxel, maxfrag); }
    }
}
// ascending bitonic sort of n2 (power of two) key/slot pairs in shared memory by `nthreads` cooperating threads (tid in [0, nthreads))
__device__ __forceinline__ void bitonic_sort(unsigned long long* k, uint32_t* v, int n2, int tid, int nthreads, bool whole_block) {
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (whole_block) __syncthreads(); else __syncwarp();
            for (int t = tid; t < (n2 >> 1); t += nthreads) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; const uint32_t x = v[lo]; v[lo] = v[hi]; v[hi] = x; }
            }
        }
    if (whole_block) __syncthreads(); else __syncwarp();
}
// replay of the last r entries of an ascending key/slot array (entries [n2 - r, n2)) -> final words
__device__ __forceinline__ void replay_last(const Frag* __restrict__ frags, const uint32_t* __restrict__ slots, int n2, int r, uint32_t& cw, uint32_t& nw) {
    cw = 0u; nw = 0u;
    for (int q = n2 - r; q < n2; ++q) {
        const Frag& f = frags[slots[q]];
        cw = rgba8_avg_insert(cw, f.cr, f.cg, f.cb); nw = rgba8_avg_insert(nw, f.nr, f.ng, f.nb);
    }
}
// the same in two steps for a warp / a CTA: the r records are fetched by all threads into shared memory (6 floats each), then one thread
// replays them without a global load on its dependent chain
__device__ __forceinline__ void stage_last(const Frag* __restrict__ frags, const uint32_t* __restrict__ slots, int n2, int r, float* __restrict__ stage, int tid, int nthreads) {
    for (int q = tid; q < r; q += nthreads) {
        const Frag& f = frags[slots[n2 - r + q]];
        float* o = stage + 6 * q;
        o[0] = f.cr; o[1] = f.cg; o[2] = f.cb; o[3] = f.nr; o[4] = f.ng; o[5] = f.nb;
    }
}
__device__ __forceinline__ void replay_staged(const float* __restrict__ stage, int r, uint32_t& cw, uint32_t& nw) {
    cw = 0u; nw = 0u;
    for (int q = 0; q < r; ++q) { const float* o = stage + 6 * q; cw = rgba8_avg_insert(cw, o[0], o[1], o[2]); nw = rgba8_avg_insert(nw, o[3], o[4], o[5]); }
}
// One warp per queued voxel with N > kSortMax fragments.  N <= kMedMax: list walk, sort, replay.  Longer lists go to the huge table (no walk:
// the scan kernels below) — the giant ones (N > kWalkMax) always, the others while fewer than kHugeMax / 2 of them have asked; a list that gets
// no slot is walked here in chunks of kMedMax, keeping the r largest keys met so far between chunks (exact; 0.3 us per node, for scenes
// with MANY voxels beyond 1024 fragments: a dense mesh on a coarse grid).
constexpr int kMedWarps = 2;
template <bool TRANSFER>
__global__ void __launch_bounds__(kMedWarps * 32) k_voxel_resolve_medium(const Frag* __restrict__ frags, Counters* __restrict__ counters, uint32_t* __restrict__ color,
                                                                          uint32_t* __restrict__ normal, uint32_t* __restrict__ radiance, float opacity, LongArgs lq) {
    __shared__ unsigned long long s_key[kMedWarps][kMedMax];
    __shared__ uint32_t s_slot[kMedWarps][kMedMax];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned n = min(counters->long_count, lq.cap);
    unsigned uniq = 0, maxfrag = 0;
    for (unsigned e = blockIdx.x * kMedWarps + w; e < n; e += gridDim.x * kMedWarps) {
        const LongEntry le = lq.queue[e];
        if (le.count > (uint32_t)kMedMax) {
            int placed = 0;
            if (lane == 0) {
                const bool giant = le.count > kWalkMax;
                if (giant || atomicAdd(&counters->huge_tickets, 1u) < (unsigned)(kHugeMax / 2)) {
                    const unsigned h = atomicAdd(&counters->huge_count, 1u);
                    if (h < (unsigned)kHugeMax) { lq.huge[h] = le; placed = 1; }
                }
                if (giant && !placed) { vct_flag_overflow(counters); placed = 1; }      // more than kHugeMax pile-ups beyond 65536 fragments: reported
            }
            if (__shfl_sync(0xffffffffu, placed, 0)) continue;
        }
        const int r = (int)((le.count - 1u) & 255u) + 1;                     // only the last r in canonical order matter (voxelize.frag:127-133)
        uint32_t j = le.head + 1u;                                           // (lane 0's copy walks)
        int kept = 0, cnt = 0, n2 = 32;
        for (;;) {
            int m = kept;
            if (lane == 0) for (; j && m < kMedMax; j = frags[j - 1u].next) s_slot[w][m++] = j - 1u;
            cnt = __shfl_sync(0xffffffffu, m, 0);
            const bool more = __shfl_sync(0xffffffffu, (int)(j != 0u), 0) != 0;
            n2 = 32; while (n2 < cnt) n2 <<= 1;
            __syncwarp();
            for (int k = kept + lane; k < n2; k += 32) {                     // entries [0, kept) carry their keys over from the last chunk
                if (k < cnt) s_key[w][k] = order_key(frags[s_slot[w][k]]); else { s_key[w][k] = 0ull; s_slot[w][k] = 0u; }
            }
            bitonic_sort(s_key[w], s_slot[w], n2, lane, 32, false);
            if (!more) break;
            // the list goes on: the r largest keys so far move to the front (through registers: the ranges may overlap), the rest is dropped
            const int keep = min(r, cnt);
            unsigned long long kk[8]; uint32_t ss[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { const int i = lane + 32 * q; if (i < keep) { kk[q] = s_key[w][n2 - keep + i]; ss[q] = s_slot[w][n2 - keep + i]; } }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) { const int i = lane + 32 * q; if (i < keep) { s_key[w][i] = kk[q]; s_slot[w][i] = ss[q]; } }
            __syncwarp();
            kept = keep;
        }
        float* stage = reinterpret_cast<float*>(s_key[w]);                   // the keys are done with: 256 x 24 bytes fit in their 8 KB
        __syncwarp();
        const int rr = min(r, cnt);
        stage_last(frags, s_slot[w], n2, rr, stage, lane, 32);
        __syncwarp();
        if (lane == 0) {
            uint32_t cw, nw;
            replay_staged(stage, rr, cw, nw);
            finish_voxel<TRANSFER>(le.key, cw, nw, color, normal, radiance, opacity, uniq, maxfrag);
        }
        __syncwarp();
    }
    if (TRANSFER && lane == 0) { if (uniq) atomicAdd(&counters->unique_voxels, uniq); if (maxfrag) atomicMax(&counters->max_fragments_per_voxel, maxfrag); }
}
// huge voxels, step 1: every fragment record of a voxel in the huge table -> (order key, slot, table index)
__global__ void __launch_bounds__(kThreads) k_voxel_huge_compact(const Frag* __restrict__ frags, Counters* __restrict__ counters, unsigned frag_cap, LongArgs lq) {
    __shared__ uint32_t s_keys[kHugeMax];
    const unsigned nh = min(counters->huge_count, (unsigned)kHugeMax);
    if (!nh) return;
    if (threadIdx.x < kHugeMax) s_keys[threadIdx.x] = threadIdx.x < nh ? lq.huge[threadIdx.x].key : 0xFFFFFFFFu;
    __syncthreads();
    const unsigned n = min(counters->n_frag_slots, frag_cap);
    const int lane = threadIdx.x & 31;
    const unsigned n_round = (n + 31u) & ~31u;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        int h = -1;
        if (i < n) {
            const uint32_t key = __ldg(&frags[i].key);
#pragma unroll
            for (int k = 0; k < kHugeMax; ++k) if (key == s_keys[k]) h = k;
        }
        const unsigned m = __ballot_sync(0xffffffffu, h >= 0);
        if (!m) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&counters->huge_items, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (h >= 0) {
            const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < lq.item_cap) { HugeItem it; it.k = order_key(frags[i]); it.slot = i; it.h = (uint32_t)h; lq.items[pos] = it; }
            else vct_flag_overflow(counters);
            // coarse histogram over the triangle index; neighbouring records mostly belong to the same triangles: one atomic per bin per warp
            const unsigned bin = min((unsigned)(frags[i].tri >> lq.tri_shift), (unsigned)kHugeBins - 1u);
            const unsigned grp = __match_any_sync(m, (unsigned)h << 16 | bin);
            if (lane == __ffs(grp) - 1) atomicAdd(&lq.aux->hist[h][bin], (unsigned)__popc(grp));
        }
    }
}
// huge voxels, step 2: the r fragments that matter are those of the LAST triangles.  From the coarse histogram one CTA per voxel finds
// the bin that holds the r-th largest key; everything at or above that bin goes to a short list (step 3), a few hundred to a few
// thousand items instead of hundreds of thousands.  The histogram is left zeroed for the next frame.
__global__ void __launch_bounds__(256) k_voxel_huge_pick(Counters* __restrict__ counters, LongArgs lq) {
    __shared__ unsigned s_sum[256];
    const unsigned nh = min(counters->huge_count, (unsigned)kHugeMax), h = blockIdx.x;
    if (h >= nh) return;
    const int r = (int)((lq.huge[h].count - 1u) & 255u) + 1;
    unsigned* hist = lq.aux->hist[h];
    // thread t owns bins [16 t, 16 t + 16); suffix sums over the 256 chunk totals, then inside the chunk
    unsigned mine = 0;
    for (int b = 0; b < 16; ++b) mine += hist[16 * threadIdx.x + b];
    s_sum[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0; int t = 255;
        for (; t > 0; --t) { if (acc + s_sum[t] >= (unsigned)r) break; acc += s_sum[t]; }
        int b = 16 * t + 15;
        for (; b > 16 * t; --b) { if (acc + hist[b] >= (unsigned)r) break; acc += hist[b]; }
        lq.aux->thresh[h] = ((unsigned long long)((unsigned)b << lq.tri_shift) << 32) + 1ull;       // every key of a triangle >= b << shift
        lq.aux->small_n[h] = 0u;
    }
    __syncthreads();
    for (int b = 0; b < 16; ++b) hist[16 * threadIdx.x + b] = 0u;
}
// huge voxels, step 3: items at or above the voxel's threshold -> its short list
__global__ void __launch_bounds__(kThreads) k_voxel_huge_gather(Counters* __restrict__ counters, LongArgs lq) {
    const unsigned nh = min(counters->huge_count, (unsigned)kHugeMax);
    if (!nh) return;
    const unsigned n_items = min(counters->huge_items, lq.item_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
        const HugeItem it = lq.items[i];
        if (it.k >= lq.aux->thresh[it.h]) {
            const unsigned pos = atomicAdd(&lq.aux->small_n[it.h], 1u);
            if (pos < (unsigned)kHugeSmall) lq.small[(size_t)it.h * kHugeSmall + pos] = it;
        }
    }
}
// huge voxels, step 4: one CTA per voxel.  Radix select (12-bit digits, top down) of the r-th largest order key among the voxel's
// items, then the r items at or above it are sorted and replayed.
constexpr int kSelThreads = 1024, kSelBins = 4096;
template <bool TRANSFER>
__global__ void __launch_bounds__(kSelThreads) k_voxel_huge_select(const Frag* __restrict__ frags, Counters* __restrict__ counters, uint32_t* __restrict__ color,
                                                                    uint32_t* __restrict__ normal, uint32_t* __restrict__ radiance, float opacity, LongArgs lq) {
    __shared__ unsigned s_hist[kSelBins];
    __shared__ unsigned long long s_key[256]; __shared__ uint32_t s_slot[256];
    __shared__ float s_stage[256 * 6];
    __shared__ unsigned s_warp[kSelThreads / 32];
    __shared__ unsigned s_pick, s_need, s_n;
    const unsigned nh = min(counters->huge_count, (unsigned)kHugeMax);
    const unsigned h = blockIdx.x;
    if (h >= nh) return;
    const LongEntry le = lq.huge[h];
    // the voxel's short list when it fits (the usual case), else every item of every huge voxel (one huge triangle range: slow, still exact)
    const bool use_small = lq.aux->small_n[h] <= (unsigned)kHugeSmall;
    const HugeItem* __restrict__ items = use_small ? lq.small + (size_t)h * kHugeSmall : lq.items;
    const unsigned n_items = use_small ? lq.aux->small_n[h] : min(counters->huge_items, lq.item_cap);
    const int r = (int)((le.count - 1u) & 255u) + 1;
    unsigned long long prefix = 0ull;                         // the digits of the r-th largest key fixed so far (top down)
    unsigned need = (unsigned)r;                              // its rank (from the top) among the items that share the prefix
    for (int shift = 60; shift >= 0; shift -= 12) {
        for (int b = threadIdx.x; b < kSelBins; b += kSelThreads) s_hist[b] = 0u;
        __syncthreads();
        const unsigned long long hi_mask = shift + 12 >= 64 ? 0ull : ~0ull << (shift + 12);
        for (unsigned i = threadIdx.x; i < n_items; i += kSelThreads) {
            const HugeItem it = items[i];
            if (it.h == h && (it.k & hi_mask) == prefix) atomicAdd(&s_hist[(unsigned)(it.k >> shift) & (kSelBins - 1)], 1u);
        }
        __syncthreads();
        {   // largest digit first: the digit that holds the `need`-th largest.  Thread t owns the 4 bins 4095-4t .. 4092-4t; a CTA-wide
            // inclusive scan of the threads' sums (from the top digit down) finds the one thread whose bins straddle `need`
            static_assert(kSelBins == 4 * kSelThreads, "4 bins per thread");
            const int top = kSelBins - 1 - 4 * (int)threadIdx.x;
            const unsigned h0 = s_hist[top], h1 = s_hist[top - 1], h2 = s_hist[top - 2], h3 = s_hist[top - 3];
            const unsigned mine = h0 + h1 + h2 + h3;
            unsigned inc = mine;
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            if (lane == 31) s_warp[w] = inc;
            if (threadIdx.x == 0) { s_pick = 0u; s_need = 1u; }                  // (unreachable default: the voxel holds at least `need` items)
            __syncthreads();
            unsigned before = 0;
            for (int k = 0; k < w; ++k) before += s_warp[k];
            const unsigned lo = before + inc - mine;                              // items in the digits above this thread's bins
            if (lo < need && need <= lo + mine) {
                unsigned acc = lo; int b = top;
                if (acc + h0 >= need) b = top; else { acc += h0; if (acc + h1 >= need) b = top - 1; else { acc += h1; if (acc + h2 >= need) b = top - 2; else { acc += h2; b = top - 3; } } }
                s_pick = (unsigned)b; s_need = need - acc;
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)s_pick << shift; need = s_need;
        __syncthreads();
    }
    // prefix is now the r-th largest key: gather the r items with key >= prefix (keys are unique)
    if (threadIdx.x == 0) s_n = 0u;
    if (threadIdx.x < 256) { s_key[threadIdx.x] = 0ull; s_slot[threadIdx.x] = 0u; }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n_items; i += kSelThreads) {
        const HugeItem it = items[i];
        if (it.h == h && it.k >= prefix) { const unsigned pos = atomicAdd(&s_n, 1u); if (pos < 256u) { s_key[pos] = it.k; s_slot[pos] = it.slot; } }
    }
    __syncthreads();
    bitonic_sort(s_key, s_slot, 256, threadIdx.x, kSelThreads, true);
    const int rr = min(r, (int)min(s_n, 256u));
    stage_last(frags, s_slot, 256, rr, s_stage, threadIdx.x, kSelThreads);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t cw, nw; unsigned uniq = 0, maxfrag = 0;
        replay_staged(s_stage, rr, cw, nw);
        finish_voxel<TRANSFER>(le.key, cw, nw, color, normal, radiance, opacity, uniq, maxfrag);
        if (TRANSFER) { if (uniq) atomicAdd(&counters->unique_voxels, uniq); if (maxfrag) atomicMax(&counters->max_fragments_per_voxel, maxfrag); }
    }
}

__global__ void k_voxel_reset(Counters* c) { c->overflow = 0; c->long_count = 0; c->huge_count = 0; c->huge_items = 0; c->huge_tickets = 0; c->n_frag_slots = 0; c->tile_queue_count = 0; c->setup_count = 0; c->expand_count = 0; c->pixel_count = 0; }

// ================================================================ tessellation voxeliser (reference default)
// testTesselation.tesc/.tese + the fixed-function tessellator (triangles, equal_spacing, point_mode), Application.cpp:585-665.
// One warp per patch: every lane evaluates the control shader (levels from the triangle's size in voxels; a patch inside one
// voxel gets level 0 and is discarded), then the lanes stride over the distinct vertices of the subdivision — ring 0 with the
// outer levels, concentric rings j >= 1 of the inner level n with corners (1-4j/3n, 2j/3n, 2j/3n) and n-2j segments — and
// run the evaluation shader on each: unlit diffuse texel (base level, NEAREST) stored at the point's voxel.  Barycentrics are
// single fp32 divisions of small integers ("canonical tessellator").
__device__ __forceinline__ float glsl_max(float x, float y) { return x < y ? y : x; }
__device__ __forceinline__ int tess_round(float level) { return !(level > 1.0f) ? 1 : (level >= 64.0f ? 64 : (int)ceilf(level)); }
struct TessPatch { V3 w[3], n[3]; float uv[3][2]; int dt; };
template <int MODE>
__device__ __forceinline__ void tess_eval(const VoxArgs& a, const FrameConst& fc, const TessPatch& P, float u, float v, float ww) {
    const V3 pos = (P.w[0] * u + P.w[1] * v) + P.w[2] * ww;
    const V3 nn = (P.n[0] * u + P.n[1] * v) + P.n[2] * ww;
    const float tu = (u * P.uv[0][0] + v * P.uv[1][0]) + ww * P.uv[2][0], tv = (u * P.uv[0][1] + v * P.uv[1][1]) + ww * P.uv[2][1];
    V4 col = mk4(0.f, 0.f, 0.f, 1.f);
    if (P.dt >= 0) col = sample2d(a.tex[P.dt], tu, tv, 0.0f);
    // tese:134 voxelIndex(position, ..., false): warpVoxels is passed as false and the host never sets this program's warpTexture
    // (Application.cpp:630-640), so only the tessellation warp or the linear mapping can apply
    const V3 vp = fc.p.voxelize_tesselation_warp ? tess_warp_position(pos, fc.p) : voxel_linear_position(pos, fc.p);
    const float Df = (float)a.D;
    int ix, iy, iz;
    if (!to_voxel_index(mk3(Df * vp.x, Df * vp.y, Df * vp.z), a.D, ix, iy, iz)) return;
    if (!owns_z(fc.st, iz)) return;                                              // another rank's layers
    const uint32_t o = (uint32_t)(((size_t)iz * a.D + iy) * a.D + ix);
    a.seg[o >> 3] = 1;
    const V3 N = normalize3(nn);
    const V3 nenc = mk3(N.x * 0.5f + 0.5f, N.y * 0.5f + 0.5f, N.z * 0.5f + 0.5f);
    if (MODE == MODE_MAX) {                                                       // tese:69-74 (the texel's alpha is part of the word)
        atomicMax(a.color + o, pack_unorm(col));
        atomicMax(a.normal + o, pack_unorm(mk4(nenc.x, nenc.y, nenc.z, 1.0f)));
    } else {                                                                      // tese:75-78, the CAS loop as written
        rgba8_avg_atomic(a.color + o, col.x, col.y, col.z);
        rgba8_avg_atomic(a.normal + o, nenc.x, nenc.y, nenc.z);
    }
}
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_voxel_tess(VoxArgs a) {
    const FrameConst& fc = *a.fc;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const float Df = (float)a.D;
    for (uint32_t t = warp; t < a.n_tris; t += n_warps) {
        TessPatch P;
        uint32_t ix[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ix[k] = __ldg(a.indices + 3 * (size_t)t + k);
            const float4 w4 = __ldg(a.wpos + ix[k]), n4 = __ldg(a.wnrm + ix[k]);
            P.w[k] = mk3(w4.x, w4.y, w4.z); P.n[k] = mk3(n4.x, n4.y, n4.z);
            P.uv[k][0] = __ldg(a.verts + 14 * (size_t)ix[k] + 6); P.uv[k][1] = __ldg(a.verts + 14 * (size_t)ix[k] + 7);
        }
        // ---- testTesselation.tesc:32-78
        int vox[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const V3 vp = voxel_linear_position(P.w[k], fc.p);
            vox[k][0] = __float2int_rz(Df * vp.x); vox[k][1] = __float2int_rz(Df * vp.y); vox[k][2] = __float2int_rz(Df * vp.z);
        }
        if (vox[0][0] == vox[1][0] && vox[0][0] == vox[2][0] && vox[0][1] == vox[1][1] && vox[0][1] == vox[2][1] && vox[0][2] == vox[1][2] && vox[0][2] == vox[2][2]) continue;   // levels 0: discarded
        const V3 A = P.w[2] - P.w[1], B = P.w[2] - P.w[0], C = P.w[1] - P.w[0];
        const float lx = length3(A), ly = length3(B), lz = length3(C);
        const float s = ((lx + ly) + lz) * 0.5f;
        const float area = sqrtf(((s * (s - lx)) * (s - ly)) * (s - lz));
        const float max_alt = glsl_max((2.0f * area) / lx, glsl_max((2.0f * area) / ly, (2.0f * area) / lz));
        const V3 vs = mk3((fc.p.voxel_max[0] - fc.p.voxel_min[0]) / Df, (fc.p.voxel_max[1] - fc.p.voxel_min[1]) / Df, (fc.p.voxel_max[2] - fc.p.voxel_min[2]) / Df);
        const float dx = fabsf(length3(normalize3(A) * vs)), dy = fabsf(length3(normalize3(B) * vs)), dz = fabsf(length3(normalize3(C) * vs));
        const float lev_in = glsl_max(1.0f, max_alt / vs.x);
        const float lev_o[3] = {glsl_max(1.0f, lz / dz), glsl_max(1.0f, lx / dx), glsl_max(1.0f, ly / dy)};   // outer[0] = C, [1] = A, [2] = B
        if (!(lev_o[0] > 0.0f) || !(lev_o[1] > 0.0f) || !(lev_o[2] > 0.0f)) continue;
        // ---- fixed function: clamp, round, enumerate the distinct vertices
        int n = tess_round(lev_in);
        const int o0 = tess_round(lev_o[0]), o1 = tess_round(lev_o[1]), o2 = tess_round(lev_o[2]);
        P.dt = a.mats[__ldg(a.trimat + t)].diffuse_tex;
        if (lane < 3) tess_eval<MODE>(a, fc, P, lane == 0 ? 1.0f : 0.0f, lane == 1 ? 1.0f : 0.0f, lane == 2 ? 1.0f : 0.0f);
        if (n == 1 && o0 == 1 && o1 == 1 && o2 == 1) continue;
        if (n == 1) n = 2;
        const int e2 = o2 - 1, e0 = o0 - 1, e1 = o1 - 1;                          // interior points of the edges w = 0, u = 0, v = 0
        for (int q = lane; q < e2 + e0 + e1; q += 32) {
            if (q < e2) { const int i = q + 1; tess_eval<MODE>(a, fc, P, (float)(o2 - i) / (float)o2, (float)i / (float)o2, 0.0f); }
            else if (q < e2 + e0) { const int i = q - e2 + 1; tess_eval<MODE>(a, fc, P, 0.0f, (float)(o0 - i) / (float)o0, (float)i / (float)o0); }
            else { const int i = q - e2 - e0 + 1; tess_eval<MODE>(a, fc, P, (float)i / (float)o1, 0.0f, (float)(o1 - i) / (float)o1); }
        }
        for (int j = 1; n - 2 * j >= 0; ++j) {
            const int m = n - 2 * j;
            if (m == 0) { if (lane == 0) { const float third = 1.0f / 3.0f; tess_eval<MODE>(a, fc, P, third, third, third); } break; }
            const float den = (float)(3 * n * m), small = (float)(2 * j) / (float)(3 * n);
            for (int q = lane; q < 3 * m; q += 32) {
                const int e = q / m, i = q - e * m;
                const float big = (float)((3 * n - 4 * j) * (m - i) + 2 * j * i) / den;
                const float rise = (float)(2 * j * (m - i) + (3 * n - 4 * j) * i) / den;
                if (e == 0) tess_eval<MODE>(a, fc, P, big, rise, small); else if (e == 1) tess_eval<MODE>(a, fc, P, small, big, rise); else tess_eval<MODE>(a, fc, P, rise, small, big);
            }
        }
    }
}

template <int MODE>
int run_mode(vct_ctx* c, const VoxArgs& a, const char* bin_name, const char* tiles_name) {
    const int grid = (int)std::min<size_t>((c->n_tris + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_voxel_bin<MODE><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, bin_name);
    // one resident wave (4 CTAs per SM at 64 registers): the warps share the queues round-robin
    if (a.msaa) k_voxel_tiles<MODE, true><<<VCT_SM_COUNT * 4, kThreads, 0, c->stream>>>(a);
    else k_voxel_tiles<MODE, false><<<VCT_SM_COUNT * 4, kThreads, 0, c->stream>>>(a);
    VCT_LAUNCH_CHECK(c, tiles_name);
    return 0;
}

}  // namespace

size_t vctk_vox_setup_bytes() { return sizeof(VoxSetup); }
size_t vctk_frag_bytes() { return sizeof(Frag); }
size_t vctk_huge_aux_bytes() { return sizeof(HugeAux) + (size_t)kHugeMax * kHugeSmall * sizeof(HugeItem); }

static void normal_matrix_host(const float* m, float n[9]) {
    // mat3(transpose(inverse(M))) = cofactor(M3)/det — same operation order as the oracle
    const float a = m[0], b = m[4], c = m[8], d = m[1], e = m[5], f = m[9], g = m[2], h = m[6], i = m[10];
    const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    const float c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
    const float c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
    const float det = (a * c00 + b * c01) + c * c02;
    n[0] = c00 / det; n[1] = c01 / det; n[2] = c02 / det; n[3] = c10 / det; n[4] = c11 / det; n[5] = c12 / det;
    n[6] = c20 / det; n[7] = c21 / det; n[8] = c22 / det;
}

// per-actor model and normal matrices of this frame, in the layout k_transform_vertices reads (uploaded by upload_frame)
void vctk_fill_models(vct_ctx* c, Mat4* models, float* nmats) {
    for (int a = 0; a < c->n_actors; ++a) { for (int i = 0; i < 16; ++i) models[a].m[i] = (i % 5 == 0) ? 1.0f : 0.0f; for (int i = 0; i < 9; ++i) nmats[9 * a + i] = (i % 4 == 0) ? 1.0f : 0.0f; }
    for (auto& m : c->meshes) { models[m.actor] = m.model; normal_matrix_host(m.model.m, &nmats[9 * (size_t)m.actor]); }
}

int vctk_transform_vertices(vct_ctx* c) {
    if (!c->n_vertices) return 0;
    const int grid = (int)std::min<size_t>((c->n_vertices + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_transform_vertices<<<grid, kThreads, 0, c->stream>>>(c->d_vertices, c->d_vactor, c->d_models, c->d_nmats, c->n_vertices, c->d_wpos, c->d_wnrm, c->d_wT, c->d_wB);
    VCT_LAUNCH_CHECK(c, "k_transform_vertices");
    return 0;
}

int vctk_voxelize(vct_ctx* c, bool occupancy, bool counters_already_reset, bool fuse_transfer, bool* transfer_done) {
    if (transfer_done) *transfer_done = false;
    if (!c->n_tris) return 0;
    const vct_frame_params& p = c->h_fc.p;
    VoxArgs a{};
    a.fc = c->d_fc; a.D = occupancy ? VCT_WARP_DIM : c->D;
    a.V = occupancy || !(p.voxelize_multiplier > 0.0f) ? a.D : (int)(p.voxelize_multiplier * (float)a.D);
    a.indices = c->d_indices; a.trimat = c->d_trimat; a.verts = c->d_vertices; a.n_tris = (uint32_t)c->n_tris;
    a.wpos = c->d_wpos; a.wnrm = c->d_wnrm; a.tex = c->d_tex; a.mats = c->d_mat; a.shadow = c->d_shadow; a.warpmap = c->d_warpmap;
    a.setups = reinterpret_cast<VoxSetup*>(c->d_setup); a.setup_cap = (unsigned)c->setup_cap;
    a.q = vctk_tile_queues(c);
    a.frags = reinterpret_cast<Frag*>(c->d_frags); a.frag_cap = (unsigned)c->frag_cap; a.displaced = c->d_displaced;
    a.color = c->d_color; a.normal = c->d_normal; a.occ = c->d_occ; a.counters = c->d_counters; a.seg = c->d_seg[c->seg_cur];
    // the three ortho views reproduce the linear mapping only for a symmetric cube: cull only then
    const bool cube = p.voxel_min[0] == -p.voxel_max[0] && p.voxel_min[1] == -p.voxel_max[1] && p.voxel_min[2] == -p.voxel_max[2] &&
                      p.voxel_max[0] == p.voxel_max[1] && p.voxel_max[1] == p.voxel_max[2] && p.voxel_max[0] > 0.0f;
    // Settings::conservativeRasterization == MSAA (both voxelisation passes, Application.cpp:244-249, 673-678): sample offsets in 1/256 pixel
    a.msaa = p.conservative_raster == VCT_RASTER_MSAA;
    a.ms = SampleSet{};
    if (a.msaa) {
        static const float standard4x[8] = {0.375f, 0.125f, 0.875f, 0.375f, 0.125f, 0.625f, 0.625f, 0.875f};
        bool given = false;
        for (int i = 0; i < 8; ++i) given = given || p.msaa_samples[i] != 0.0f;
        const float* sp = given ? p.msaa_samples : standard4x;
        a.ms.n = 4; a.ms.x_min = a.ms.y_min = 256; a.ms.x_max = a.ms.y_max = 0;
        for (int i = 0; i < 4; ++i) {
            a.ms.x[i] = std::min(255, std::max(0, (int)lrintf(sp[2 * i] * 256.0f))); a.ms.y[i] = std::min(255, std::max(0, (int)lrintf(sp[2 * i + 1] * 256.0f)));
            a.ms.x_min = std::min(a.ms.x_min, a.ms.x[i]); a.ms.x_max = std::max(a.ms.x_max, a.ms.x[i]);
            a.ms.y_min = std::min(a.ms.y_min, a.ms.y[i]); a.ms.y_max = std::max(a.ms.y_max, a.ms.y[i]);
        }
    }
    // (a multisampled fragment's inputs are extrapolated to the pixel centre: its voxel can lie outside the triangle's padded box, so no slab cull then)
    a.slab_cull = !occupancy && c->cfg.world_size > 1 && !p.warp_voxels && !p.warp_texture && cube && p.axis_override < 0 && !a.msaa;
    if (!counters_already_reset) { k_voxel_reset<<<1, 1, 0, c->stream>>>(c->d_counters); VCT_LAUNCH_CHECK(c, "k_voxel_reset"); }
    if (occupancy) {
        VCT_CHECK(c, cudaMemsetAsync(c->d_occ, 0, sizeof(uint32_t) * VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM, c->stream));
        vct_prof_mark(c, "memset");
        return run_mode<MODE_OCC>(c, a, "k_voxel_bin_occ", "k_voxel_tiles_occ");
    }
    if (p.voxelize_tesselation) {                               // the reference's default voxeliser: no rasterisation at all
        const int grid = (int)std::min<size_t>((c->n_tris * 32 + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
        if (p.voxelize_atomic_max) { k_voxel_tess<MODE_MAX><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_tess_max"); }
        else { k_voxel_tess<MODE_CAS><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_tess_cas"); }
        return 0;
    }
    if (p.voxelize_atomic_max) return run_mode<MODE_MAX>(c, a, "k_voxel_bin_max", "k_voxel_tiles_max");
    if (!p.deterministic) return run_mode<MODE_CAS>(c, a, "k_voxel_bin_cas", "k_voxel_tiles_cas");
    // deterministic running average: per-voxel lists, then ordered sequential replay (fused with transferVoxels when the caller asks)
    if (run_mode<MODE_SORTED>(c, a, "k_voxel_bin", "k_voxel_tiles")) return 1;
    // long per-voxel lists: queue + huge table + scan buffer.  The three extra launches run when a warp mode is on (pile-ups) or once
    // the resolve kernel has met a long list in this context (it tells the host through the mapped flag word and resolves that frame's
    // lists itself); otherwise they are skipped: ~15 us of launches per frame that the common case does not need.
    const bool long_path = p.warp_voxels || p.warp_texture || (c->h_overflow && ((volatile unsigned*)c->h_overflow)[1]);
    LongArgs lq{reinterpret_cast<LongEntry*>(c->d_long_queue), (unsigned)c->long_cap, reinterpret_cast<LongEntry*>(c->d_long_queue) + c->long_cap,
                reinterpret_cast<HugeItem*>(c->d_huge_items), c->d_huge_items ? (unsigned)c->frag_cap : 0u, long_path ? 0 : 1,
                reinterpret_cast<HugeAux*>(c->d_huge_aux), reinterpret_cast<HugeItem*>((char*)c->d_huge_aux + sizeof(HugeAux)), 0};
    while (((size_t)c->n_tris >> lq.tri_shift) > (size_t)kHugeBins) lq.tri_shift++;
    const float op = fuse_transfer ? p.voxel_set_opacity : 0.0f;
    if (fuse_transfer) k_voxel_resolve<true><<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(a.frags, c->d_counters, a.frag_cap, a.displaced, c->d_color, c->d_normal, c->d_radiance, op, lq);
    else k_voxel_resolve<false><<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(a.frags, c->d_counters, a.frag_cap, a.displaced, c->d_color, c->d_normal, c->d_radiance, op, lq);
    VCT_LAUNCH_CHECK(c, "k_voxel_resolve");
    if (!long_path) { if (transfer_done) *transfer_done = fuse_transfer; return 0; }
    if (fuse_transfer) k_voxel_resolve_medium<true><<<VCT_SM_COUNT * 4, kMedWarps * 32, 0, c->stream>>>(a.frags, c->d_counters, c->d_color, c->d_normal, c->d_radiance, op, lq);
    else k_voxel_resolve_medium<false><<<VCT_SM_COUNT * 4, kMedWarps * 32, 0, c->stream>>>(a.frags, c->d_counters, c->d_color, c->d_normal, c->d_radiance, op, lq);
    VCT_LAUNCH_CHECK(c, "k_voxel_resolve_medium");
    {
        k_voxel_huge_compact<<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(a.frags, c->d_counters, a.frag_cap, lq);
        VCT_LAUNCH_CHECK(c, "k_voxel_huge_compact");
        k_voxel_huge_pick<<<kHugeMax, 256, 0, c->stream>>>(c->d_counters, lq);
        VCT_LAUNCH_CHECK(c, "k_voxel_huge_pick");
        k_voxel_huge_gather<<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(c->d_counters, lq);
        VCT_LAUNCH_CHECK(c, "k_voxel_huge_gather");
        if (fuse_transfer) k_voxel_huge_select<true><<<kHugeMax, kSelThreads, 0, c->stream>>>(a.frags, c->d_counters, c->d_color, c->d_normal, c->d_radiance, op, lq);
        else k_voxel_huge_select<false><<<kHugeMax, kSelThreads, 0, c->stream>>>(a.frags, c->d_counters, c->d_color, c->d_normal, c->d_radiance, op, lq);
        VCT_LAUNCH_CHECK(c, "k_voxel_huge_select");
    }
    if (transfer_done) *transfer_done = fuse_transfer;
    return 0;
}
