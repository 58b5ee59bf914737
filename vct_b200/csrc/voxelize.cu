// voxelize.cu — dominant-axis triangle voxelisation with running-average RGBA8 accumulation
// (compiled with -fmad=false: coverage, voxel index and colours are bit-exact vs the oracle).
//
// Replaces the GL draw of voxelize.vert/.geom/.frag (reference src/Application.cpp:666-754) and the 32^3
// occupancy draw (:235-301).  Stages, all on the context stream, no host synchronisation:
//   k_transform_vertices  voxelize.vert:15-23 / phong.vert:37-54 : world position, normalMatrix*normal, T, B
//   k_voxel_count         coverage only: fragments per triangle (canonical order needs the exact count)
//   exclusive scan        per-triangle base index = position of its first fragment in draw order
//   k_voxel_emit          voxelize.geom:26-73 + voxelize.frag:186-280: shade each fragment, write
//                         (voxel key, payload) at its canonical sequence number — or, in the free-running /
//                         atomicMax / occupancy modes, apply the image atomic directly
//   radix sort            stable LSD sort of (voxel key, sequence) pairs, 8 bits per pass
//   k_voxel_apply         per voxel: replay imageAtomicRGBA8Avg (voxelize.frag:111-139) over its fragments in
//                         draw order — the sequential semantics of the CAS loop, order fixed
// Work distribution: one thread per triangle for the (vast majority of) sub-voxel / small triangles; a
// triangle whose bounding box exceeds kCoopArea pixels is rasterised by the whole warp, 32 pixels of a row at
// a time, with ballot-ranked in-order compaction.
#include "raster.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCoopArea = 48;
constexpr unsigned kOobFlag = 0x80000000u;
enum { MODE_SORTED = 0, MODE_CAS = 1, MODE_MAX = 2, MODE_OCC = 3 };

// ------------------------------------------------------------------------------------ vertex transform
__global__ void __launch_bounds__(kThreads) k_transform_vertices(const float* __restrict__ verts, const int32_t* __restrict__ vactor,
                                                                 const Mat4* __restrict__ models, const float* __restrict__ nmats, size_t n,
                                                                 float4* __restrict__ wpos, float4* __restrict__ wnrm, float4* __restrict__ wT, float4* __restrict__ wB) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
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

// ------------------------------------------------------------------------------------------- per-triangle
struct VoxTri {
    RV cv[3]; V3 w[3], n[3]; float uv[3][2];
    int axis, material; float rho2;
    TriSetup s; bool valid;
};
__device__ __forceinline__ V3 ld3(const float4* __restrict__ p, uint32_t i) { const float4 q = __ldg(p + i); return mk3(q.x, q.y, q.z); }

// voxelize.geom:26-73
__device__ __forceinline__ void load_tri(const FrameConst& fc, int D, const uint32_t* __restrict__ indices, const float4* __restrict__ wpos,
                                         const float4* __restrict__ wnrm, uint32_t t, VoxTri& T) {
    const uint32_t i0 = __ldg(indices + 3 * (size_t)t), i1 = __ldg(indices + 3 * (size_t)t + 1), i2 = __ldg(indices + 3 * (size_t)t + 2);
    T.w[0] = ld3(wpos, i0); T.w[1] = ld3(wpos, i1); T.w[2] = ld3(wpos, i2);
    T.n[0] = ld3(wnrm, i0); T.n[1] = ld3(wnrm, i1); T.n[2] = ld3(wnrm, i2);
    const V3 f = normalize3((T.n[0] + T.n[1]) + T.n[2]);
    const float ax = fabsf(f.x), ay = fabsf(f.y), az = fabsf(f.z);
    int axis;
    if (ax > ay && ax > az) axis = 0; else if (ay > ax && ay > az) axis = 1; else axis = 2;
    if (fc.p.axis_override >= 0 && fc.p.axis_override <= 2) axis = fc.p.axis_override;
    T.axis = axis;
    const Mat4& mvp = axis == 0 ? fc.mvp_x : axis == 1 ? fc.mvp_y : fc.mvp_z;
#pragma unroll
    for (int k = 0; k < 3; ++k) { const V4 c = mul44(mvp, mk4(T.w[k].x, T.w[k].y, T.w[k].z, 1.0f)); T.cv[k].x = c.x; T.cv[k].y = c.y; T.cv[k].z = c.z; T.cv[k].w = c.w; }
    T.valid = tri_setup(T.cv, D, D, false, T.s);
}
// voxelize.frag:79-108
__device__ __forceinline__ bool frag_voxel(const FrameConst& fc, const VoxTri& T, const float l[3], int D, const uint16_t* __restrict__ warpmap, bool occupancy,
                                           int& ix, int& iy, int& iz) {
    const V3 ndc = mk3(interp1(l, T.cv[0].x, T.cv[1].x, T.cv[2].x), interp1(l, T.cv[0].y, T.cv[1].y, T.cv[2].y), interp1(l, T.cv[0].z, T.cv[1].z, T.cv[2].z));
    V3 u = mk3((ndc.x + 1.0f) * 0.5f, (ndc.y + 1.0f) * 0.5f, (ndc.z + 1.0f) * 0.5f);
    if (T.axis == 0) u = mk3(1.0f - u.z, u.y, u.x);
    else if (T.axis == 1) u = mk3(u.x, 1.0f - u.z, u.y);
    u.z = 1.0f - u.z;
    if (fc.p.warp_voxels) u = voxel_warp(u, voxel_linear_position(eye_of(fc.p), fc.p));
    else if (fc.p.warp_texture && !occupancy) u = warp_sample(warpmap, u);
    return to_voxel_index(mk3((float)D * u.x, (float)D * u.y, (float)D * u.z), D, ix, iy, iz);
}
// fragment exists (coverage + near/far clip) and belongs to this rank's slab
__device__ __forceinline__ bool frag_test(const FrameConst& fc, const VoxTri& T, int px, int py, int D, const uint16_t* __restrict__ warpmap, bool occupancy,
                                          float l[3], bool& oob, int& ix, int& iy, int& iz) {
    if (!tri_cover(T.s, px, py, l)) return false;
    const float z = interp1(l, T.cv[0].z, T.cv[1].z, T.cv[2].z);
    if (z < -1.0f || z > 1.0f) return false;
    oob = !frag_voxel(fc, T, l, D, warpmap, occupancy, ix, iy, iz);
    if (oob) return fc.z_lo == 0;                                         // counted once, by the rank owning z = 0
    return iz >= fc.z_lo && iz < fc.z_hi;
}

struct Shaded { V3 color, nenc; };
// voxelize.frag:195-228
__device__ __forceinline__ Shaded shade_fragment(const FrameConst& fc, const VoxTri& T, const float l[3], const DevTexture* __restrict__ tex,
                                                 const DevMaterial* __restrict__ mats, const float* __restrict__ shadow) {
    const V3 wp = interp3(l, T.w[0], T.w[1], T.w[2]);
    const V3 nn = interp3(l, T.n[0], T.n[1], T.n[2]);
    const float u = interp1(l, T.uv[0][0], T.uv[1][0], T.uv[2][0]), v = interp1(l, T.uv[0][1], T.uv[1][1], T.uv[2][1]);
    V3 col = mk3(0.f, 0.f, 0.f);
    const int dt = mats[T.material].diffuse_tex;
    if (dt >= 0) { const V4 a = sample2d(tex[dt], u, v, T.rho2); col = mk3(a.x, a.y, a.z); }
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

// ---------------------------------------------------------------------------------------- count / emit
struct VoxArgs {
    const FrameConst* fc; int D;
    const uint32_t* indices; const int32_t* trimat; const float* verts; uint32_t n_tris;
    const float4 *wpos, *wnrm;
    const DevTexture* tex; const DevMaterial* mats; const float* shadow; const uint16_t* warpmap;
    uint32_t* tri_count; const uint32_t* tri_base;
    uint32_t *key, *val; float4 *fcolor, *fnormal; uint32_t frag_cap;
    uint32_t *color, *normal, *occ;
    Counters* counters;
};

__device__ __forceinline__ void load_shading_inputs(const VoxArgs& a, uint32_t t, VoxTri& T) {
    T.material = __ldg(a.trimat + t);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t vi = __ldg(a.indices + 3 * (size_t)t + k);
        T.uv[k][0] = __ldg(a.verts + 14 * (size_t)vi + 6); T.uv[k][1] = __ldg(a.verts + 14 * (size_t)vi + 7);
    }
    const int dt = a.mats[T.material].diffuse_tex;
    T.rho2 = dt >= 0 ? tri_rho2_affine(T.cv, T.uv, a.D, a.D, a.tex[dt]) : 0.0f;
}

template <int MODE>
__device__ __forceinline__ void emit_fragment(const VoxArgs& a, const FrameConst& fc, const VoxTri& T, const float l[3], bool oob, int ix, int iy, int iz, uint32_t seq) {
    const int D = a.D;
    if (MODE == MODE_OCC) { if (!oob) atomicOr(a.occ + ((size_t)iz * D + iy) * D + ix, 1u); return; }
    if (MODE != MODE_SORTED && oob) return;
    const Shaded sh = shade_fragment(fc, T, l, a.tex, a.mats, a.shadow);
    const size_t o = oob ? 0 : ((size_t)iz * D + iy) * D + ix;
    if (MODE == MODE_SORTED) {
        if (seq >= a.frag_cap) { a.counters->overflow = 1u; return; }
        a.key[seq] = (uint32_t)o;
        a.val[seq] = seq | (oob ? kOobFlag : 0u);
        a.fcolor[seq] = make_float4(sh.color.x, sh.color.y, sh.color.z, 0.f);
        a.fnormal[seq] = make_float4(sh.nenc.x, sh.nenc.y, sh.nenc.z, 0.f);
    } else if (MODE == MODE_CAS) {
        rgba8_avg_atomic(a.color + o, sh.color.x, sh.color.y, sh.color.z);
        rgba8_avg_atomic(a.normal + o, sh.nenc.x, sh.nenc.y, sh.nenc.z);
    } else {                                                               // voxelize.frag:271-274
        atomicMax(a.color + o, pack_unorm(mk4(sh.color.x, sh.color.y, sh.color.z, 1.0f)));
        atomicMax(a.normal + o, pack_unorm(mk4(sh.nenc.x, sh.nenc.y, sh.nenc.z, 1.0f)));
    }
}

// COUNT=true: write the number of fragments of each triangle; COUNT=false: emit them.
template <bool COUNT, int MODE>
__global__ void __launch_bounds__(kThreads) k_voxel_raster(VoxArgs a) {
    const FrameConst& fc = *a.fc;
    const int D = a.D, lane = threadIdx.x & 31;
    const bool occupancy = MODE == MODE_OCC;
    unsigned my_total = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (a.n_tris + 31u) & ~31u;                       // keep warps converged for the ballots
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        VoxTri T; T.valid = false;
        bool big = false;
        if (t < a.n_tris) {
            load_tri(fc, D, a.indices, a.wpos, a.wnrm, t, T);
            if (T.valid) big = (T.s.x1 - T.s.x0 + 1) * (T.s.y1 - T.s.y0 + 1) > kCoopArea;
        }
        // ---- serial path: this lane walks its own small bounding box in raster-scan order
        if (T.valid && !big) {
            uint32_t cnt = 0; bool ready = false; uint32_t base = 0;
            for (int py = T.s.y0; py <= T.s.y1; ++py)
                for (int px = T.s.x0; px <= T.s.x1; ++px) {
                    float l[3]; bool oob; int ix, iy, iz;
                    if (!frag_test(fc, T, px, py, D, a.warpmap, occupancy, l, oob, ix, iy, iz)) continue;
                    if (!COUNT) {
                        if (!ready) { if (MODE != MODE_OCC) load_shading_inputs(a, t, T); if (MODE == MODE_SORTED) base = a.tri_base[t]; ready = true; }
                        emit_fragment<MODE>(a, fc, T, l, oob, ix, iy, iz, base + cnt);
                    }
                    cnt++;
                }
            if (COUNT) a.tri_count[t] = cnt; else my_total += cnt;
        } else if (COUNT && t < a.n_tris && !big) a.tri_count[t] = 0;
        // ---- cooperative path: the warp rasterises each big triangle together, 32 pixels of a row per step
        unsigned todo = __ballot_sync(0xffffffffu, big);
        while (todo) {
            const int src = __ffs(todo) - 1; todo &= todo - 1;
            const uint32_t tt = __shfl_sync(0xffffffffu, t, src);
            VoxTri B; load_tri(fc, D, a.indices, a.wpos, a.wnrm, tt, B);     // every lane rebuilds the same setup
            uint32_t base = 0;
            if (!COUNT) { if (MODE != MODE_OCC) load_shading_inputs(a, tt, B); if (MODE == MODE_SORTED) base = a.tri_base[tt]; }
            uint32_t cnt = 0;
            for (int py = B.s.y0; py <= B.s.y1; ++py)
                for (int cx = B.s.x0; cx <= B.s.x1; cx += 32) {
                    const int px = cx + lane;
                    float l[3]; bool oob = false; int ix = 0, iy = 0, iz = 0;
                    const bool hit = px <= B.s.x1 && frag_test(fc, B, px, py, D, a.warpmap, occupancy, l, oob, ix, iy, iz);
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (!COUNT && hit) emit_fragment<MODE>(a, fc, B, l, oob, ix, iy, iz, base + cnt + __popc(m & ((1u << lane) - 1u)));
                    cnt += __popc(m);
                }
            if (lane == src) { if (COUNT) a.tri_count[tt] = cnt; else my_total += cnt; }
        }
    }
    if (!COUNT && MODE != MODE_OCC && MODE != MODE_SORTED) {
#pragma unroll
        for (int o = 16; o; o >>= 1) my_total += __shfl_xor_sync(0xffffffffu, my_total, o);
        if (lane == 0 && my_total) atomicAdd(&a.counters->total_fragments, my_total);
    }
}

// --------------------------------------------------------------------------------------- exclusive scan
constexpr int kScanTile = 2048;             // 256 threads x 8
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < kThreads / 32 ? s_warp[lane] : 0u, xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += n; }
        if (lane < kThreads / 32) s_warp[lane] = xi - x;
        if (lane == 31) s_warp[kThreads / 32] = xi;
    }
    __syncthreads();
    block_total = s_warp[kThreads / 32];
    const uint32_t r = s_warp[w] + inc - v;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(kThreads) k_scan_reduce(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums, size_t n) {
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    const size_t base = (size_t)blockIdx.x * kScanTile + threadIdx.x * 8;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (base + k < n) v += in[base + k];
    uint32_t total; block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kThreads) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ sums, size_t n,
                                                         uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    const size_t base = (size_t)blockIdx.x * kScanTile + threadIdx.x * 8;
    uint32_t x[8], v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = base + k < n ? in[base + k] : 0u; v += x[k]; }
    uint32_t total; uint32_t run = block_exclusive_scan(v, s_warp, total) + (sums ? sums[blockIdx.x] : 0u);
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (base + k < n) out[base + k] = run; run += x[k]; }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kThreads - 1) *total_out = run;
}

// ------------------------------------------------------------------------------------------- radix sort
// Stable LSD radix sort of (key,val) pairs, 8 bits per pass, count read from device memory.  A fixed grid of
// kSortBlocks CTAs each owns a contiguous range of 2048-key tiles, so the per-(digit, block) histogram has a
// fixed shape [256][kSortBlocks] and nothing depends on a host-visible count.
constexpr int kSortBlocks = 592;             // 148 SMs x 4
constexpr int kSortTile = 2048;
__device__ __forceinline__ void sort_range(uint32_t n, uint32_t& t0, uint32_t& t1) {
    const uint32_t ntiles = (n + kSortTile - 1) / kSortTile, per = (ntiles + kSortBlocks - 1) / kSortBlocks;
    t0 = min(ntiles, blockIdx.x * per); t1 = min(ntiles, t0 + per);
}
__global__ void __launch_bounds__(kThreads) k_sort_hist(const uint32_t* __restrict__ keys, const unsigned* __restrict__ n_ptr, int shift, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0; __syncthreads();
    const uint32_t n = *n_ptr; uint32_t t0, t1; sort_range(n, t0, t1);
    for (uint32_t i = t0 * kSortTile + threadIdx.x; i < min(n, t1 * kSortTile); i += kThreads) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[threadIdx.x * kSortBlocks + blockIdx.x] = h[threadIdx.x];
}
// one CTA per digit: exclusive scan of its row, row total -> totals[digit]
__global__ void __launch_bounds__(kThreads) k_sort_rowscan(uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    uint32_t* row = hist + (size_t)blockIdx.x * kSortBlocks;
    uint32_t carry = 0;
    for (int base = 0; base < kSortBlocks; base += kThreads) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < kSortBlocks ? row[i] : 0u;
        uint32_t total; const uint32_t ex = block_exclusive_scan(v, s_warp, total);
        if (i < kSortBlocks) row[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(kThreads) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out, const unsigned* __restrict__ n_ptr, int shift,
                                                           const uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[kThreads / 32 + 1];
    __shared__ uint32_t run[256];                       // next free output slot per digit for this CTA
    __shared__ uint32_t cnt[kThreads / 32][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {   // digit base = exclusive scan of the 256 totals + this CTA's offset inside the digit
        uint32_t total; const uint32_t ex = block_exclusive_scan(totals[threadIdx.x], s_warp, total);
        run[threadIdx.x] = ex + hist[threadIdx.x * kSortBlocks + blockIdx.x];
    }
    __syncthreads();
    const uint32_t n = *n_ptr; uint32_t t0, t1; sort_range(n, t0, t1);
    for (uint32_t tile = t0; tile < t1; ++tile) {
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) cnt[k][threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[8], off[8]; bool ok[8];
        const uint32_t wbase = tile * kSortTile + w * 256;
#pragma unroll
        for (int c = 0; c < 8; ++c) {                   // each warp: 8 chunks of 32 consecutive keys, in order
            const uint32_t i = wbase + c * 32 + lane;
            ok[c] = i < n; key[c] = ok[c] ? keys_in[i] : 0xFFFFFFFFu;
            const uint32_t d = ok[c] ? (key[c] >> shift) & 255u : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (ok[c] && lane == leader) { old = cnt[w][d]; cnt[w][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            off[c] = old + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
        }
        __syncthreads();
        {   // per digit: exclusive prefix over the warps, advancing the CTA's running slot
            uint32_t b = run[threadIdx.x];
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) { const uint32_t v = cnt[k][threadIdx.x]; cnt[k][threadIdx.x] = b; b += v; }
            run[threadIdx.x] = b;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (ok[c]) {
                const uint32_t i = wbase + c * 32 + lane, dst = cnt[w][(key[c] >> shift) & 255u] + off[c];
                keys_out[dst] = key[c]; vals_out[dst] = vals_in[i];
            }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ apply
__global__ void __launch_bounds__(kThreads) k_voxel_apply(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const unsigned* __restrict__ n_ptr,
                                                          const float4* __restrict__ fcolor, const float4* __restrict__ fnormal, uint32_t* __restrict__ color,
                                                          uint32_t* __restrict__ normal) {
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        if (i > 0 && keys[i - 1] == key) continue;                          // not the head of this voxel's run
        uint32_t cw = 0u, nw = 0u; bool any = false;
        for (uint32_t j = i; j < n && keys[j] == key; ++j) {
            const uint32_t v = vals[j];
            if (v & kOobFlag) continue;
            const float4 c = fcolor[v], m = fnormal[v];
            cw = rgba8_avg_insert(cw, c.x, c.y, c.z); nw = rgba8_avg_insert(nw, m.x, m.y, m.z); any = true;
        }
        if (any) { color[key] = cw; normal[key] = nw; }
    }
}

__global__ void k_set_frag_count(Counters* c, unsigned cap) {
    c->total_fragments = c->n_frag_slots;
    if (c->n_frag_slots > cap) { c->n_frag_slots = cap; c->overflow = 1u; }
}

int exclusive_scan(vct_ctx* c, const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, uint32_t* total_out) {
    const int nb = (int)((n + kScanTile - 1) / kScanTile);
    if (nb <= 1) {
        k_scan_apply<<<1, kThreads, 0, c->stream>>>(in, out, nullptr, n, total_out); VCT_LAUNCH_CHECK(c, "k_scan_apply");
        return 0;
    }
    k_scan_reduce<<<nb, kThreads, 0, c->stream>>>(in, tmp, n); VCT_LAUNCH_CHECK(c, "k_scan_reduce");
    if (exclusive_scan(c, tmp, tmp, (size_t)nb, tmp + ((nb + 31) & ~31), nullptr)) return 1;
    k_scan_apply<<<nb, kThreads, 0, c->stream>>>(in, out, tmp, n, total_out); VCT_LAUNCH_CHECK(c, "k_scan_apply");
    return 0;
}

}  // namespace

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

int vctk_transform_vertices(vct_ctx* c) {
    if (!c->n_vertices) return 0;
    std::vector<Mat4>& models = c->h_models; std::vector<float>& nm = c->h_nmats;
    models.resize(c->n_actors); nm.resize((size_t)c->n_actors * 9);
    for (auto& m : c->meshes) { models[m.actor] = m.model; normal_matrix_host(m.model.m, &nm[9 * (size_t)m.actor]); }
    VCT_CHECK(c, cudaMemcpyAsync(c->d_models, models.data(), models.size() * sizeof(Mat4), cudaMemcpyHostToDevice, c->stream));
    VCT_CHECK(c, cudaMemcpyAsync(c->d_nmats, nm.data(), nm.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    vct_prof_mark(c, "h2d_params");
    const int grid = (int)std::min<size_t>((c->n_vertices + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_transform_vertices<<<grid, kThreads, 0, c->stream>>>(c->d_vertices, c->d_vactor, c->d_models, c->d_nmats, c->n_vertices, c->d_wpos, c->d_wnrm, c->d_wT, c->d_wB);
    VCT_LAUNCH_CHECK(c, "k_transform_vertices");
    return 0;
}

int vctk_voxelize(vct_ctx* c, bool occupancy) {
    if (!c->n_tris) return 0;
    const vct_frame_params& p = c->h_fc.p;
    VoxArgs a{};
    a.fc = c->d_fc; a.D = occupancy ? VCT_WARP_DIM : c->D;
    a.indices = c->d_indices; a.trimat = c->d_trimat; a.verts = c->d_vertices; a.n_tris = (uint32_t)c->n_tris;
    a.wpos = c->d_wpos; a.wnrm = c->d_wnrm; a.tex = c->d_tex; a.mats = c->d_mat; a.shadow = c->d_shadow; a.warpmap = c->d_warpmap;
    a.tri_count = c->d_tri_count; a.tri_base = c->d_tri_base;
    a.key = c->d_key[0]; a.val = c->d_val[0]; a.fcolor = c->d_frag_color; a.fnormal = c->d_frag_normal; a.frag_cap = (uint32_t)c->frag_cap;
    a.color = c->d_color; a.normal = c->d_normal; a.occ = c->d_occ; a.counters = c->d_counters;
    const int grid = (int)std::min<size_t>((c->n_tris + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    if (occupancy) {
        VCT_CHECK(c, cudaMemsetAsync(c->d_occ, 0, sizeof(uint32_t) * VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM, c->stream));
        vct_prof_mark(c, "memset");
        k_voxel_raster<false, MODE_OCC><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_raster_occ");
        return 0;
    }
    if (p.voxelize_atomic_max) { k_voxel_raster<false, MODE_MAX><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_raster_max"); return 0; }
    if (!p.deterministic) { k_voxel_raster<false, MODE_CAS><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_raster_cas"); return 0; }
    // deterministic running average: count -> scan -> emit -> sort by voxel -> sequential apply
    k_voxel_raster<true, MODE_SORTED><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_raster_count");
    if (exclusive_scan(c, c->d_tri_count, c->d_tri_base, c->n_tris, c->d_scan_tmp, &c->d_counters->n_frag_slots)) return 1;
    k_set_frag_count<<<1, 1, 0, c->stream>>>(c->d_counters, (unsigned)c->frag_cap); VCT_LAUNCH_CHECK(c, "k_set_frag_count");
    k_voxel_raster<false, MODE_SORTED><<<grid, kThreads, 0, c->stream>>>(a); VCT_LAUNCH_CHECK(c, "k_voxel_raster_emit");
    int bits = 0; while ((1ull << bits) < (unsigned long long)c->D * c->D * c->D) bits++;
    int cur = 0;
    for (int shift = 0; shift < bits; shift += 8) {
        k_sort_hist<<<kSortBlocks, kThreads, 0, c->stream>>>(c->d_key[cur], &c->d_counters->n_frag_slots, shift, c->d_hist); VCT_LAUNCH_CHECK(c, "k_sort_hist");
        k_sort_rowscan<<<256, kThreads, 0, c->stream>>>(c->d_hist, c->d_hist + 256 * kSortBlocks); VCT_LAUNCH_CHECK(c, "k_sort_rowscan");
        k_sort_scatter<<<kSortBlocks, kThreads, 0, c->stream>>>(c->d_key[cur], c->d_val[cur], c->d_key[cur ^ 1], c->d_val[cur ^ 1], &c->d_counters->n_frag_slots, shift,
                                                                c->d_hist, c->d_hist + 256 * kSortBlocks);
        VCT_LAUNCH_CHECK(c, "k_sort_scatter");
        cur ^= 1;
    }
    k_voxel_apply<<<VCT_SM_COUNT * 8, kThreads, 0, c->stream>>>(c->d_key[cur], c->d_val[cur], &c->d_counters->n_frag_slots, c->d_frag_color, c->d_frag_normal, c->d_color, c->d_normal);
    VCT_LAUNCH_CHECK(c, "k_voxel_apply");
    return 0;
}
