// volume_passes.cu — the HBM-bound voxel-space passes (compiled with -fmad=false: bit-exact vs the oracle).
//
//   k_clear            glClearTexImage of voxelColor/voxelNormal           reference src/Application.cpp:686-687
//   k_transfer         transferVoxels.comp:29-70 fused with the radiance clear of Application.cpp:762-764
//   k_inject           injectRadiance.comp:40-103
//   k_fill_holes       voxelFillHoles.comp:8-36 (+ the copy back, Application.cpp:867-872)
//   k_mip_box2(_small) filterRadiance.comp:25-35 (BOX2), k_mip_generic for BOX3/CUBE (:36-58)
//   k_publish          linear level -> level of the cudaMipmappedArray the cone tracer samples
//   dead variants      setVoxelOpacity.comp, temporalRadianceFilter.comp, filter3d.comp, normalizeVoxels.comp
//
// Data layout: every volume level is a linear uint32[d^3], x fastest (RGBA8, R in bits 0-7).  All streaming
// kernels move 16 bytes per thread per access (4 voxels) and are launched with enough CTAs to fill 148 SMs
// several times over; grid-stride where the element count is large.
#include <algorithm>
#include <cuda_fp16.h>
#include <string.h>

#include "raster.cuh"

namespace {

constexpr int kThreads = 256;
static inline int grid_for(size_t items, int threads, int max_waves = 16) {
    size_t g = (items + threads - 1) / threads;
    const size_t cap = (size_t)VCT_SM_COUNT * 8 * max_waves;      // 8 CTAs of 256 threads per SM
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------ clear
// `reset` (whole-frame entry points): the first thread also zeroes the per-frame counters — VoxelizeInfo
// (glClearNamedBufferData, Application.cpp:581), the raster work queues and the cone-step count — which saves two memsets
// and a one-thread kernel per frame.
__global__ void __launch_bounds__(kThreads) k_clear(uint4* __restrict__ a, uint4* __restrict__ b, size_t n16, Counters* __restrict__ reset) {
    if (reset && blockIdx.x == 0 && threadIdx.x == 0) {
        reset->total_fragments = 0; reset->unique_voxels = 0; reset->max_fragments_per_voxel = 0;
        reset->n_frag_slots = 0; reset->tile_queue_count = 0; reset->setup_count = 0; reset->expand_count = 0; reset->pixel_count = 0;
        reset->cone_steps = 0ull; reset->overflow = 0; reset->long_count = 0; reset->huge_count = 0; reset->huge_items = 0; reset->huge_tickets = 0;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        a[i] = z;
        if (b) b[i] = z;
    }
}

// Sparse frame: one thread = 4 consecutive segments (32 voxels, 128 bytes per volume).  Segments flagged in last frame's
// mask are zeroed in voxelColor, voxelNormal and (unless the temporal filter keeps it) voxelRadiance; this frame's mask
// starts empty (temporal: inherits, because the decaying radiance keeps its support).  Also resets the frame counters.
// (w_stripe mask words per owned stripe, n_words in all of them: the loop index runs over this rank's stripes back to back)
__device__ __forceinline__ void clear_masked_part(uint4* __restrict__ color, uint4* __restrict__ normal, uint4* __restrict__ radiance,
                                                  const uint32_t* __restrict__ seg_prev, uint32_t* __restrict__ seg_cur, Stripes st, size_t w_stripe, size_t n_words, int temporal,
                                                  Counters* __restrict__ reset, unsigned block, unsigned n_blocks) {
    if (block == 0 && threadIdx.x == 0) {
        reset->total_fragments = 0; reset->unique_voxels = 0; reset->max_fragments_per_voxel = 0;
        reset->n_frag_slots = 0; reset->tile_queue_count = 0; reset->setup_count = 0; reset->expand_count = 0; reset->pixel_count = 0;
        reset->cone_steps = 0ull; reset->overflow = 0; reset->long_count = 0; reset->huge_count = 0; reset->huge_items = 0; reset->huge_tickets = 0;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t tl = block * (size_t)blockDim.x + threadIdx.x; tl < n_words; tl += (size_t)n_blocks * blockDim.x) {
        const size_t t = stripe_index(st, tl, w_stripe);
        const uint32_t flags = __ldg(seg_prev + t);
        seg_cur[t] = temporal ? flags : 0u;
        if (!flags) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!(flags >> (8 * (j >> 1)) & 0xFFu)) continue;
            __stcs(color + 8 * t + j, z); __stcs(normal + 8 * t + j, z);
            if (!temporal) __stcs(radiance + 8 * t + j, z);
        }
    }
}
__global__ void __launch_bounds__(kThreads) k_clear_masked(uint4* __restrict__ color, uint4* __restrict__ normal, uint4* __restrict__ radiance,
                                                           const uint32_t* __restrict__ seg_prev, uint32_t* __restrict__ seg_cur, Stripes st, size_t w_stripe, size_t n_words, int temporal,
                                                           Counters* __restrict__ reset) {
    clear_masked_part(color, normal, radiance, seg_prev, seg_cur, st, w_stripe, n_words, temporal, reset, blockIdx.x, gridDim.x);
}
// Frame begin of vct_gi_passes on a sparse frame: the vertex transform and the masked clear are independent, both are
// short and neither fills the GPU, so they share one launch (the first `t_blocks` CTAs transform, the rest clear).
struct TransformArgs { const float* verts; const int32_t* vactor; const Mat4* models; const float* nmats; size_t n; float4 *wpos, *wnrm, *wT, *wB; };
__global__ void __launch_bounds__(kThreads) k_frame_begin(TransformArgs t, unsigned t_blocks, uint4* __restrict__ color, uint4* __restrict__ normal, uint4* __restrict__ radiance,
                                                          const uint32_t* __restrict__ seg_prev, uint32_t* __restrict__ seg_cur, Stripes st, size_t w_stripe, size_t n_words, int temporal,
                                                          Counters* __restrict__ reset) {
    if (blockIdx.x < t_blocks) transform_vertices_part(t.verts, t.vactor, t.models, t.nmats, t.n, t.wpos, t.wnrm, t.wT, t.wB, blockIdx.x, t_blocks);
    else clear_masked_part(color, normal, radiance, seg_prev, seg_cur, st, w_stripe, n_words, temporal, reset, blockIdx.x - t_blocks, gridDim.x - t_blocks);
}

// --------------------------------------------------------------------------------------------- transfer
// transferVoxels.comp:39-62 on four voxels (one 16-byte word group).  Returns true when voxelColor must be written back.
__device__ __forceinline__ bool transfer_quad(uint4& cw, const uint4 pw, uint4& rw, float opacity, int temporal, float decay, unsigned& uniq, unsigned& maxfrag) {
    uint32_t c[4] = {cw.x, cw.y, cw.z, cw.w};
    uint32_t r[4] = {0, 0, 0, 0};
    const uint32_t prev[4] = {pw.x, pw.y, pw.z, pw.w};
    bool dirty = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // imageLoad -> float -> imageStore of an unchanged channel returns the same byte
        // (round(b/255*255) == b), so only alpha needs arithmetic: transferVoxels.comp:39-50
        const uint32_t a8 = c[k] >> 24;
        float aw = 0.0f;
        if (a8) {
            uniq++;
            aw = (float)a8 / 255.0f;
            maxfrag = max(maxfrag, f2u_trunc(255.0f * aw));
            if (opacity > 0.0f) aw = opacity;
            c[k] = (c[k] & 0x00FFFFFFu) | unorm8(aw) << 24;
            dirty = true;
        }
        if (temporal) {                                             // mix(prev, vec4(0,0,0,aw), 1 - decay), :55-62
            const V4 p = unpack_unorm(prev[k]);
            const float a = 1.0f - decay;
            r[k] = pack_unorm(mk4(mixf(p.x, 0.0f, a), mixf(p.y, 0.0f, a), mixf(p.z, 0.0f, a), mixf(p.w, aw, a)));
        } else if (aw > 0.0f) r[k] = unorm8(aw) << 24;
    }
    cw = make_uint4(c[0], c[1], c[2], c[3]); rw = make_uint4(r[0], r[1], r[2], r[3]);
    return dirty;
}
// block reduction of the VoxelizeInfo counters -> one atomic pair per CTA
__device__ __forceinline__ void transfer_counters(unsigned uniq, unsigned maxfrag, Counters* __restrict__ counters) {
    __shared__ unsigned s_u[kThreads / 32], s_m[kThreads / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) { uniq += __shfl_xor_sync(0xffffffffu, uniq, o); maxfrag = max(maxfrag, __shfl_xor_sync(0xffffffffu, maxfrag, o)); }
    if ((threadIdx.x & 31) == 0) { s_u[threadIdx.x >> 5] = uniq; s_m[threadIdx.x >> 5] = maxfrag; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned u = 0, m = 0;
        for (int w = 0; w < kThreads / 32; ++w) { u += s_u[w]; m = max(m, s_m[w]); }
        if (u) atomicAdd(&counters->unique_voxels, u);
        if (m) atomicMax(&counters->max_fragments_per_voxel, m);
    }
}
// One thread = 4 voxels (one 16-byte load of voxelColor).  Writes voxelColor back only where a fragment
// landed; writes EVERY radiance voxel (0 where empty), which is the reference's clear + conditional store.
// `seg_mark` (temporal filter, dense frame): segments whose decayed radiance is still non-zero are flagged in this frame's
// mask, so that the mask bounds the radiance support and the next frame can be sparse.
__global__ void __launch_bounds__(kThreads) k_transfer(uint4* __restrict__ color, uint4* __restrict__ radiance, size_t n16,
                                                       float opacity, int temporal, float decay, Counters* __restrict__ counters, uint8_t* __restrict__ seg_mark) {
    unsigned uniq = 0, maxfrag = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n16; i0 += 2 * stride) {
        // two independent 16-byte items per thread per trip: both loads are in flight before any arithmetic
        const size_t i1 = i0 + stride; const bool has1 = i1 < n16;
        uint4 cws[2], pws[2];
        cws[0] = color[i0]; cws[1] = has1 ? color[i1] : make_uint4(0, 0, 0, 0);
        pws[0] = pws[1] = make_uint4(0, 0, 0, 0);
        if (temporal) { pws[0] = radiance[i0]; if (has1) pws[1] = radiance[i1]; }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !has1) break;
            const size_t i = u ? i1 : i0;
            uint4 cw = cws[u]; const uint4 pw = pws[u];
            if ((cw.x | cw.y | cw.z | cw.w) == 0u) {                       // ~97 % of the grid: four empty voxels
                if (!temporal) { radiance[i] = make_uint4(0, 0, 0, 0); continue; }      // = the reference's clear
                if ((pw.x | pw.y | pw.z | pw.w) == 0u) continue;                        // mix(0, 0, a) = 0: already stored
            }
            uint4 rw;
            if (transfer_quad(cw, pw, rw, opacity, temporal, decay, uniq, maxfrag)) color[i] = cw;
            radiance[i] = rw;
            if (seg_mark && (rw.x | rw.y | rw.z | rw.w) != 0u) seg_mark[i >> 1] = 1;
        }
    }
    transfer_counters(uniq, maxfrag, counters);
}
// Sparse frame.  A warp scans 32 mask words (128 segments), compacts the flagged segments into a shared list, and then
// works through their 16-byte quads with all lanes (about one segment in ten is flagged: without the compaction most
// lanes would idle behind the few that found work).  Only flagged segments can hold a fragment; their radiance was
// zeroed by k_clear_masked (non-temporal) or holds last frame's value (temporal).
__global__ void __launch_bounds__(kThreads) k_transfer_masked(uint4* __restrict__ color, uint4* __restrict__ radiance, const uint32_t* __restrict__ seg, Stripes st, size_t w_stripe, size_t n_words,
                                                              float opacity, int temporal, float decay, Counters* __restrict__ counters) {
    __shared__ uint32_t s_list[kThreads / 32][128];
    unsigned uniq = 0, maxfrag = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t n_round = (n_words + 31) & ~(size_t)31;                                     // n_words: over this rank's stripes, back to back
    for (size_t tl = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tl < n_round; tl += (size_t)gridDim.x * blockDim.x) {
        const size_t t = tl < n_words ? stripe_index(st, tl, w_stripe) : 0;
        const uint32_t flags = tl < n_words ? __ldg(seg + t) : 0u;
        const unsigned nz = (flags & 0xFFu ? 1u : 0u) | (flags & 0xFF00u ? 2u : 0u) | (flags & 0xFF0000u ? 4u : 0u) | (flags & 0xFF000000u ? 8u : 0u);
        int inc = __popc(nz);
        const int mine = inc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        if (!total) continue;
        int pos = inc - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (nz >> j & 1u) s_list[w][pos++] = (uint32_t)(4 * t + j);
        __syncwarp();
        for (int q = lane; q < 2 * total; q += 32) {
            const size_t i = 2 * (size_t)s_list[w][q >> 1] + (q & 1);
            uint4 cw = __ldcs(color + i);
            const uint4 pw = temporal ? __ldcs(radiance + i) : make_uint4(0, 0, 0, 0);
            if ((cw.x | cw.y | cw.z | cw.w) == 0u && (!temporal || (pw.x | pw.y | pw.z | pw.w) == 0u)) continue;
            uint4 rw;
            if (transfer_quad(cw, pw, rw, opacity, temporal, decay, uniq, maxfrag)) color[i] = cw;
            radiance[i] = rw;
        }
        __syncwarp();
    }
    transfer_counters(uniq, maxfrag, counters);
}

// setVoxelOpacity.comp:18-35 (dead variant)
__global__ void __launch_bounds__(kThreads) k_set_voxel_opacity(uint32_t* __restrict__ color, uint32_t* __restrict__ radiance, size_t n,
                                                                float opacity, Counters* __restrict__ counters) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        V4 v = unpack_unorm(color[i]);
        if (v.w > 0.0f) {
            atomicAdd(&counters->unique_voxels, 1u);
            atomicMax(&counters->max_fragments_per_voxel, f2u_trunc(255.0f * v.w));
            if (opacity > 0.0f) v.w = opacity;
            color[i] = pack_unorm(v);
            radiance[i] = pack_unorm(mk4(0.f, 0.f, 0.f, v.w));
        }
    }
}
// temporalRadianceFilter.comp:9-19 (dead variant)
__global__ void __launch_bounds__(kThreads) k_temporal_decay(uint32_t* __restrict__ vol, size_t n, float decay) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const V4 v = unpack_unorm(vol[i]);
        vol[i] = pack_unorm(mk4(v.x * decay, v.y * decay, v.z * decay, v.w * decay));
    }
}
// normalizeVoxels.comp:19-42 (dead variant; RGBA16F volumes)
__global__ void __launch_bounds__(kThreads) k_normalize_f16(__half* __restrict__ col, __half* __restrict__ nrm, uint32_t* __restrict__ radiance,
                                                            size_t n, float opacity, Counters* __restrict__ counters) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = __half2float(col[4 * i + k]);
        if (c[3] > 0.0f) {
            atomicAdd(&counters->unique_voxels, 1u);
            atomicMax(&counters->max_fragments_per_voxel, f2u_trunc(c[3]));
            const float a = c[3];
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = c[k] / a;
            if (opacity > 0.0f) c[3] = opacity;
#pragma unroll
            for (int k = 0; k < 4; ++k) col[4 * i + k] = __float2half_rn(c[k]);
            radiance[i] = pack_unorm(mk4(0.f, 0.f, 0.f, c[3]));
        }
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __half2float(nrm[4 * i + k]);
        if (m[3] > 0.0f) {
            const float a = m[3];
#pragma unroll
            for (int k = 0; k < 4; ++k) nrm[4 * i + k] = __float2half_rn(m[k] / a);
        }
    }
}

// ----------------------------------------------------------------------------------------------- inject
// One thread per shadow-map texel, 32x8 tiles so that a warp reads one 128-byte row segment of depth.
// The bilinear fetch at a texel CORNER averages the 2x2 neighbourhood (x-1..x, y-1..y): staged through a
// (32+1)x(8+1) shared tile so each depth value is read from HBM once.
__global__ void __launch_bounds__(256) k_inject_generic(const FrameConst* __restrict__ fcp, const float* __restrict__ shadow,
                                                const uint32_t* __restrict__ color, const uint32_t* __restrict__ normal,
                                                const uint16_t* __restrict__ warpmap, uint32_t* __restrict__ radiance) {
    const FrameConst& fc = *fcp;
    const int S = fc.S, D = fc.D;
    __shared__ float tile[9][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 9 * 33; i += 256) {
        const int lx = i % 33, ly = i / 33;
        tile[ly][lx] = shadow_texel(shadow, S, bx + lx - 1, by + ly - 1);
    }
    __syncthreads();
    const int x = bx + tx, y = by + ty;
    bool valid = x < S && y < S;
    uint32_t o = 0xFFFFFFFFu;
    int ix = 0, iy = 0, iz = 0;
    if (valid) {
        // x/S: S is a power of two in every configuration, where the division is an exact scaling (bit-identical)
        const bool pow2 = (S & (S - 1)) == 0;
        const float inv = 1.0f / (float)S;
        const float tu = pow2 ? (float)x * inv : (float)x / (float)S, tv = pow2 ? (float)y * inv : (float)y / (float)S;
        float d;
        {   // shadow_linear(tu, tv) with texels taken from the tile when the footprint is the expected one
            const float fxp = tu * (float)S - 0.5f, fyp = tv * (float)S - 0.5f;
            const float fx0 = floorf(fxp), fy0 = floorf(fyp);
            const int x0 = (int)fx0, y0 = (int)fy0; const float fx = fxp - fx0, fy = fyp - fy0;
            const int lx = x0 - bx + 1, ly = y0 - by + 1;
            if (lx >= 0 && lx + 1 < 33 && ly >= 0 && ly + 1 < 9) {
                const float top = tile[ly][lx] * (1.0f - fx) + tile[ly][lx + 1] * fx;
                const float bot = tile[ly + 1][lx] * (1.0f - fx) + tile[ly + 1][lx + 1] * fx;
                d = top * (1.0f - fy) + bot * fy;
            } else d = shadow_linear(shadow, S, tu, tv, 0, 0);
        }
        const float nx = tu * 2.0f - 1.0f, ny = tv * 2.0f - 1.0f, nz = d * 2.0f - 1.0f;
        const V4 w = mul44(fc.ls_inverse, mk4(nx, ny, nz, 1.0f));
        V3 vp = get_voxel_position(mk3(w.x, w.y, w.z), fc.p, warpmap);
        vp = mk3((float)D * vp.x, (float)D * vp.y, (float)D * vp.z);
        valid = to_voxel_index(vp, D, ix, iy, iz) && owns_z(fc.st, iz);                  // z ownership (multi-GPU)
        if (valid) o = (uint32_t)(((size_t)iz * D + iy) * D + ix);
    }
    // Neighbouring texels of a row mostly land in the same voxel (4096^2 texels onto ~4e5 voxels) and every writer
    // stores the same word, so a lane whose left neighbour targets the same voxel leaves the store to it.
    const uint32_t left = __shfl_up_sync(0xffffffffu, o, 1);
    if (!valid || (tx > 0 && left == o)) return;
    const uint32_t cw = __ldg(color + o);
    if (!fc.p.radiance_lighting) { radiance[o] = cw; return; }   // packUnorm4x8(unpackUnorm4x8(c)) == c for every byte
    V4 c = unpack_unorm(cw);
    const V4 n4 = unpack_unorm(__ldg(normal + o));
    const V3 n = mk3(2.0f * n4.x - 1.0f, 2.0f * n4.y - 1.0f, 2.0f * n4.z - 1.0f);
    const vct_light& L0 = fc.lights[0];
    V3 lpv = get_voxel_position(mk3(L0.position[0], L0.position[1], L0.position[2]), fc.p, warpmap);
    lpv = mk3((float)D * lpv.x, (float)D * lpv.y, (float)D * lpv.z);
    const V3 lv = normalize3(lpv - mk3((float)ix, (float)iy, (float)iz));
    const float diff = maxsel(dot3(n, lv), 0.0f);
    c.x = (diff * L0.color[0]) * c.x; c.y = (diff * L0.color[1]) * c.y; c.z = (diff * L0.color[2]) * c.z;
    radiance[o] = pack_unorm(c);
}

// Fast path for power-of-two shadow maps (every configuration): one thread = 4 consecutive texels of a row.
// With S = 2^k the texel-corner coordinate x/S is exact, so the LINEAR footprint is exactly texels (x-1..x, y-1..y)
// with weights 0.5 — the same expressions as shadow_linear() are evaluated, only the addressing is simplified: two
// 16-byte loads + two scalars feed four texels, and nothing is staged through shared memory.
__device__ __forceinline__ uint32_t inject_target(const FrameConst& fc, const uint16_t* __restrict__ warpmap, int x, int y, float t00, float t10, float t01, float t11,
                                                  float inv_s, int& ix, int& iy, int& iz) {
    const int D = fc.D;
    const float tu = (float)x * inv_s, tv = (float)y * inv_s;
    const float fx = 0.5f, fy = 0.5f;
    const float top = t00 * (1.0f - fx) + t10 * fx, bot = t01 * (1.0f - fx) + t11 * fx;
    const float d = top * (1.0f - fy) + bot * fy;
    const float nx = tu * 2.0f - 1.0f, ny = tv * 2.0f - 1.0f, nz = d * 2.0f - 1.0f;
    const V4 w = mul44(fc.ls_inverse, mk4(nx, ny, nz, 1.0f));
    V3 vp = get_voxel_position(mk3(w.x, w.y, w.z), fc.p, warpmap);
    vp = mk3((float)D * vp.x, (float)D * vp.y, (float)D * vp.z);
    if (!to_voxel_index(vp, D, ix, iy, iz) || !owns_z(fc.st, iz)) return 0xFFFFFFFFu;              // z ownership (multi-GPU)
    return (uint32_t)(((size_t)iz * D + iy) * D + ix);
}
__device__ __forceinline__ void inject_store(const FrameConst& fc, const uint16_t* __restrict__ warpmap, const uint32_t* __restrict__ color,
                                             const uint32_t* __restrict__ normal, uint32_t* __restrict__ radiance, uint32_t o, int ix, int iy, int iz) {
    const int D = fc.D;
    const uint32_t cw = __ldg(color + o);
    if (!fc.p.radiance_lighting) { radiance[o] = cw; return; }   // packUnorm4x8(unpackUnorm4x8(c)) == c for every byte
    V4 c = unpack_unorm(cw);
    const V4 n4 = unpack_unorm(__ldg(normal + o));
    const V3 n = mk3(2.0f * n4.x - 1.0f, 2.0f * n4.y - 1.0f, 2.0f * n4.z - 1.0f);
    const vct_light& L0 = fc.lights[0];
    V3 lpv = get_voxel_position(mk3(L0.position[0], L0.position[1], L0.position[2]), fc.p, warpmap);
    lpv = mk3((float)D * lpv.x, (float)D * lpv.y, (float)D * lpv.z);
    const V3 lv = normalize3(lpv - mk3((float)ix, (float)iy, (float)iz));
    const float diff = maxsel(dot3(n, lv), 0.0f);
    c.x = (diff * L0.color[0]) * c.x; c.y = (diff * L0.color[1]) * c.y; c.z = (diff * L0.color[2]) * c.z;
    radiance[o] = pack_unorm(c);
}
__global__ void __launch_bounds__(256) k_inject(const FrameConst* __restrict__ fcp, const float* __restrict__ shadow,
                                                const uint32_t* __restrict__ color, const uint32_t* __restrict__ normal,
                                                const uint16_t* __restrict__ warpmap, uint32_t* __restrict__ radiance) {
    const FrameConst& fc = *fcp;
    const int S = fc.S, qx = S >> 2;
    const float inv_s = 1.0f / (float)S;                                    // exact: S is a power of two
    const int q = blockIdx.x * 256 + threadIdx.x;                           // grid covers exactly S*S/4 quads (S >= 32)
    const int x0 = (q % qx) * 4, y = q / qx;
    const float* row1 = shadow + (size_t)y * S + x0;
    const float4 b = __ldg(reinterpret_cast<const float4*>(row1));
    const float bm1 = x0 > 0 ? __ldg(row1 - 1) : 1.0f;                      // CLAMP_TO_BORDER, border 1
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f); float am1 = 1.0f;
    if (y > 0) { a = __ldg(reinterpret_cast<const float4*>(row1 - S)); if (x0 > 0) am1 = __ldg(row1 - S - 1); }
    const float ta[5] = {am1, a.x, a.y, a.z, a.w}, tb[5] = {bm1, b.x, b.y, b.z, b.w};
    uint32_t o[4]; int ix[4], iy[4], iz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = inject_target(fc, warpmap, x0 + k, y, ta[k], ta[k + 1], tb[k], tb[k + 1], inv_s, ix[k], iy[k], iz[k]);
    // Neighbouring texels of a row mostly land in the same voxel (4096^2 texels onto ~4e5 voxels) and every writer
    // stores the same word, so a texel whose left neighbour targets the same voxel leaves the store to it.
    uint32_t left = __shfl_up_sync(0xffffffffu, o[3], 1);
    if ((threadIdx.x & 31) == 0) left = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (o[k] != 0xFFFFFFFFu && o[k] != left) inject_store(fc, warpmap, color, normal, radiance, o[k], ix[k], iy[k], iz[k]);
        left = o[k];
    }
}

// Linear voxel mapping (no warp mode), the case of every benchmark configuration.  Same IEEE results as k_inject with
// about half the instructions: the three divisions by (max - min) of voxelLinearPosition (common.glsl:6-9) use the
// host-rounded reciprocal and two FMA residual corrections — q0 = a*rc; q1 = q0 + (a - q0*c)*rc; q = q1 + (a - q1*c)*rc —
// which is the correctly rounded quotient (Markstein; checked against a/c for every float mantissa), and
// the warp-mode branches are gone.  Results that differ from a/c only for non-finite a are rejected by the bounds test
// either way.
struct InjectLinear {                                 // passed by value: no dependent loads of the frame constants
    float rc[3], c[3], sub0[3], sub1[3];              // per axis: 1/(max-min), max-min, center, min
    float m[16];                                      // ls_inverse, column-major
    int S, log2_qx, D; Stripes st;
    int skip_far;                                     // the light's far plane (depth 1 = nothing rendered) misses the volume by > 1 voxel
    // Block cull: voxel coordinates are affine in (ndc x, ndc y, ndc depth): P_k = bx[k]*nx + by[k]*ny + bz[k]*nz + b0[k].  With the
    // min / max filtered depth of a 64x16 texel block (k_shadow_minmax*, written by the shadow pass) interval arithmetic bounds P over
    // the block; k_inject_cull lists the blocks that can reach the volume — and this rank's z layers —, k_inject_linear runs one CTA per
    // listed block.  The list is a function of the shadow map and of (ls_inverse, volume, grid, slab): it is rebuilt when one of them
    // changes (a new shadow map, a moved volume), not every frame.  Sponza: 2/3 of the shadow map sees no geometry or lies outside the
    // volume and is never read.  (Measured and dropped: the test per thread on 4x4 blocks, 40 -> 49 us, and per CTA inside the inject
    // kernel, 55 us — a dependent load in front of the depth loads either way.)
    const float2* minmax;                             // coarse level (one entry per 64x16 texels); nullptr: every block is listed
    const uint32_t* list;                             // [0] = number of active blocks, [1..] = their ids (k_inject_cull)
    float bx[3], by[3], bz[3], b0[3];
};
constexpr int kInjBlockW = 64, kInjBlockH = 16;      // texels per CTA of k_inject_linear: 16 threads x 4 texels wide, 16 rows
constexpr float kInjectMargin = 0.02f;               // voxels; the float evaluation of the bound is good to ~1e-4 voxel

// min / max over each 4x4 block of shadow texels of the depth injectRadiance.comp sees at the texel CORNER (LINEAR filter: the mean
// of the 2x2 texels around it, border 1) — the very expression k_inject_linear evaluates, so d is inside [min, max] exactly
__global__ void __launch_bounds__(256) k_shadow_minmax(const float* __restrict__ shadow, int S, float2* __restrict__ out) {
    const int nb = S >> 2, b = blockIdx.x * 256 + threadIdx.x;
    if (b >= nb * nb) return;
    const int bx = b % nb, by = b / nb;
    float t[5][5];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) t[j][i] = shadow_texel(shadow, S, 4 * bx - 1 + i, 4 * by - 1 + j);
    float lo = 2.0f, hi = -1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float top = t[j][i] * 0.5f + t[j][i + 1] * 0.5f, bot = t[j + 1][i] * 0.5f + t[j + 1][i + 1] * 0.5f;
            const float d = top * 0.5f + bot * 0.5f;
            lo = fminf(lo, d); hi = fmaxf(hi, d);
        }
    out[b] = make_float2(lo, hi);
}
// coarse level: min / max over the 16 x 4 fine entries of a 64x16 texel block
__global__ void __launch_bounds__(256) k_shadow_minmax_coarse(const float2* __restrict__ fine, int S, float2* __restrict__ out) {
    const int nbx = S / kInjBlockW, nby = S / kInjBlockH, b = blockIdx.x * 256 + threadIdx.x;
    if (b >= nbx * nby) return;
    const int bx = b % nbx, by = b / nbx, nf = S >> 2;
    float lo = 2.0f, hi = -1.0f;
    for (int j = 0; j < kInjBlockH / 4; ++j)
        for (int i = 0; i < kInjBlockW / 4; ++i) { const float2 v = __ldg(fine + (size_t)(by * (kInjBlockH / 4) + j) * nf + bx * (kInjBlockW / 4) + i); lo = fminf(lo, v.x); hi = fmaxf(hi, v.y); }
    out[b] = make_float2(lo, hi);
}
// can the 64x16 texel block `b` of the shadow map land in the volume (and this rank's z layers) this frame?  CTA-uniform.
__device__ __forceinline__ bool inject_block_active(const InjectLinear& lin, int b) {
    if (!lin.minmax) return true;
    const int S = lin.S, nbx = S / kInjBlockW;
    const float inv_s = 1.0f / (float)S, fd = (float)lin.D;
    const float2 mm = __ldg(lin.minmax + b);
    const int x0 = (b % nbx) * kInjBlockW, y0 = (b / nbx) * kInjBlockH;
    const float nx0 = ((float)x0 * inv_s) * 2.0f - 1.0f, nx1 = ((float)(x0 + kInjBlockW - 1) * inv_s) * 2.0f - 1.0f;
    const float ny0 = ((float)y0 * inv_s) * 2.0f - 1.0f, ny1 = ((float)(y0 + kInjBlockH - 1) * inv_s) * 2.0f - 1.0f;
    const float nz0 = mm.x * 2.0f - 1.0f, nz1 = mm.y * 2.0f - 1.0f;
    bool active = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ax0 = lin.bx[k] * nx0, ax1 = lin.bx[k] * nx1, ay0 = lin.by[k] * ny0, ay1 = lin.by[k] * ny1, az0 = lin.bz[k] * nz0, az1 = lin.bz[k] * nz1;
        const float lo = ((fminf(ax0, ax1) + fminf(ay0, ay1)) + fminf(az0, az1)) + lin.b0[k];
        const float hi = ((fmaxf(ax0, ax1) + fmaxf(ay0, ay1)) + fmaxf(az0, az1)) + lin.b0[k];
        if (hi < -1.0f - kInjectMargin || lo > fd + kInjectMargin) active = false;           // (NaN compares false: stays active)
        // this rank's z layers: (int)P truncates toward zero (P in (-1, 0) is layer 0); the interval widened by the margin, clamped to the volume
        if (k == 2 && active && !owns_any_z(lin.st, max((int)floorf(lo - kInjectMargin), 0), min((int)floorf(hi + kInjectMargin), lin.D - 1))) active = false;
    }
    return active;
}
// which 64x16 texel blocks of the shadow map can land in the volume (and this rank's slab); list[0] was zeroed by the host
__global__ void __launch_bounds__(256) k_inject_cull(const __grid_constant__ InjectLinear lin, uint32_t* __restrict__ list) {
    const int nb = (lin.S / kInjBlockW) * (lin.S / kInjBlockH), b = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    const bool active = b < nb && inject_block_active(lin, b);
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(list, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (active) list[1u + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)b;
}
// the finished count also goes to mapped host memory, tagged with the list's generation: later frames size their grid with it
__global__ void k_inject_publish(const uint32_t* __restrict__ list, unsigned* __restrict__ host_words, unsigned gen) {
    host_words[2] = list[0]; __threadfence_system(); host_words[3] = gen;
}
__device__ __forceinline__ float div_by_const(float a, float c, float rc) {
    const float q0 = __fmul_rn(a, rc);
    const float q1 = __fmaf_rn(__fmaf_rn(-q0, c, a), rc, q0);
    return __fmaf_rn(__fmaf_rn(-q1, c, a), rc, q1);
}
__global__ void __launch_bounds__(256) k_inject_linear(const float* __restrict__ shadow, const uint32_t* __restrict__ color, uint32_t* __restrict__ radiance,
                                                       const __grid_constant__ InjectLinear lin) {
    const int S = lin.S, D = lin.D;
    const float inv_s = 1.0f / (float)S, fd = (float)D;                     // exact: S is a power of two
    if (blockIdx.x >= __ldg(lin.list)) return;                              // CTA i works on the i-th active 64x16 texel block
    const int blk = (int)__ldg(lin.list + 1 + blockIdx.x), nbx = S / kInjBlockW;
    const int x0 = (blk % nbx) * kInjBlockW + (threadIdx.x & 15) * 4, y = (blk / nbx) * kInjBlockH + (threadIdx.x >> 4);
    const float* row1 = shadow + (size_t)y * S + x0;
    const float4 b = __ldcs(reinterpret_cast<const float4*>(row1));        // one-touch stream: evict first, keep L2 for the cone tracer's inputs
    const float bm1 = x0 > 0 ? __ldcs(row1 - 1) : 1.0f;                      // CLAMP_TO_BORDER, border 1
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f); float am1 = 1.0f;
    if (y > 0) { a = __ldcs(reinterpret_cast<const float4*>(row1 - S)); if (x0 > 0) am1 = __ldcs(row1 - S - 1); }
    const float ta[5] = {am1, a.x, a.y, a.z, a.w}, tb[5] = {bm1, b.x, b.y, b.z, b.w};
    // Texels that saw no geometry (all four depths exactly 1) unproject onto the far plane; when the host has shown that
    // plane to lie outside the volume, the bounds test below would reject them anyway.
    // (a flag, not a return: the lanes of the warp meet again in the shuffle below)
    const bool far = lin.skip_far && fminf(fminf(fminf(am1, bm1), fminf(fminf(a.x, a.y), fminf(a.z, a.w))), fminf(fminf(b.x, b.y), fminf(b.z, b.w))) == 1.0f;
    const float* m = lin.m;
    const float ny = ((float)y * inv_s) * 2.0f - 1.0f;
    const float my0 = m[4] * ny, my1 = m[5] * ny, my2 = m[6] * ny;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // shadow_linear() at the texel corner: weights 0.5/0.5 (k_inject), then injectRadiance.comp:46-52
        const float top = ta[k] * 0.5f + ta[k + 1] * 0.5f, bot = tb[k] * 0.5f + tb[k + 1] * 0.5f;
        const float d = top * 0.5f + bot * 0.5f;
        const float nx = ((float)(x0 + k) * inv_s) * 2.0f - 1.0f, nz = d * 2.0f - 1.0f;
        const float wx = ((m[0] * nx + my0) + m[8] * nz) + m[12];           // mul44 with w == 1
        const float wy = ((m[1] * nx + my1) + m[9] * nz) + m[13];
        const float wz = ((m[2] * nx + my2) + m[10] * nz) + m[14];
        const float px = fd * div_by_const(wx - lin.sub0[0] - lin.sub1[0], lin.c[0], lin.rc[0]);
        const float py = fd * div_by_const(wy - lin.sub0[1] - lin.sub1[1], lin.c[1], lin.rc[1]);
        const float pz = fd * div_by_const(wz - lin.sub0[2] - lin.sub1[2], lin.c[2], lin.rc[2]);
        int ix, iy, iz;
        o[k] = 0xFFFFFFFFu;
        if (!far && to_voxel_index(mk3(px, py, pz), D, ix, iy, iz) && owns_z(lin.st, iz)) o[k] = (uint32_t)((iz * D + iy) * D + ix);
    }
    uint32_t left = __shfl_up_sync(0xffffffffu, o[3], 1);
    if ((threadIdx.x & 15) == 0) left = 0xFFFFFFFFu;                        // 16 threads per texel row of the block
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (o[k] != 0xFFFFFFFFu && o[k] != left) radiance[o[k]] = __ldg(color + o[k]);     // packUnorm4x8(unpackUnorm4x8(c)) == c
        left = o[k];
    }
}

// ------------------------------------------------------------------------------------------- fill holes
__global__ void __launch_bounds__(kThreads) k_fill_holes(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int D, int z_lo, int z_hi) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), z = z_lo + blockIdx.z;
    if (x >= D || y >= D || z >= z_hi) return;
    const size_t o = ((size_t)z * D + y) * D + x;
    const uint32_t w = __ldg(src + o);
    V4 cur = unpack_unorm(w);
    if (cur.w == 0.0f) {
        float count = 0.0f;
        for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) for (int k = -1; k <= 1; ++k) {
            const int xx = x + i, yy = y + j, zz = z + k;
            if (xx < 0 || yy < 0 || zz < 0 || xx >= D || yy >= D || zz >= D) continue;
            const uint32_t nw = __ldg(src + ((size_t)zz * D + yy) * D + xx);
            if ((nw >> 24) != 0u) { const V4 v = unpack_unorm(nw); cur = mk4(cur.x + v.x, cur.y + v.y, cur.z + v.z, cur.w + v.w); count += 1.0f; }
        }
        if (count > 0.0f) cur = mk4(cur.x / count, cur.y / count, cur.z / count, cur.w / count);
    }
    dst[o] = pack_unorm(cur);
}

// -------------------------------------------------------------------------------------------------- mip
// BOX2: dst texel = 0.125 * sum of 8 children, accumulated in the shader's offset order
// (0,0,0),(0,0,1),(0,1,0),(0,1,1),(1,0,0),(1,0,1),(1,1,0),(1,1,1) where each (i,j,k) is an (x,y,z) offset.
// A 256-entry table of b/255.0f lives in shared memory (the division is exact-rounded, a table read is cheaper).
__device__ __forceinline__ void acc_word(float acc[4], uint32_t w, const float* __restrict__ lut) {
    acc[0] += lut[w & 255u]; acc[1] += lut[(w >> 8) & 255u]; acc[2] += lut[(w >> 16) & 255u]; acc[3] += lut[w >> 24];
}
__device__ __forceinline__ uint32_t finish_word(const float acc[4], float k) {
    return pack_unorm(mk4(acc[0] * k, acc[1] * k, acc[2] * k, acc[3] * k));
}
// one thread -> 4 consecutive dst texels in x: 8 x 16-byte loads, 1 x 16-byte store.  Requires Dd % 4 == 0.
__global__ void __launch_bounds__(kThreads) k_mip_box2(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int Ds, int zd_lo, int zd_hi) {
    __shared__ float lut[256];
    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int Dd = Ds >> 1, qx = Dd >> 2;                       // quads per dst row
    const size_t total = (size_t)qx * Dd * (zd_hi - zd_lo);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % qx); const size_t r = i / qx;
        const int y = (int)(r % Dd), z = zd_lo + (int)(r / Dd);
        const uint4* row00 = reinterpret_cast<const uint4*>(src + ((size_t)(2 * z) * Ds + 2 * y) * Ds) + 2 * q;        // (y0,z0)
        const uint4* row10 = reinterpret_cast<const uint4*>(src + ((size_t)(2 * z) * Ds + 2 * y + 1) * Ds) + 2 * q;    // (y1,z0)
        const uint4* row01 = reinterpret_cast<const uint4*>(src + ((size_t)(2 * z + 1) * Ds + 2 * y) * Ds) + 2 * q;    // (y0,z1)
        const uint4* row11 = reinterpret_cast<const uint4*>(src + ((size_t)(2 * z + 1) * Ds + 2 * y + 1) * Ds) + 2 * q;
        uint4 a[2], b[2], c[2], d[2];
        a[0] = __ldg(row00); a[1] = __ldg(row00 + 1); b[0] = __ldg(row01); b[1] = __ldg(row01 + 1);
        c[0] = __ldg(row10); c[1] = __ldg(row10 + 1); d[0] = __ldg(row11); d[1] = __ldg(row11 + 1);
        const uint32_t* A = reinterpret_cast<const uint32_t*>(a); const uint32_t* B = reinterpret_cast<const uint32_t*>(b);
        const uint32_t* Cc = reinterpret_cast<const uint32_t*>(c); const uint32_t* Dv = reinterpret_cast<const uint32_t*>(d);
        uint32_t out[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            // offsets in shader order: x0:(y0z0),(y0z1),(y1z0),(y1z1) then x1: same
            acc_word(acc, A[2 * t], lut); acc_word(acc, B[2 * t], lut); acc_word(acc, Cc[2 * t], lut); acc_word(acc, Dv[2 * t], lut);
            acc_word(acc, A[2 * t + 1], lut); acc_word(acc, B[2 * t + 1], lut); acc_word(acc, Cc[2 * t + 1], lut); acc_word(acc, Dv[2 * t + 1], lut);
            out[t] = finish_word(acc, 0.125f);
        }
        reinterpret_cast<uint4*>(dst + ((size_t)z * Dd + y) * Dd)[q] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}
// Whole BOX2 chain in ONE launch: a CTA of 128 threads owns a B^3 block of level 0 (B = 2^R, R <= 4), streams it
// once from HBM (one thread = 8 x 16-byte loads in flight), writes its (B/2)^3 block of level 1 and keeps it in
// shared memory, from which levels 2..R follow without touching HBM again; the last CTA to finish (ticket counter)
// reduces the few remaining coarse levels.  Every level is computed from the ROUNDED unorm8 words of the level below,
// exactly like the reference's one-dispatch-per-level loop (src/Application.cpp:889-902).  With `publish`, the
// level-0 words that pass through registers and every produced level are also written to the surfaces of the
// mipmapped array the cone tracer samples, which replaces the separate linear -> array copy.
constexpr int kMipThreads = 128;
struct MipChain {
    const uint32_t* src0; uint32_t* lvl[VCT_MAX_LEVELS];      // lvl[k] = linear level k (k >= 1 written)
    cudaSurfaceObject_t surf[VCT_MAX_LEVELS];
    uint8_t* pub_mask;                            // 1 byte per 8 level-0 texels of a row: array holds non-zero data there
    int publish;
};
struct MipChainArgs {
    MipChain chain[2]; int n_chains;              // radiance and colour pyramids filtered by one launch
    unsigned* ticket;                             // last-CTA detection for the tail levels
    int D, R, L, tail;                            // tail: reduce levels R..L-2 -> R+1..L-1 in the last CTA
    // sharded frames: the levels >= 1 that chain `push_chain` writes also go to the same place of every peer's pyramid (NVLink stores;
    // only bricks that hold or held something are written at all, so the exchange of the upper levels is as sparse as the scene)
    uint32_t* peer[VCT_MAX_PEERS]; unsigned long long lvl_off[VCT_MAX_LEVELS]; int n_peers, push_chain;
    Stripes st;                                   // blockIdx.z counts the bricks of this rank's stripes, back to back
    const uint8_t *seg_a, *seg_b;                 // sparse frames: this and last frame's segment masks (nullptr: dense)
};
__device__ __forceinline__ uint32_t box2_words(const uint32_t w[8], const float* __restrict__ lut) {
    if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0u) return 0u;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) acc_word(acc, w[k], lut);
    return finish_word(acc, 0.125f);
}
__global__ void __launch_bounds__(kMipThreads) k_mip_chain(const __grid_constant__ MipChainArgs args) {
    __shared__ float lut[256];
    __shared__ uint32_t s1[8 * 8 * 8], s2[4 * 4 * 4], s3[2 * 2 * 2];
    __shared__ unsigned s_last;
    lut[threadIdx.x] = (float)threadIdx.x / 255.0f; lut[threadIdx.x + 128] = (float)(threadIdx.x + 128) / 255.0f;   // visible after the first barrier below
    // one CTA = one 16^3 brick of BOTH pyramids (radiance, then colour): the segment masks are scanned once, and half as many CTAs pay
    // the fixed cost of a launch slot, the table set-up and the barriers (8192 -> 4096 CTAs at 256^3: 48 -> see DESIGN.md)
    const int B = 1 << args.R, H = B >> 1, D = args.D;
    const int bps = args.st.T >> args.R;          // bricks per stripe along z
    const int bx = blockIdx.x * B, by = blockIdx.y * B, bz = stripe_z(args.st, blockIdx.z / bps) + (blockIdx.z % bps) * B;
    // ---- level 0 -> 1: one thread = 4 consecutive level-1 texels (8 x 16-byte loads, 1 x 16-byte store)
    const int qx = H >> 2, nquads = qx * H * H, D1 = D >> 1;
    const bool masked = args.seg_a != nullptr;
    bool block_active = !masked;
    if (masked) {
        // A thread's four source rows are four mask segments.  Unflagged (now and last frame) means: zero in the linear
        // level, in the texture array, and in everything above it that this CTA would write.  A CTA with no flagged
        // segment at all has nothing to read and nothing to change.
        int any = 0;
        for (int q = threadIdx.x; q < nquads; q += kMipThreads) {
            const int lx = (q % qx) * 4, ly = (q / qx) % H, lz = q / (qx * H);
            const int x0 = bx + 2 * lx, y0 = by + 2 * ly, z0 = bz + 2 * lz;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t sidx = (((size_t)(z0 + (j & 1)) * D + y0 + (j >> 1)) * D + x0) >> 3;
                any |= args.seg_a[sidx] | args.seg_b[sidx];
            }
        }
        block_active = __syncthreads_or(any) != 0;
    } else __syncthreads();
    if (block_active)
    for (int which = 0; which < args.n_chains; ++which) {
    const MipChain& a = args.chain[which];
    const bool push = which == args.push_chain;
    if (which) __syncthreads();                                           // the shared levels of the previous chain have been read
    for (int q = threadIdx.x; q < nquads; q += kMipThreads) {
        const int lx = (q % qx) * 4, ly = (q / qx) % H, lz = q / (qx * H);
        const int x0 = bx + 2 * lx, y0 = by + 2 * ly, z0 = bz + 2 * lz;
        const uint4* r00 = reinterpret_cast<const uint4*>(a.src0 + ((size_t)z0 * D + y0) * D + x0);
        const uint4* r10 = reinterpret_cast<const uint4*>(a.src0 + ((size_t)z0 * D + y0 + 1) * D + x0);
        const uint4* r01 = reinterpret_cast<const uint4*>(a.src0 + ((size_t)(z0 + 1) * D + y0) * D + x0);
        const uint4* r11 = reinterpret_cast<const uint4*>(a.src0 + ((size_t)(z0 + 1) * D + y0 + 1) * D + x0);
        bool f[4] = {true, true, true, true};                               // rows (y0,z0) (y0,z1) (y1,z0) (y1,z1)
        if (masked) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t sidx = (((size_t)(z0 + (j & 1)) * D + y0 + (j >> 1)) * D + x0) >> 3;
                f[j] = (args.seg_a[sidx] | args.seg_b[sidx]) != 0;
            }
        }
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        uint4 v[8];
        v[0] = f[0] ? __ldcs(r00) : zero4; v[1] = f[0] ? __ldcs(r00 + 1) : zero4; v[2] = f[1] ? __ldcs(r01) : zero4; v[3] = f[1] ? __ldcs(r01 + 1) : zero4;
        v[4] = f[2] ? __ldcs(r10) : zero4; v[5] = f[2] ? __ldcs(r10 + 1) : zero4; v[6] = f[3] ? __ldcs(r11) : zero4; v[7] = f[3] ? __ldcs(r11 + 1) : zero4;
        if (a.publish) {
            // Level 0 goes to the texture array only where it is non-zero now or was non-zero in the array (the
            // volume is ~97 % empty and surface stores are the slowest part of this kernel): array == linear always.
            uint8_t* m[4]; uint8_t was[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { m[j] = a.pub_mask + (((size_t)(z0 + (j & 1)) * D + y0 + (j >> 1)) * D + x0) / 8; was[j] = f[j] ? *m[j] : (uint8_t)0; }   // loads in flight with v[]
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                   // rows (y0,z0) (y0,z1) (y1,z0) (y1,z1)
                if (!f[j]) continue;
                const int yy = y0 + (j >> 1), zz = z0 + (j & 1);
                const uint4 p = v[2 * j], r = v[2 * j + 1];
                const bool now = (p.x | p.y | p.z | p.w | r.x | r.y | r.z | r.w) != 0u;
                if (now || was[j]) { surf3Dwrite(p, a.surf[0], x0 * 4, yy, zz); surf3Dwrite(r, a.surf[0], x0 * 4 + 16, yy, zz); *m[j] = now ? 1 : 0; }
            }
        }
        const uint32_t A[8] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w};      // (y0,z0)
        const uint32_t Bq[8] = {v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};    // (y0,z1)
        const uint32_t Cq[8] = {v[4].x, v[4].y, v[4].z, v[4].w, v[5].x, v[5].y, v[5].z, v[5].w};    // (y1,z0)
        const uint32_t Dq[8] = {v[6].x, v[6].y, v[6].z, v[6].w, v[7].x, v[7].y, v[7].z, v[7].w};    // (y1,z1)
        uint32_t out[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            // shader offset order (x,y,z): (0,0,0),(0,0,1),(0,1,0),(0,1,1),(1,0,0),(1,0,1),(1,1,0),(1,1,1)
            const uint32_t w[8] = {A[2 * t], Bq[2 * t], Cq[2 * t], Dq[2 * t], A[2 * t + 1], Bq[2 * t + 1], Cq[2 * t + 1], Dq[2 * t + 1]};
            out[t] = box2_words(w, lut);
        }
        const uint4 o = make_uint4(out[0], out[1], out[2], out[3]);
        const int gx = (bx >> 1) + lx, gy = (by >> 1) + ly, gz = (bz >> 1) + lz;
        if (f[0] || f[1] || f[2] || f[3]) {                                  // otherwise level 1 is, and stays, zero here
            const size_t at = ((size_t)gz * D1 + gy) * D1 + gx;
            *reinterpret_cast<uint4*>(a.lvl[1] + at) = o;
            if (a.publish) surf3Dwrite(o, a.surf[1], gx * 4, gy, gz);
            if (push) for (int r = 0; r < args.n_peers; ++r) if (args.peer[r]) *reinterpret_cast<uint4*>(args.peer[r] + args.lvl_off[1] + at) = o;
        }
        *reinterpret_cast<uint4*>(s1 + (lz * H + ly) * H + lx) = o;
    }
    // ---- levels 1 -> 2 -> ... -> R from shared memory
    for (int k = 1; k < args.R; ++k) {
        __syncthreads();
        const int Hs = B >> k, Hd = Hs >> 1, Dd = D >> (k + 1);
        const uint32_t* src = k == 1 ? s1 : k == 2 ? s2 : s3;
        uint32_t* dsts = k == 1 ? s2 : k == 2 ? s3 : nullptr;
        uint32_t* gl = a.lvl[k + 1];
        const cudaSurfaceObject_t gs = a.surf[k + 1];
        for (int i = threadIdx.x; i < Hd * Hd * Hd; i += kMipThreads) {
            const int lx = i % Hd, ly = (i / Hd) % Hd, lz = i / (Hd * Hd);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = src[((2 * lz + (j & 1)) * Hs + 2 * ly + ((j >> 1) & 1)) * Hs + 2 * lx + (j >> 2)];
            const uint32_t o = box2_words(w, lut);
            const int gx = (bx >> (k + 1)) + lx, gy = (by >> (k + 1)) + ly, gz = (bz >> (k + 1)) + lz;
            const size_t at = ((size_t)gz * Dd + gy) * Dd + gx;
            gl[at] = o;
            if (a.publish) surf3Dwrite(o, gs, gx * 4, gy, gz);
            if (push) for (int r = 0; r < args.n_peers; ++r) if (args.peer[r]) args.peer[r][args.lvl_off[k + 1] + at] = o;
            if (dsts) dsts[(lz * Hd + ly) * Hd + lx] = o;
        }
    }
    }   // block_active
    // ---- tail: the last CTA reduces the remaining coarse levels (a few thousand texels) straight from L2
    if (!args.tail) return;
    if (block_active) __threadfence();                                     // an inactive CTA wrote nothing
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned t = atomicAdd(args.ticket, 1u);
        s_last = t == total - 1u;
        if (s_last) *args.ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int ch = 0; ch < args.n_chains; ++ch) {
        const MipChain& c = args.chain[ch];
        for (int l = args.R; l + 1 < args.L; ++l) {
            const int Ds = D >> l, Dd = Ds >> 1;
            const uint32_t* src = c.lvl[l]; uint32_t* dst = c.lvl[l + 1];
            for (int i = threadIdx.x; i < Dd * Dd * Dd; i += kMipThreads) {
                const int x = i % Dd, y = (i / Dd) % Dd, z = i / (Dd * Dd);
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = __ldcg(src + ((size_t)(2 * z + (j & 1)) * Ds + 2 * y + ((j >> 1) & 1)) * Ds + 2 * x + (j >> 2));
                const uint32_t o = box2_words(w, lut);
                dst[i] = o;
                if (c.publish) surf3Dwrite(o, c.surf[l + 1], x * 4, y, z);
            }
            __threadfence();
            __syncthreads();
        }
    }
}

// generic (any size, any kernel mode) — used for the small top levels and for BOX3 / CUBE
__global__ void __launch_bounds__(kThreads) k_mip_generic(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int Ds, int mode, int zd_lo, int zd_hi) {
    __shared__ float lut[256];
    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int Dd = Ds >> 1;
    const size_t total = (size_t)Dd * Dd * (zd_hi - zd_lo);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Dd); const size_t r = i / Dd; const int y = (int)(r % Dd), z = zd_lo + (int)(r / Dd);
        const int sx = 2 * x, sy = 2 * y, sz = 2 * z;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}; float k = 0.125f;
        auto ld = [&](int xx, int yy, int zz) { if (xx < 0 || yy < 0 || zz < 0 || xx >= Ds || yy >= Ds || zz >= Ds) return; acc_word(acc, __ldg(src + ((size_t)zz * Ds + yy) * Ds + xx), lut); };
        if (mode == 0) { for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) ld(sx + a, sy + b, sz + c); }
        else if (mode == 1) { for (int a = -1; a <= 1; ++a) for (int b = -1; b <= 1; ++b) for (int c = -1; c <= 1; ++c) ld(sx + a, sy + b, sz + c); k = 0.037f; }
        else { ld(sx, sy, sz); ld(sx, sy, sz + 1); ld(sx, sy + 1, sz); ld(sx + 1, sy, sz); ld(sx, sy, sz - 1); ld(sx, sy - 1, sz); ld(sx - 1, sy, sz); k = 0.143f; }
        dst[((size_t)z * Dd + y) * Dd + x] = finish_word(acc, k);
    }
}
// filter3d.comp:16-47 (dead variant): sampler fetches at texel centres with lod 0 -> mag NEAREST -> texel reads
__global__ void __launch_bounds__(kThreads) k_filter3d(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int Ds) {
    const int Dd = Ds >> 1;
    const size_t total = (size_t)Dd * Dd * Dd;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Dd); const size_t r = i / Dd; const int y = (int)(r % Dd), z = (int)(r / Dd);
        const float h = 0.5f * (1.0f / (float)Ds);
        const int bx = (int)floorf(((float)x / (float)Dd + h) * (float)Ds), by = (int)floorf(((float)y / (float)Dd + h) * (float)Ds),
                  bz = (int)floorf(((float)z / (float)Dd + h) * (float)Ds);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
            const int xx = bx + a, yy = by + b, zz = bz + c;
            if (xx < 0 || yy < 0 || zz < 0 || xx >= Ds || yy >= Ds || zz >= Ds) continue;
            const V4 t = unpack_unorm(__ldg(src + ((size_t)zz * Ds + yy) * Ds + xx));
            acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
        }
        dst[i] = finish_word(acc, 0.125f);
    }
}

// ---------------------------------------------------------------------------------------------- publish
// linear level -> surface of the mipmapped array (block-linear), 16 bytes per thread along x
__global__ void __launch_bounds__(kThreads) k_publish(const uint32_t* __restrict__ src, cudaSurfaceObject_t surf, int d) {
    const int qx = d >= 4 ? d >> 2 : 1;
    const size_t total = (size_t)qx * d * d;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % qx); const size_t r = i / qx; const int y = (int)(r % d), z = (int)(r / d);
        if (d >= 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)z * d + y) * d) + q);
            surf3Dwrite(v, surf, q * 16, y, z);
        } else {
            for (int x = 0; x < d; ++x) surf3Dwrite(__ldg(src + ((size_t)z * d + y) * d + x), surf, x * 4, y, z);
        }
    }
}

// levels 1..L-1 in one launch (they are small: 1/7 of level 0 together; one launch per level is launch-bound)
struct PublishUpper { const uint32_t* src[VCT_MAX_LEVELS]; cudaSurfaceObject_t surf[VCT_MAX_LEVELS]; int d[VCT_MAX_LEVELS]; unsigned long long first[VCT_MAX_LEVELS + 1]; int n; };
__global__ void __launch_bounds__(kThreads) k_publish_upper(const __grid_constant__ PublishUpper p) {
    const unsigned long long total = p.first[p.n];
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < p.n && i >= p.first[k + 1]) ++k;
        const int d = p.d[k];
        const unsigned long long j = i - p.first[k];
        const int qx = d >= 4 ? d >> 2 : 1;
        const int q = (int)(j % qx); const unsigned long long r = j / qx; const int y = (int)(r % d), z = (int)(r / d);
        const uint32_t* src = p.src[k];
        if (d >= 4) surf3Dwrite(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)z * d + y) * d) + q), p.surf[k], q * 16, y, z);
        else for (int x = 0; x < d; ++x) surf3Dwrite(__ldg(src + ((size_t)z * d + y) * d + x), p.surf[k], x * 4, y, z);
    }
}

}  // namespace

// ============================================================================================ host side
int vctk_clear_voxels(vct_ctx* c, bool reset_frame_counters) {
    // only this rank's z layers of level 0 are cleared (single GPU: the whole volume); dense frames are rare: one launch per stripe
    for (int k = 0; k < c->st.count; ++k) {
        const size_t off = (size_t)stripe_z(c->st, k) * c->D * c->D, n = (size_t)c->st.T * c->D * c->D;
        k_clear<<<grid_for(n / 4, kThreads), kThreads, 0, c->stream>>>(reinterpret_cast<uint4*>(c->d_color + off), reinterpret_cast<uint4*>(c->d_normal + off), n / 4,
                                                                      reset_frame_counters && k == 0 ? c->d_counters : nullptr);
        VCT_LAUNCH_CHECK(c, "k_clear");
    }
    return 0;
}
int vctk_transfer(vct_ctx* c) {
    const vct_frame_params& p = c->h_fc.p;
    for (int k = 0; k < c->st.count; ++k) {
        const size_t off = (size_t)stripe_z(c->st, k) * c->D * c->D, n = (size_t)c->st.T * c->D * c->D;
        // seg_mark is indexed relative to the stripe like the volume pointers (off voxels = off / 8 segments)
        k_transfer<<<grid_for(n / 8, kThreads), kThreads, 0, c->stream>>>(reinterpret_cast<uint4*>(c->d_color + off), reinterpret_cast<uint4*>(c->d_radiance + off), n / 4,
                                                                        p.voxel_set_opacity, p.temporal_filter_radiance, p.temporal_decay, c->d_counters,
                                                                        (p.temporal_filter_radiance && c->d_seg[c->seg_cur]) ? c->d_seg[c->seg_cur] + off / 8 : nullptr);
        VCT_LAUNCH_CHECK(c, "k_transfer");
    }
    return 0;
}
// Sparse frames need whole segments per x-row and the fused mip-chain kernel with 16^3 blocks (multi-GPU: inside the slab).
bool vctk_sparse_supported(const vct_ctx* c) {
    if (c->cfg.world_size > 1 && (c->st.T % 16 || c->st.count < 1)) return false;   // whole 16^3 blocks per stripe
    return !c->seg_disabled && c->D >= 16 && c->D % 16 == 0 && c->L >= 5 && c->d_seg[0] && c->d_seg[1];
}
int vctk_clear_masked(vct_ctx* c) {
    const size_t w_stripe = (size_t)c->D * c->D / 32 * c->st.T, n_words = w_stripe * c->st.count;   // mask words of this rank's stripes
    const int temporal = c->h_fc.p.temporal_filter_radiance;
    k_clear_masked<<<grid_for(n_words, kThreads), kThreads, 0, c->stream>>>(reinterpret_cast<uint4*>(c->d_color), reinterpret_cast<uint4*>(c->d_normal), reinterpret_cast<uint4*>(c->d_radiance),
                                                                            reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur ^ 1]), reinterpret_cast<uint32_t*>(c->d_seg[c->seg_cur]), c->st, w_stripe, n_words,
                                                                            temporal, c->d_counters);
    VCT_LAUNCH_CHECK(c, "k_clear");
    return 0;
}
int vctk_frame_begin_masked(vct_ctx* c) {
    const size_t w_stripe = (size_t)c->D * c->D / 32 * c->st.T, n_words = w_stripe * c->st.count;
    TransformArgs t{c->d_vertices, c->d_vactor, c->d_models, c->d_nmats, c->n_vertices, c->d_wpos, c->d_wnrm, c->d_wT, c->d_wB};
    const unsigned tb = (unsigned)std::min<size_t>((c->n_vertices + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    const unsigned cb = (unsigned)grid_for(n_words, kThreads);
    k_frame_begin<<<tb + cb, kThreads, 0, c->stream>>>(t, tb, reinterpret_cast<uint4*>(c->d_color), reinterpret_cast<uint4*>(c->d_normal), reinterpret_cast<uint4*>(c->d_radiance),
                                                       reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur ^ 1]), reinterpret_cast<uint32_t*>(c->d_seg[c->seg_cur]), c->st, w_stripe, n_words,
                                                       c->h_fc.p.temporal_filter_radiance, c->d_counters);
    VCT_LAUNCH_CHECK(c, "k_frame_begin");
    return 0;
}
int vctk_transfer_masked(vct_ctx* c) {
    const vct_frame_params& p = c->h_fc.p;
    const size_t w_stripe = (size_t)c->D * c->D / 32 * c->st.T, n_words = w_stripe * c->st.count;
    k_transfer_masked<<<grid_for(n_words, kThreads), kThreads, 0, c->stream>>>(reinterpret_cast<uint4*>(c->d_color), reinterpret_cast<uint4*>(c->d_radiance),
                                                                               reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur]), c->st, w_stripe, n_words,
                                                                               p.voxel_set_opacity, p.temporal_filter_radiance, p.temporal_decay, c->d_counters);
    VCT_LAUNCH_CHECK(c, "k_transfer");
    return 0;
}
int vctk_inject(vct_ctx* c) {
    const vct_frame_params& p = c->h_fc.p;
    const bool pow2 = (c->S & (c->S - 1)) == 0 && c->S >= 32;
    if (pow2 && c->S >= 64 && !p.warp_voxels && !p.warp_texture && !p.voxelize_tesselation_warp && !p.radiance_lighting && c->D <= 1024) {
        InjectLinear lin; memset(&lin, 0, sizeof lin); bool sane = true;   // (zeroed padding: the struct is compared bytewise below)
        for (int i = 0; i < 3; ++i) {
            volatile float ext = p.voxel_max[i] - p.voxel_min[i];           // one fp32 rounding, like the shader's (max - min)
            lin.c[i] = ext; lin.rc[i] = (float)(1.0 / (double)lin.c[i]); lin.sub0[i] = p.voxel_center[i]; lin.sub1[i] = p.voxel_min[i];
            sane = sane && lin.c[i] > 1e-6f && lin.c[i] < 1e6f;            // well inside the normal range: no under/overflow in the corrections
        }
        if (sane) {
            memcpy(lin.m, c->h_fc.ls_inverse.m, 64);
            lin.S = c->S; lin.D = c->D; lin.st = c->st;
            lin.log2_qx = 0; while ((4 << lin.log2_qx) < c->S) lin.log2_qx++;
            {   // far plane (ndc z = 1) of the light frustum in voxel coordinates: affine image of a quad, extremes at the corners
                double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
                for (int k = 0; k < 4; ++k) {
                    const double nx = (k & 1) ? 1.0 : -1.0, ny = (k & 2) ? 1.0 : -1.0;
                    for (int i = 0; i < 3; ++i) {
                        const double w = lin.m[i] * nx + lin.m[4 + i] * ny + lin.m[8 + i] + lin.m[12 + i];
                        const double v = (w - lin.sub0[i] - lin.sub1[i]) / lin.c[i] * c->D;
                        lo[i] = std::min(lo[i], v); hi[i] = std::max(hi[i], v);
                    }
                }
                lin.skip_far = 0;
                for (int i = 0; i < 3; ++i) if (hi[i] < -2.0 || lo[i] > c->D + 2.0) lin.skip_far = 1;      // (NaN compares false: no skip)
            }
            const size_t n_fine = (size_t)(c->S / 4) * (c->S / 4);
            lin.minmax = c->shadow_mm_valid ? reinterpret_cast<const float2*>(c->d_shadow_mm) + n_fine : nullptr;
            for (int i = 0; i < 3; ++i) {                                   // D * ((ls_inverse * ndc)[i] - center - min) / (max - min), in double
                const double k = (double)c->D / (double)lin.c[i];
                lin.bx[i] = (float)(k * lin.m[i]); lin.by[i] = (float)(k * lin.m[4 + i]); lin.bz[i] = (float)(k * lin.m[8 + i]);
                lin.b0[i] = (float)(k * ((double)lin.m[12 + i] - (double)lin.sub0[i] - (double)lin.sub1[i]));
            }
            lin.list = c->d_inject_list;
            // (re)build the active-block list when the shadow map or anything the test reads has changed
            InjectLinear key = lin; key.minmax = nullptr; key.list = nullptr; key.skip_far = (int)c->shadow_gen;
            static_assert(sizeof(InjectLinear) <= sizeof(c->inject_key), "inject_key");
            if (!c->inject_list_valid || memcmp(&key, c->inject_key, sizeof key) != 0) {
                const int nb = (c->S / kInjBlockW) * (c->S / kInjBlockH);
                VCT_CHECK(c, cudaMemsetAsync(c->d_inject_list, 0, 4, c->stream));
                vct_prof_mark(c, "memset");
                k_inject_cull<<<(nb + 255) / 256, 256, 0, c->stream>>>(lin, c->d_inject_list);
                VCT_LAUNCH_CHECK(c, "k_inject_cull");
                memcpy(c->inject_key, &key, sizeof key); c->inject_list_valid = true;
                c->inject_gen++;
                if (c->h_overflow) {
                    unsigned* dp = nullptr;
                    if (cudaHostGetDevicePointer((void**)&dp, c->h_overflow, 0) == cudaSuccess) { k_inject_publish<<<1, 1, 0, c->stream>>>(c->d_inject_list, dp, c->inject_gen); VCT_LAUNCH_CHECK(c, "k_inject_cull"); }
                }
            }
            // one CTA per LISTED block once the host has seen the count of this list (mapped words [2], [3]); until then one per block
            unsigned grid = (unsigned)((c->S / kInjBlockW) * (c->S / kInjBlockH));
            if (c->h_overflow) {
                const volatile unsigned* hw = (const volatile unsigned*)c->h_overflow;
                if (hw[3] == c->inject_gen) { const unsigned n = hw[2]; if (n < grid) grid = n; }
            }
            if (grid) k_inject_linear<<<grid, 256, 0, c->stream>>>(c->d_shadow, c->d_color, c->d_radiance, lin);
            VCT_LAUNCH_CHECK(c, "k_inject");
            return 0;
        }
    }
    if (pow2) {
        k_inject<<<(unsigned)((size_t)c->S * c->S / 4 / 256), 256, 0, c->stream>>>(c->d_fc, c->d_shadow, c->d_color, c->d_normal, c->d_warpmap, c->d_radiance);
        VCT_LAUNCH_CHECK(c, "k_inject");
        return 0;
    }
    dim3 grid((c->S + 31) / 32, (c->S + 7) / 8);
    k_inject_generic<<<grid, 256, 0, c->stream>>>(c->d_fc, c->d_shadow, c->d_color, c->d_normal, c->d_warpmap, c->d_radiance);
    VCT_LAUNCH_CHECK(c, "k_inject_generic");
    return 0;
}
// after every write of the shadow map (vct_shadowmap, vct_write_shadowmap): block depth bounds for the inject pass
int vctk_shadow_minmax(vct_ctx* c) {
    c->shadow_mm_valid = false; c->shadow_gen++;             // a new shadow map: the inject pass rebuilds its block list
    if (!c->d_shadow_mm || (c->S & (c->S - 1)) || c->S < 32) return 0;
    const int nb = c->S / 4;
    if (c->S < 64) return 0;
    float2* fine = reinterpret_cast<float2*>(c->d_shadow_mm);
    k_shadow_minmax<<<(nb * nb + 255) / 256, 256, 0, c->stream>>>(c->d_shadow, c->S, fine);
    VCT_LAUNCH_CHECK(c, "k_shadow_minmax");
    const int ncoarse = (c->S / kInjBlockW) * (c->S / kInjBlockH);
    k_shadow_minmax_coarse<<<(ncoarse + 255) / 256, 256, 0, c->stream>>>(fine, c->S, fine + (size_t)nb * nb);
    VCT_LAUNCH_CHECK(c, "k_shadow_minmax");
    c->shadow_mm_valid = true;
    return 0;
}
int vctk_fill_holes(vct_ctx* c) {
    dim3 grid((c->D + 31) / 32, (c->D + 7) / 8, c->D);     // (single GPU only: upload_frame refuses voxel_fill_holes with world_size > 1)
    k_fill_holes<<<grid, kThreads, 0, c->stream>>>(c->d_radiance, c->d_scratch, c->D, 0, c->D);
    VCT_LAUNCH_CHECK(c, "k_fill_holes");
    const size_t off = 0, n = (size_t)c->D * c->D * c->D;
    VCT_CHECK(c, cudaMemcpyAsync(c->d_radiance + off, c->d_scratch + off, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    vct_prof_mark(c, "memcpy_d2d");
    return 0;
}
static int publish_levels(vct_ctx* c, int which, int l_begin, int l_end) {
    uint32_t* base = which == VCT_VOL_COLOR ? c->d_color : c->d_radiance;
    uint8_t* mask = which == VCT_VOL_COLOR ? c->d_pub_mask_color : c->d_pub_mask_radiance;
    if (l_begin == 0 && mask) {                     // the array now mirrors the linear level 0: anything may be non-zero
        VCT_CHECK(c, cudaMemsetAsync(mask, 0xFF, (size_t)c->D * c->D * c->D / 8, c->stream));
        vct_prof_mark(c, "memset");
    }
    cudaSurfaceObject_t* surf = which == VCT_VOL_COLOR ? c->color_surf : c->radiance_surf;
    for (int l = l_begin; l < l_end; ++l) {
        const int d = level_dim(c->D, l);
        const size_t items = (size_t)(d >= 4 ? d / 4 : 1) * d * d;
        k_publish<<<grid_for(items, kThreads), kThreads, 0, c->stream>>>(base + c->level_off[l], surf[l], d);
        VCT_LAUNCH_CHECK(c, "k_publish");
    }
    return 0;
}
// levels 1..L-1 (the reference's last loop iteration targets a non-existent level: Application.cpp:889-902) of up to
// two pyramids.  publish[i] != 0 (single GPU): that pyramid also lands in the mipmapped array the cone tracer samples.
int vctk_mip_chains(vct_ctx* c, int n, const int* which, const int* publish_in, int mode, bool masked, bool push_to_peers) {
    c->mip_pushed_upto = 0;
    int publish[2] = {0, 0};
    uint32_t* base[2]; cudaSurfaceObject_t* surf[2];
    for (int i = 0; i < n; ++i) {
        base[i] = which[i] == VCT_VOL_COLOR ? c->d_color : c->d_radiance;
        surf[i] = which[i] == VCT_VOL_COLOR ? c->color_surf : c->radiance_surf;
        publish[i] = publish_in[i] && surf[i][0];
    }
    int first = 0;                                  // first source level still to be filtered
    int published_upto = 0;                         // levels [0, published_upto) are already in the array
    // fused chain: R reductions per CTA, block edge 2^R; needs whole blocks inside this rank's stripes
    const Stripes& st = c->st;
    const bool multi = c->cfg.world_size > 1;
    int R = c->L - 1 < 4 ? c->L - 1 : 4;
    while (R >= 3 && (st.T % (1 << R) || c->D % (1 << R))) R--;
    if (mode == 0 && R >= 3) {
        MipChainArgs a{};
        for (int i = 0; i < n; ++i) {
            MipChain& ch = a.chain[i];
            ch.src0 = base[i] + c->level_off[0];
            for (int k = 1; k < c->L; ++k) ch.lvl[k] = base[i] + c->level_off[k];
            for (int k = 0; k < c->L; ++k) ch.surf[k] = surf[i][k];
            ch.publish = publish[i];
            ch.pub_mask = which[i] == VCT_VOL_COLOR ? c->d_pub_mask_color : c->d_pub_mask_radiance;
            if (publish[i] && !ch.pub_mask) { c->error = "vct_mip: publish mask missing"; return 1; }
        }
        a.n_chains = n; a.D = c->D; a.R = R; a.L = c->L; a.st = st;
        a.push_chain = -1; a.n_peers = 0;
        if (push_to_peers && multi) {
            const int traced = c->h_fc.p.draw_radiance ? VCT_VOL_RADIANCE : VCT_VOL_COLOR;
            for (int i = 0; i < n; ++i) if (which[i] == traced) a.push_chain = i;
            if (a.push_chain >= 0) {
                if (vctk_xchg_mip_peers(c, a.peer)) return 1;
                a.n_peers = c->cfg.world_size;
                for (int k = 0; k < c->L; ++k) a.lvl_off[k] = c->level_off[k];
                c->mip_pushed_upto = R;
            }
        }
        a.tail = !multi && c->L - 1 > R;                        // sharded: the levels above the stripes follow the exchange (vctk_mip_tail)
        a.ticket = &c->d_counters->mip_ticket;
        a.seg_a = masked && R == 4 ? c->d_seg[c->seg_cur] : nullptr; a.seg_b = masked && R == 4 ? c->d_seg[c->seg_cur ^ 1] : nullptr;
        const int B = 1 << R;
        dim3 grid(c->D / B, c->D / B, st.T / B * st.count);
        k_mip_chain<<<grid, kMipThreads, 0, c->stream>>>(a);
        VCT_LAUNCH_CHECK(c, "k_mip_chain");
        first = a.tail ? c->L - 1 : R;
        published_upto = first + 1;
    }
    const int top = multi ? vctk_mip_top_sharded_level(c) : c->L - 1;     // sharded: the last level whose z layers lie inside one stripe
    for (int i = 0; i < n; ++i) {
        for (int l = first; l + 1 <= top; ++l) {
            const int Ds = level_dim(c->D, l), Dd = Ds >> 1;
            if (Dd < 1) break;
            const uint32_t* src = base[i] + c->level_off[l]; uint32_t* dst = base[i] + c->level_off[l + 1];
            for (int k = 0; k < (multi ? st.count : 1); ++k) {
                // z layers of the destination level owned by this rank: one run per stripe
                const int zd_lo = multi ? stripe_z(st, k) >> (l + 1) : 0, zd_hi = multi ? (stripe_z(st, k) + st.T) >> (l + 1) : Dd;
                if (mode == 0 && Dd % 4 == 0) {
                    const size_t items = (size_t)(Dd / 4) * Dd * (zd_hi - zd_lo);
                    k_mip_box2<<<grid_for(items, kThreads), kThreads, 0, c->stream>>>(src, dst, Ds, zd_lo, zd_hi);
                    VCT_LAUNCH_CHECK(c, "k_mip_box2");
                } else {
                    const size_t items = (size_t)Dd * Dd * (zd_hi - zd_lo);
                    k_mip_generic<<<grid_for(items, kThreads), kThreads, 0, c->stream>>>(src, dst, Ds, mode, zd_lo, zd_hi);
                    VCT_LAUNCH_CHECK(c, "k_mip_generic");
                }
            }
        }
        const int pub_end = top + 1;
        if (publish[i] && published_upto < pub_end && publish_levels(c, which[i], published_upto, pub_end)) return 1;
    }
    return 0;
}
// Sharded frames: the highest level a rank can filter from its own stripes (texel layers do not straddle a stripe boundary).
int vctk_mip_top_sharded_level(const vct_ctx* c) {
    int l = 0;
    while (l + 1 < c->L && c->st.T % (1 << (l + 1)) == 0) ++l;
    return l;
}
// Sharded frames, after the exchange has completed level `top` on every rank: the few small levels above it, redundantly on every rank
// (16^3 texels and fewer at the default stripe of 16 layers) — cheaper than another exchange round.
int vctk_mip_tail(vct_ctx* c, int which) {
    uint32_t* base = which == VCT_VOL_COLOR ? c->d_color : c->d_radiance;
    for (int l = vctk_mip_top_sharded_level(c); l + 1 < c->L; ++l) {
        const int Ds = level_dim(c->D, l), Dd = Ds >> 1;
        if (Dd < 1) break;
        const size_t items = (size_t)Dd * Dd * Dd;
        k_mip_generic<<<grid_for(items, kThreads), kThreads, 0, c->stream>>>(base + c->level_off[l], base + c->level_off[l + 1], Ds, 0, 0, Dd);
        VCT_LAUNCH_CHECK(c, "k_mip_generic");
    }
    return 0;
}
int vctk_mip(vct_ctx* c, int which, int mode, int publish) { return vctk_mip_chains(c, 1, &which, &publish, mode); }
int vctk_publish(vct_ctx* c, int which) { return publish_levels(c, which, 0, c->L); }
int vctk_publish_upper(vct_ctx* c, int which) {
    if (c->L < 2) return 0;
    uint32_t* base = which == VCT_VOL_COLOR ? c->d_color : c->d_radiance;
    cudaSurfaceObject_t* surf = which == VCT_VOL_COLOR ? c->color_surf : c->radiance_surf;
    PublishUpper p{};
    unsigned long long off = 0;
    for (int l = 1; l < c->L; ++l) {
        const int d = level_dim(c->D, l), k = p.n++;
        p.src[k] = base + c->level_off[l]; p.surf[k] = surf[l]; p.d[k] = d; p.first[k] = off;
        off += (unsigned long long)(d >= 4 ? d / 4 : 1) * d * d;
    }
    p.first[p.n] = off;
    k_publish_upper<<<grid_for((size_t)off, kThreads), kThreads, 0, c->stream>>>(p);
    VCT_LAUNCH_CHECK(c, "k_publish");
    return 0;
}
int vctk_set_voxel_opacity(vct_ctx* c, float opacity) {
    const size_t n = (size_t)c->D * c->D * c->D;
    k_set_voxel_opacity<<<grid_for(n, kThreads), kThreads, 0, c->stream>>>(c->d_color, c->d_radiance, n, opacity, c->d_counters);
    VCT_LAUNCH_CHECK(c, "k_set_voxel_opacity");
    return 0;
}
int vctk_temporal_radiance_filter(vct_ctx* c, float decay) {
    const size_t n = (size_t)c->D * c->D * c->D;
    k_temporal_decay<<<grid_for(n, kThreads), kThreads, 0, c->stream>>>(c->d_radiance, n, decay);
    VCT_LAUNCH_CHECK(c, "k_temporal_decay");
    return 0;
}
int vctk_filter3d(vct_ctx* c, int which, int src_level) {
    if (src_level < 0 || src_level + 1 >= c->L) { c->error = "filter3d: level out of range"; return 1; }
    uint32_t* base = which == VCT_VOL_COLOR ? c->d_color : c->d_radiance;
    const int Ds = level_dim(c->D, src_level);
    const size_t n = (size_t)(Ds / 2) * (Ds / 2) * (Ds / 2);
    k_filter3d<<<grid_for(n, kThreads), kThreads, 0, c->stream>>>(base + c->level_off[src_level], base + c->level_off[src_level + 1], Ds);
    VCT_LAUNCH_CHECK(c, "k_filter3d");
    return 0;
}
int vctk_normalize_voxels_f16(vct_ctx* c, void* col, void* nrm, float opacity) {
    const size_t n = (size_t)c->D * c->D * c->D;
    k_normalize_f16<<<grid_for(n, kThreads), kThreads, 0, c->stream>>>((__half*)col, (__half*)nrm, c->d_radiance, n, opacity, c->d_counters);
    VCT_LAUNCH_CHECK(c, "k_normalize_f16");
    return 0;
}
