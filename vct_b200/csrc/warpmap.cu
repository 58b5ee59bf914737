// warpmap.cu — the warped-grid warp map (a9), generated entirely on the device (compiled with -fmad=false).
//
// The reference reads the 32^3 occupancy grid back to the CPU, builds three prefix-sum tables and a weight table
// there, uploads them and draws two layered quads (src/Application.cpp:303-577, generateWarpmapWeights.frag:37-74,
// generateWarpmap.frag:44-100) — a GPU->CPU->GPU bubble every frame.  Here 32 CTAs of 1024 threads do all of it, one
// z-slice of the warp map each: the occupancy becomes three bit tables in shared memory, prefix sums are popcounts of masked words,
// the weight table (Application.cpp:346-370) is 33 entries in shared memory, and both reference passes run per texel without a round
// trip through memory (the weight texel that pass 2 fetches is recomputed in place).
// Quirk kept: the quad is drawn at 0.8 scale (quad.vert:12) so only texels 3..28 in x and y are
// written, with tc = ((i+.5)/32-.5)/.8+.5; unwritten texels stay 0.
#include <cuda_fp16.h>

#include "raster.cuh"

namespace {
constexpr int N = VCT_WARP_DIM;

__device__ __forceinline__ bool quad_covered(int i) { const float c = (float)i + 0.5f; return c >= 3.2f && c <= 28.8f; }
__device__ __forceinline__ float quad_tc(int i) { return (((float)i + 0.5f) / (float)N - 0.5f) / 0.8f + 0.5f; }

// Occupancy of the 32^3 grid as three bit tables in shared memory: bx[z][y] bit x, by[z][x] bit y, bz[y][x] bit z.  Inclusive prefix counts
// along an axis (Application.cpp:315-343) are popcounts of a masked word; the row totals are popcounts of the whole word.
struct OccBits { uint32_t bx[N][N], by[N][N], bz[N][N]; };
__device__ __forceinline__ uint32_t upto(int i) { return 0xffffffffu >> (31 - i); }
__device__ __forceinline__ int tot_x(const OccBits& b, int y, int z) { return __popc(b.bx[z][y]); }
__device__ __forceinline__ int tot_y(const OccBits& b, int x, int z) { return __popc(b.by[z][x]); }
__device__ __forceinline__ int tot_z(const OccBits& b, int x, int y) { return __popc(b.bz[y][x]); }

// grid = 32 CTAs, one z-slice of the warp map each; thread = one texel (warp = y, lane = x).  Every CTA builds the bit tables from the
// whole occupancy grid (128 KiB from L2, 32 independent loads per thread), so nothing is exchanged between CTAs.
__global__ void __launch_bounds__(1024) k_warpmap(const uint32_t* __restrict__ occ, const FrameConst* __restrict__ fcp, ushort4* __restrict__ warpmap,
                                                  ushort4* __restrict__ wlo, ushort4* __restrict__ whi) {
    const vct_frame_params& p = fcp->p;
    __shared__ float wl[N + 1], wh[N + 1];
    __shared__ OccBits B;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid <= N) {                                                       // Application.cpp:346-370
        const int o = tid;
        if (o == 0 || o == N) { wl[o] = 1.0f; wh[o] = 1.0f; }
        else {
            const int empty = N - o;
            float h = p.warp_texture_high_resolution, l = ((float)N - h * (float)o) / (float)empty;
            if (l < p.warp_texture_low_resolution) { l = p.warp_texture_low_resolution; h = ((float)N - l * (float)empty) / (float)o; }
            wl[o] = l; wh[o] = h;
        }
    }
    {
        uint32_t v[N];
#pragma unroll
        for (int z = 0; z < N; ++z) v[z] = __ldg(occ + (z * N + w) * N + lane);          // warp = y, lane = x
#pragma unroll
        for (int z = 0; z < N; ++z) { const unsigned m = __ballot_sync(0xffffffffu, v[z] > 0u); if (lane == 0) B.bx[z][w] = m; }
    }
    __syncthreads();
    {
        uint32_t my = 0u, mz = 0u;                                        // by[w][lane]: bit y of column (x = lane, z = w); bz[w][lane]: bit z of column (x = lane, y = w)
#pragma unroll
        for (int k = 0; k < N; ++k) { my |= (B.bx[w][k] >> lane & 1u) << k; mz |= (B.bx[k][w] >> lane & 1u) << k; }
        B.by[w][lane] = my; B.bz[w][lane] = mz;
    }
    __syncthreads();
    const int x = lane, y = w, z = blockIdx.x, i = (z * N + y) * N + x;
    // the quad is drawn at 0.8 scale: texels outside it keep the cleared value 0 in all three targets
    if (!quad_covered(x) || !quad_covered(y)) {
        warpmap[i] = make_ushort4(0, 0, 0, 0); wlo[i] = make_ushort4(0, 0, 0, 0); whi[i] = make_ushort4(0, 0, 0, 0);
        return;
    }
    const float tc[3] = {quad_tc(x), quad_tc(y), ((float)z + 0.5f) / (float)N};
    // pass 1 of the reference: generateWarpmapWeights.frag:37-74 (two RGBA16F targets) for a texel (tx, ty, tz) of the weight volumes
    auto weights_texel = [&](int tx, int ty, int tz, ushort4& a, ushort4& b) {
        if (!p.use_warpmap_weights_texture || !quad_covered(tx) || !quad_covered(ty)) { a = make_ushort4(0, 0, 0, 0); b = make_ushort4(0, 0, 0, 0); return; }
        const float t[3] = {quad_tc(tx), quad_tc(ty), ((float)tz + 0.5f) / (float)N};
        const int cx = (int)truncf(t[0] * (float)N), cy = (int)truncf(t[1] * (float)N), cz = (int)truncf(t[2] * (float)N);
        const bool od = (B.bx[cz][cy] >> cx & 1u) != 0u;
        const int tot[3] = {tot_x(B, cy, cz), tot_y(B, cx, cz), tot_z(B, cx, cy)};
        a = make_ushort4(__half_as_ushort(__float2half_rn(wl[tot[0]])), __half_as_ushort(__float2half_rn(wl[tot[1]])),
                         __half_as_ushort(__float2half_rn(wl[tot[2]])), __half_as_ushort(__float2half_rn(od ? 0.0f : 1.0f)));
        b = make_ushort4(__half_as_ushort(__float2half_rn(wh[tot[0]])), __half_as_ushort(__float2half_rn(wh[tot[1]])),
                         __half_as_ushort(__float2half_rn(wh[tot[2]])), __half_as_ushort(__float2half_rn(od ? 1.0f : 0.0f)));
    };
    {
        ushort4 a, b;
        weights_texel(x, y, z, a, b);
        wlo[i] = a; whi[i] = b;
    }
    // pass 2: generateWarpmap.frag:44-100
    int cid[3]; float fr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { const float lt = tc[k] * (float)N, fl = truncf(lt); fr[k] = lt - fl; cid[k] = (int)fl; }
    const bool od = (B.bx[cid[2]][cid[1]] >> cid[0] & 1u) != 0u;
    const int part[3] = {__popc(B.bx[cid[2]][cid[1]] & upto(cid[0])), __popc(B.by[cid[2]][cid[0]] & upto(cid[1])), __popc(B.bz[cid[1]][cid[0]] & upto(cid[2]))};
    const int tot[3] = {tot_x(B, cid[1], cid[2]), tot_y(B, cid[0], cid[2]), tot_z(B, cid[0], cid[1])};
    float lo[3], hi[3];
    if (p.use_warpmap_weights_texture) {                                  // NEAREST / CLAMP_TO_EDGE fetch at tc: that texel of pass 1, recomputed
        int t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = min(max((int)floorf(tc[k] * (float)N), 0), N - 1);
        ushort4 a, b;
        weights_texel(t[0], t[1], t[2], a, b);
        lo[0] = __half2float(__ushort_as_half(a.x)); lo[1] = __half2float(__ushort_as_half(a.y)); lo[2] = __half2float(__ushort_as_half(a.z));
        hi[0] = __half2float(__ushort_as_half(b.x)); hi[1] = __half2float(__ushort_as_half(b.y)); hi[2] = __half2float(__ushort_as_half(b.z));
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = wl[tot[k]]; hi[k] = wh[tot[k]]; }
    }
    float out[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float res = od ? hi[k] : lo[k];
        const float prev = od ? (float)part[k] - 1.0f : (float)part[k];
        const float off = lo[k] * ((float)cid[k] - prev) + hi[k] * prev;
        const float inner = fr[k] * res;
        const float warped = (off + inner) / (float)N;
        const bool use = p.warp_texture_linear ? false : p.warp_texture_axes[k] != 0;
        out[k] = use ? warped : tc[k];
    }
    const unsigned bits = ((unsigned)tot[0] & 31u) | ((unsigned)tot[1] & 31u) << 5 | ((unsigned)tot[2] & 31u) << 10 | (od ? 1u : 0u) << 15;
    out[3] = (float)bits / 65535.0f;
    unsigned short q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { float v = out[k]; if (!(v > 0.0f)) v = 0.0f; if (v > 1.0f) v = 1.0f; q[k] = (unsigned short)__float2uint_rn(v * 65535.0f); }
    warpmap[i] = make_ushort4(q[0], q[1], q[2], q[3]);
}
// Float copy of the warp map for the cone tracer and the voxeliser: the unorm16 -> float conversion (q / 65535, one IEEE division per channel) once
// per texel here instead of 24 times per cone step there (config 4 at 4K: 68 ms -> 9 ms); same values, bit for bit.  Also run after the host
// writes a warp map (vct_write_volume).
__global__ void __launch_bounds__(256) k_warpmap_floats(const ushort4* __restrict__ warpmap, float4* __restrict__ wf) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM) return;
    const ushort4 q = warpmap[i];
    // entry i = texel i and its +x neighbour (CLAMP_TO_EDGE): the tracer fetches both x corners of a trilinear cell with one 256-bit load
    const ushort4 r = warpmap[(i % VCT_WARP_DIM) == VCT_WARP_DIM - 1 ? i : i + 1];
    wf[2 * i] = make_float4((float)q.x / 65535.0f, (float)q.y / 65535.0f, (float)q.z / 65535.0f, (float)q.w / 65535.0f);
    wf[2 * i + 1] = make_float4((float)r.x / 65535.0f, (float)r.y / 65535.0f, (float)r.z / 65535.0f, (float)r.w / 65535.0f);
}
}  // namespace

int vctk_warpmap_floats(vct_ctx* c) {
    const int n = VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM;
    k_warpmap_floats<<<(n + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const ushort4*>(c->d_warpmap), reinterpret_cast<float4*>(c->d_warpmap + 4 * (size_t)n));
    VCT_LAUNCH_CHECK(c, "k_warpmap_floats");
    return 0;
}

int vctk_warpmap(vct_ctx* c) {
    k_warpmap<<<N, 1024, 0, c->stream>>>(c->d_occ, c->d_fc, reinterpret_cast<ushort4*>(c->d_warpmap), reinterpret_cast<ushort4*>(c->d_wlo),
                                         reinterpret_cast<ushort4*>(c->d_whi));
    VCT_LAUNCH_CHECK(c, "k_warpmap");
    return vctk_warpmap_floats(c);
}
