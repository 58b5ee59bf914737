// warpmap.cu — the warped-grid warp map (a9), generated entirely on the device (compiled with -fmad=false).
//
// The reference reads the 32^3 occupancy grid back to the CPU, builds three prefix-sum tables and a weight table
// there, uploads them and draws two layered quads (src/Application.cpp:303-577, generateWarpmapWeights.frag:37-74,
// generateWarpmap.frag:44-100) — a GPU->CPU->GPU bubble every frame.  Here one 1024-thread CTA does all of it:
// warp w / lane l own one row of the 32x32x32 grid, prefix sums are warp scans over ballots, the weight table
// (Application.cpp:346-370) is 33 entries in shared memory.
// Quirk kept: the quad is drawn at 0.8 scale (quad.vert:12) so only texels 3..28 in x and y are
// written, with tc = ((i+.5)/32-.5)/.8+.5; unwritten texels stay 0.
#include <cuda_fp16.h>

#include "raster.cuh"

namespace {
constexpr int N = VCT_WARP_DIM;

__device__ __forceinline__ bool quad_covered(int i) { const float c = (float)i + 0.5f; return c >= 3.2f && c <= 28.8f; }
__device__ __forceinline__ float quad_tc(int i) { return (((float)i + 0.5f) / (float)N - 0.5f) / 0.8f + 0.5f; }

__global__ void __launch_bounds__(1024) k_warpmap(const uint32_t* __restrict__ occ, const FrameConst* __restrict__ fcp, ushort4* __restrict__ warpmap,
                                                  ushort4* __restrict__ wlo, ushort4* __restrict__ whi, uint8_t* __restrict__ partials /* N^3 x 4 bytes scratch */) {
    const vct_frame_params& p = fcp->p;
    __shared__ float wl[N + 1], wh[N + 1];
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
    // inclusive prefix counts along each axis (Application.cpp:315-343); partials[cell] = (px, py, pz, occupied)
    for (int z = 0; z < N; ++z) {                                        // x-rows: warp = y, lane = x
        const int y = w, x = lane;
        const bool o = occ[(z * N + y) * N + x] > 0u;
        const unsigned m = __ballot_sync(0xffffffffu, o);
        uint8_t* q = partials + 4 * ((z * N + y) * N + x);
        q[0] = (uint8_t)__popc(m & (0xffffffffu >> (31 - lane))); q[3] = o ? 1 : 0;
    }
    for (int z = 0; z < N; ++z) {                                        // y-columns: warp = x, lane = y
        const int x = w, y = lane;
        const bool o = occ[(z * N + y) * N + x] > 0u;
        const unsigned m = __ballot_sync(0xffffffffu, o);
        partials[4 * ((z * N + y) * N + x) + 1] = (uint8_t)__popc(m & (0xffffffffu >> (31 - lane)));
    }
    for (int y = 0; y < N; ++y) {                                        // z-columns: warp = x, lane = z
        const int x = w, z = lane;
        const bool o = occ[(z * N + y) * N + x] > 0u;
        const unsigned m = __ballot_sync(0xffffffffu, o);
        partials[4 * ((z * N + y) * N + x) + 2] = (uint8_t)__popc(m & (0xffffffffu >> (31 - lane)));
    }
    __syncthreads();
    auto cell = [&](int x, int y, int z) { return partials + 4 * ((z * N + y) * N + x); };
    // clear outputs (unwritten texels are defined as 0)
    for (int i = tid; i < N * N * N; i += 1024) { warpmap[i] = make_ushort4(0, 0, 0, 0); wlo[i] = make_ushort4(0, 0, 0, 0); whi[i] = make_ushort4(0, 0, 0, 0); }
    __syncthreads();
    // pass 1: generateWarpmapWeights.frag:37-74 (two RGBA16F targets)
    if (p.use_warpmap_weights_texture)
        for (int i = tid; i < N * N * N; i += 1024) {
            const int x = i % N, y = (i / N) % N, z = i / (N * N);
            if (!quad_covered(x) || !quad_covered(y)) continue;
            const float tc[3] = {quad_tc(x), quad_tc(y), ((float)z + 0.5f) / (float)N};
            const int cx = (int)truncf(tc[0] * (float)N), cy = (int)truncf(tc[1] * (float)N), cz = (int)truncf(tc[2] * (float)N);
            const bool od = cell(cx, cy, cz)[3] != 0;
            const int tot[3] = {cell(N - 1, cy, cz)[0], cell(cx, N - 1, cz)[1], cell(cx, cy, N - 1)[2]};
            wlo[i] = make_ushort4(__half_as_ushort(__float2half_rn(wl[tot[0]])), __half_as_ushort(__float2half_rn(wl[tot[1]])),
                                  __half_as_ushort(__float2half_rn(wl[tot[2]])), __half_as_ushort(__float2half_rn(od ? 0.0f : 1.0f)));
            whi[i] = make_ushort4(__half_as_ushort(__float2half_rn(wh[tot[0]])), __half_as_ushort(__float2half_rn(wh[tot[1]])),
                                  __half_as_ushort(__float2half_rn(wh[tot[2]])), __half_as_ushort(__float2half_rn(od ? 1.0f : 0.0f)));
        }
    __syncthreads();
    // pass 2: generateWarpmap.frag:44-100
    for (int i = tid; i < N * N * N; i += 1024) {
        const int x = i % N, y = (i / N) % N, z = i / (N * N);
        if (!quad_covered(x) || !quad_covered(y)) continue;
        const float tc[3] = {quad_tc(x), quad_tc(y), ((float)z + 0.5f) / (float)N};
        int cid[3]; float fr[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float lt = tc[k] * (float)N, fl = truncf(lt); fr[k] = lt - fl; cid[k] = (int)fl; }
        const uint8_t* cc = cell(cid[0], cid[1], cid[2]);
        const bool od = cc[3] != 0;
        const int part[3] = {cc[0], cc[1], cc[2]};
        const int tot[3] = {cell(N - 1, cid[1], cid[2])[0], cell(cid[0], N - 1, cid[2])[1], cell(cid[0], cid[1], N - 1)[2]};
        float lo[3], hi[3];
        if (p.use_warpmap_weights_texture) {                              // NEAREST / CLAMP_TO_EDGE fetch at tc
            int t[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = min(max((int)floorf(tc[k] * (float)N), 0), N - 1);
            const ushort4 a = wlo[(t[2] * N + t[1]) * N + t[0]], b = whi[(t[2] * N + t[1]) * N + t[0]];
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
}
// float copy of the warp map for the cone tracer: the unorm16 -> float conversion (q / 65535, one IEEE division per channel) done once
// per texel here instead of 24 times per cone step there (config 4 at 4K: 68 ms -> see DESIGN.md); same values, bit for bit
__global__ void __launch_bounds__(256) k_warpmap_floats(const ushort4* __restrict__ warpmap, float4* __restrict__ wf) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM) return;
    const ushort4 q = warpmap[i];
    wf[i] = make_float4((float)q.x / 65535.0f, (float)q.y / 65535.0f, (float)q.z / 65535.0f, (float)q.w / 65535.0f);
}
}  // namespace

int vctk_warpmap_floats(vct_ctx* c) {
    const int n = VCT_WARP_DIM * VCT_WARP_DIM * VCT_WARP_DIM;
    k_warpmap_floats<<<(n + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const ushort4*>(c->d_warpmap), reinterpret_cast<float4*>(c->d_warpmap + 4 * (size_t)n));
    VCT_LAUNCH_CHECK(c, "k_warpmap_floats");
    return 0;
}

int vctk_warpmap(vct_ctx* c) {
    k_warpmap<<<1, 1024, 0, c->stream>>>(c->d_occ, c->d_fc, reinterpret_cast<ushort4*>(c->d_warpmap), reinterpret_cast<ushort4*>(c->d_wlo),
                                         reinterpret_cast<ushort4*>(c->d_whi), reinterpret_cast<uint8_t*>(c->d_warp_scratch));
    VCT_LAUNCH_CHECK(c, "k_warpmap");
    if (vctk_warpmap_floats(c)) return 1;
    return 0;
}
