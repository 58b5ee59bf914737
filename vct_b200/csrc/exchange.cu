// exchange.cu — sparse slab exchange over NVLink peer memory (multi-GPU, SURVEY §8e "exchange only non-empty bricks").
//
// The reference is a single-GPU program; this is the one real exchange step of the z-slab sharded frame.  Level 0 of
// the traced pyramid is 8/9 of its bytes and ~97 % empty, so instead of all-gathering it densely every rank PUSHES the
// x-row segments of its own slab that hold — or held last frame — a fragment (the segment masks of the sparse frame,
// common.cuh) straight into a staging buffer in every peer's memory:
//
//   k_xchg_push     a warp scans 32 mask words of the own slab, compacts the flagged segments, reserves record slots with
//                   one atomic, and every lane stores one 16-byte half segment (+ the segment id) to ALL ranks' staging
//                   (remote stores over NVLink through cudaIpc-mapped pointers; stale segments travel as zeros)
//   k_xchg_finish   publishes the record count to every rank and re-arms the local counter
//   (the NCCL all-gather of the small levels >= 1 that follows on the same stream is the cross-GPU barrier)
//   k_xchg_unpack   every rank scatters the records of all senders into its linear level 0 (remote slabs), into the
//                   3D texture the cone tracer samples, and into the publish mask — a sparse publish instead of the
//                   dense linear -> array copy
//
// Staging layout on every rank: world_size regions (one per sender): [count | pad to 128 B][ids: cap x u32][data: cap x 32 B],
// cap = segments of one slab (worst case: every segment flagged).  A region is rewritten only after its owner has passed
// the image all-gather that ends the frame in which it was unpacked, so one buffer suffices.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr size_t kHdr = 128;

struct XchgPeers { unsigned char* base[VCT_MAX_PEERS]; int world, rank; size_t region_bytes; unsigned cap; };
__device__ __forceinline__ unsigned* region_count(unsigned char* base, size_t region_bytes, int sender) { return reinterpret_cast<unsigned*>(base + sender * region_bytes); }
__device__ __forceinline__ uint32_t* region_ids(unsigned char* base, size_t region_bytes, int sender) { return reinterpret_cast<uint32_t*>(base + sender * region_bytes + kHdr); }
__device__ __forceinline__ uint4* region_data(unsigned char* base, size_t region_bytes, int sender, unsigned cap) {
    return reinterpret_cast<uint4*>(base + sender * region_bytes + kHdr + (((size_t)cap * 4 + 127) & ~(size_t)127));
}

__global__ void __launch_bounds__(kThreads) k_xchg_push(const uint4* __restrict__ level0, const uint32_t* __restrict__ seg_now, const uint32_t* __restrict__ seg_before,
                                                        size_t word_lo, size_t word_hi, unsigned* __restrict__ counter, const __grid_constant__ XchgPeers peers) {
    __shared__ uint32_t s_list[kThreads / 32][128];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t n = word_hi - word_lo, n_round = (n + 31) & ~(size_t)31;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = word_lo + i;
        const uint32_t flags = i < n ? (__ldg(seg_now + t) | __ldg(seg_before + t)) : 0u;
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
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int q = lane; q < 2 * total; q += 32) {
            const uint32_t sid = s_list[w][q >> 1];
            const unsigned slot = base + (unsigned)(q >> 1);
            if (slot >= peers.cap) continue;                                 // cannot happen: cap = segments of the slab
            const uint4 v = __ldcs(level0 + 2 * (size_t)sid + (q & 1));
            for (int r = 0; r < peers.world; ++r) {
                unsigned char* b = peers.base[r];
                region_data(b, peers.region_bytes, peers.rank, peers.cap)[2 * (size_t)slot + (q & 1)] = v;
                if (!(q & 1)) region_ids(b, peers.region_bytes, peers.rank)[slot] = sid;
            }
        }
        __syncwarp();
    }
    __threadfence_system();
}
__global__ void k_xchg_finish(unsigned* __restrict__ counter, const __grid_constant__ XchgPeers peers) {
    const unsigned n = min(*counter, peers.cap);
    if ((int)threadIdx.x < peers.world) *region_count(peers.base[threadIdx.x], peers.region_bytes, peers.rank) = n;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
}
// grid.y = sender.  Records of the own slab are already in the linear level; remote ones are written there too, so that
// vct_read_volume and the next dense frame see one consistent volume.
__global__ void __launch_bounds__(kThreads) k_xchg_unpack(unsigned char* __restrict__ staging, size_t region_bytes, unsigned cap, int rank,
                                                          uint4* __restrict__ level0, cudaSurfaceObject_t surf, uint8_t* __restrict__ pub_mask, int D) {
    const int sender = blockIdx.y;
    const unsigned n = min(*region_count(staging, region_bytes, sender), cap);
    const uint32_t* ids = region_ids(staging, region_bytes, sender);
    const uint4* data = region_data(staging, region_bytes, sender, cap);
    const unsigned spr = (unsigned)D >> 3;                                   // segments per x-row
    for (unsigned q = blockIdx.x * kThreads + threadIdx.x; q < 2u * n; q += gridDim.x * kThreads) {
        const uint32_t sid = ids[q >> 1];
        const uint4 v = data[q];
        if (sender != rank) level0[2 * (size_t)sid + (q & 1)] = v;
        const unsigned row = sid / spr, sx = sid - row * spr;
        const int y = (int)(row % (unsigned)D), z = (int)(row / (unsigned)D);
        surf3Dwrite(v, surf, (int)(sx * 32u + (q & 1u) * 16u), y, z);
        const uint4 o = data[q ^ 1u];
        if (!(q & 1)) pub_mask[sid] = ((v.x | v.y | v.z | v.w | o.x | o.y | o.z | o.w) != 0u) ? 1 : 0;
    }
}

}  // namespace

size_t vctk_xchg_region_bytes(const vct_ctx* c) {
    const size_t cap = (size_t)c->D * c->D * (c->z_hi - c->z_lo) / 8;
    return kHdr + ((cap * 4 + 127) & ~(size_t)127) + cap * 32;
}
int vctk_xchg_setup(vct_ctx* c) {
    if (c->d_xchg) return 0;
    const int ws = c->cfg.world_size;
    if (ws < 2 || ws > VCT_MAX_PEERS) { c->error = "sparse exchange: world_size must be in [2, 8]"; return 1; }
    c->xchg_cap = (unsigned)((size_t)c->D * c->D * (c->z_hi - c->z_lo) / 8);
    c->xchg_region_bytes = vctk_xchg_region_bytes(c);
    VCT_CHECK(c, cudaMalloc(&c->d_xchg, c->xchg_region_bytes * ws));
    VCT_CHECK(c, cudaMemsetAsync(c->d_xchg, 0, c->xchg_region_bytes * ws, c->stream));
    VCT_CHECK(c, cudaMalloc(&c->d_xchg_count, 128));
    VCT_CHECK(c, cudaMemsetAsync(c->d_xchg_count, 0, 128, c->stream));
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < VCT_MAX_PEERS; ++r) c->peer_xchg[r] = nullptr;
    c->peer_xchg[c->cfg.rank] = c->d_xchg;
    return 0;
}
void vctk_xchg_free(vct_ctx* c) {
    for (int r = 0; r < VCT_MAX_PEERS; ++r) {
        if (c->peer_xchg[r] && r != c->cfg.rank) cudaIpcCloseMemHandle(c->peer_xchg[r]);
        c->peer_xchg[r] = nullptr;
    }
    cudaFree(c->d_xchg); cudaFree(c->d_xchg_count);
    c->d_xchg = nullptr; c->d_xchg_count = nullptr;
}
static int xchg_peers(vct_ctx* c, XchgPeers& p) {
    p.world = c->cfg.world_size; p.rank = c->cfg.rank; p.region_bytes = c->xchg_region_bytes; p.cap = c->xchg_cap;
    for (int r = 0; r < p.world; ++r) {
        if (!c->peer_xchg[r]) { c->error = "sparse exchange: a peer's staging buffer was not imported (vct_exchange_import)"; return 1; }
        p.base[r] = reinterpret_cast<unsigned char*>(c->peer_xchg[r]);
    }
    return 0;
}
int vctk_xchg_push(vct_ctx* c) {
    XchgPeers p{};
    if (xchg_peers(c, p)) return 1;
    const bool rad = c->h_fc.p.draw_radiance != 0;
    const uint4* level0 = reinterpret_cast<const uint4*>(rad ? c->d_radiance : c->d_color);
    // gi_body swapped the masks at its end: this frame's is seg_cur ^ 1, last frame's is seg_cur
    const size_t words_per_slice = (size_t)c->D * c->D / 32;
    const size_t lo = (size_t)c->z_lo * words_per_slice, hi = (size_t)c->z_hi * words_per_slice;
    const size_t blocks = std::min<size_t>((hi - lo + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_xchg_push<<<(unsigned)std::max<size_t>(blocks, 1), kThreads, 0, c->stream>>>(level0, reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur ^ 1]),
                                                                                  reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur]), lo, hi, c->d_xchg_count, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_push");
    k_xchg_finish<<<1, 32, 0, c->stream>>>(c->d_xchg_count, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_finish");
    return 0;
}
int vctk_xchg_unpack(vct_ctx* c) {
    const bool rad = c->h_fc.p.draw_radiance != 0;
    uint4* level0 = reinterpret_cast<uint4*>(rad ? c->d_radiance : c->d_color);
    const cudaSurfaceObject_t surf = rad ? c->radiance_surf[0] : c->color_surf[0];
    uint8_t* mask = rad ? c->d_pub_mask_radiance : c->d_pub_mask_color;
    if (!surf || !mask) { c->error = "sparse exchange: the traced pyramid has no texture array yet (run one dense frame first)"; return 1; }
    dim3 grid(VCT_SM_COUNT * 2, c->cfg.world_size);
    k_xchg_unpack<<<grid, kThreads, 0, c->stream>>>(reinterpret_cast<unsigned char*>(c->d_xchg), c->xchg_region_bytes, c->xchg_cap, c->cfg.rank, level0, surf, mask, c->D);
    VCT_LAUNCH_CHECK(c, "k_xchg_unpack");
    return 0;
}
