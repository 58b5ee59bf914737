// exchange.cu — the one exchange step of the z-sharded frame (a rank owns a slab or interleaved stripes of z layers: Stripes, common.cuh) (multi-GPU), entirely in kernels over NVLink peer
// memory: no NCCL collective and no host synchronisation on the frame's critical path, so the whole N-GPU step can be captured in a
// CUDA graph like the single-GPU step.
//
// The reference is a single-GPU program; sharding leaves one real exchange: after its voxel passes every rank owns its z layers of
// every level of the traced pyramid and the cone tracer of every rank samples all of it.
//   level 0      8/9 of the bytes and ~97 % empty: a rank PUSHES the x-row segments (8 voxels, 32 bytes) of its slab that hold — or
//                held last frame — a fragment (the segment masks of the sparse frame, common.cuh) as (id, data) records into a
//                staging region in every peer's memory (k_xchg_push: warp-compacted, one atomic per 128 segments, remote 16-byte
//                stores); on a dense frame (the first, or after the masks were invalidated) every segment of the slab travels
//   levels >= 1  small and dense per slab: copied straight into the peers' linear levels at the place they have everywhere
//                (k_xchg_push_upper)
//   flags        every rank's allocation starts with a control block the PEERS write: ready[s] = the frame number whose records sender s
//                has completed, consumed[p] = the last frame rank p has unpacked.  k_xchg_finish publishes count + ready after a
//                system-wide fence; k_xchg_unpack spins on ready[] of all senders before it touches a record; k_xchg_push spins on
//                consumed[] of all peers before it overwrites the single staging region (a wait that is long satisfied in steady
//                state: a whole frame lies between).  Frame numbers live in device memory and are advanced by k_xchg_ack, so a
//                replayed graph counts on correctly.  Every wait is for an event on ANOTHER GPU that depends on nothing here: the
//                chain push(k) -> unpack(k) -> ack(k) -> push(k+1) cannot deadlock.
//   unpack       records of the remote slabs are scattered into the linear level 0, the 3D texture the cone tracer samples and the
//                publish mask (a sparse publish); the own slab was published by the mip chain; levels >= 1 follow by one launch
//   image        a rank's cone trace stores its pixels into rank 0's image as well (cone_trace.cuh); k_xchg_image_done / _wait order
//                rank 0's readers behind every rank's pixels.
// Peer memory is mapped by cudaIpc (one process per GPU: vct_exchange_export / _import) or is plain peer-accessible device memory
// (one process driving all GPUs: vct_exchange_attach).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr size_t kCtrlBytes = 256;       // XchgControl, padded
constexpr size_t kHdr = 128;

struct XchgControl {                     // first bytes of every rank's staging allocation; WRITTEN BY THE PEERS
    unsigned ready[VCT_MAX_PEERS];       // ready[s]: sender s has pushed all its records of frame `ready[s]`
    unsigned count[VCT_MAX_PEERS];       // count[s]: how many
    unsigned consumed[VCT_MAX_PEERS];    // consumed[p]: rank p has unpacked frame `consumed[p]` (its staging may be overwritten)
    unsigned image_done[VCT_MAX_PEERS];  // rank 0: rank p's pixels of frame `image_done[p]` are in this rank's image
    unsigned shadow_ready[VCT_MAX_PEERS];// shadow_ready[s]: rank s has stored its band of shadow map number `shadow_ready[s]` here
};
static_assert(sizeof(XchgControl) <= kCtrlBytes, "control block");

struct XchgPeers { unsigned char* base[VCT_MAX_PEERS]; uint32_t* pyramid[VCT_MAX_PEERS]; int world, rank; size_t region_bytes; unsigned cap; };
__device__ __forceinline__ XchgControl* ctrl(unsigned char* base) { return reinterpret_cast<XchgControl*>(base); }
__device__ __forceinline__ uint32_t* region_ids(unsigned char* base, size_t region_bytes, int sender) { return reinterpret_cast<uint32_t*>(base + kCtrlBytes + sender * region_bytes + kHdr); }
__device__ __forceinline__ uint4* region_data(unsigned char* base, size_t region_bytes, int sender, unsigned cap) {
    return reinterpret_cast<uint4*>(base + kCtrlBytes + sender * region_bytes + kHdr + (((size_t)cap * 4 + 127) & ~(size_t)127));
}
// spin until flags[r] >= want for every r != self (one thread per CTA calls this; the CTA then passes a barrier)
__device__ __forceinline__ void wait_flags(const unsigned* flags, int world, int self, unsigned want) {
    for (int r = 0; r < world; ++r) {
        if (r == self) continue;
        while ((int)(*(const volatile unsigned*)(flags + r) - want) < 0) __nanosleep(64);
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(kThreads) k_xchg_push(const uint4* __restrict__ level0, const uint32_t* __restrict__ seg_now, const uint32_t* __restrict__ seg_before, int dense,
                                                        Stripes st, size_t w_stripe, size_t n, unsigned* __restrict__ counter, const unsigned* __restrict__ seq,
                                                        const __grid_constant__ XchgPeers peers) {
    __shared__ uint32_t s_list[kThreads / 32][128];
    if (threadIdx.x == 0) wait_flags(ctrl(peers.base[peers.rank])->consumed, peers.world, peers.rank, *seq);   // the peers are done with last frame's records
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t n_round = (n + 31) & ~(size_t)31;                          // n mask words over this rank's stripes, back to back
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i < n ? stripe_index(st, i, w_stripe) : 0;
        const uint32_t flags = i < n ? (dense ? 0x01010101u : (__ldg(seg_now + t) | __ldg(seg_before + t))) : 0u;
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
                if (r == peers.rank) continue;                               // the own slab is published by the mip chain
                unsigned char* b = peers.base[r];
                region_data(b, peers.region_bytes, peers.rank, peers.cap)[2 * (size_t)slot + (q & 1)] = v;
                if (!(q & 1)) region_ids(b, peers.region_bytes, peers.rank)[slot] = sid;
            }
        }
        __syncwarp();
    }
    __threadfence_system();
}
// levels >= 1 of the own slab, dense, straight into every peer's linear pyramid (same offsets on every rank)
// one run per (level, stripe): `first` = running element count, `off` = word offset in the pyramid; runs of one level have one length
constexpr int kMaxRuns = 160;
struct UpperLevels { unsigned first[kMaxRuns + 1]; unsigned off[kMaxRuns]; int n; };
__global__ void __launch_bounds__(kThreads) k_xchg_push_upper(const uint32_t* __restrict__ pyramid, const __grid_constant__ UpperLevels lv, const __grid_constant__ XchgPeers peers) {
    const unsigned long long total = lv.first[lv.n];
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        int k = 0, hi = lv.n;                                                 // binary search: first[k] <= i < first[k + 1]
        while (hi - k > 1) { const int m = (k + hi) >> 1; if (i >= lv.first[m]) k = m; else hi = m; }
        const unsigned long long w = lv.off[k] + (i - lv.first[k]);           // word offset of this rank's chunk element in the pyramid
        const uint32_t v = __ldcg(pyramid + w);
        for (int r = 0; r < peers.world; ++r) if (r != peers.rank) peers.pyramid[r][w] = v;
    }
    __threadfence_system();
}
__global__ void k_xchg_finish(unsigned* __restrict__ counter, const unsigned* __restrict__ seq, const __grid_constant__ XchgPeers peers) {
    const unsigned n = min(*counter, peers.cap), s = *seq + 1u;
    __threadfence_system();
    if ((int)threadIdx.x < peers.world && (int)threadIdx.x != peers.rank) {
        XchgControl* c = ctrl(peers.base[threadIdx.x]);
        c->count[peers.rank] = n;
        __threadfence_system();
        *(volatile unsigned*)&c->ready[peers.rank] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
}
// grid.y = sender (remote ranks only do work).  Remote records land in the linear level 0 too, so that vct_read_volume and the next
// dense frame see one consistent volume.
__global__ void __launch_bounds__(kThreads) k_xchg_unpack(unsigned char* __restrict__ staging, size_t region_bytes, unsigned cap, int rank, int world, const unsigned* __restrict__ seq,
                                                          uint4* __restrict__ level0, cudaSurfaceObject_t surf, uint8_t* __restrict__ pub_mask, int D) {
    const int sender = blockIdx.y;
    if (sender == rank) return;
    if (threadIdx.x == 0) wait_flags(ctrl(staging)->ready, world, rank, *seq + 1u);
    __syncthreads();
    const unsigned n = min(*(const volatile unsigned*)&ctrl(staging)->count[sender], cap);
    const uint32_t* ids = region_ids(staging, region_bytes, sender);
    const uint4* data = region_data(staging, region_bytes, sender, cap);
    const unsigned spr = (unsigned)D >> 3;                                   // segments per x-row
    for (unsigned q = blockIdx.x * kThreads + threadIdx.x; q < 2u * n; q += gridDim.x * kThreads) {
        const uint32_t sid = __ldcg(ids + (q >> 1));
        const uint4 v = __ldcg(data + q);
        level0[2 * (size_t)sid + (q & 1)] = v;
        const unsigned row = sid / spr, sx = sid - row * spr;
        const int y = (int)(row % (unsigned)D), z = (int)(row / (unsigned)D);
        surf3Dwrite(v, surf, (int)(sx * 32u + (q & 1u) * 16u), y, z);
        const uint4 o = __ldcg(data + (q ^ 1u));
        if (!(q & 1)) pub_mask[sid] = ((v.x | v.y | v.z | v.w | o.x | o.y | o.z | o.w) != 0u) ? 1 : 0;
    }
}
__global__ void k_xchg_ack(unsigned* __restrict__ seq, const __grid_constant__ XchgPeers peers) {
    const unsigned s = *seq + 1u;
    __threadfence_system();
    if ((int)threadIdx.x < peers.world && (int)threadIdx.x != peers.rank) *(volatile unsigned*)&ctrl(peers.base[threadIdx.x])->consumed[peers.rank] = s;
    __syncthreads();
    if (threadIdx.x == 0) *seq = s;
}
// before a kernel that stores into the peers' pyramids directly (the mip chain): every peer has acknowledged last frame, i.e. is done
// reading its pyramid (publish of the upper levels, tail levels)
__global__ void k_xchg_wait_consumed(const unsigned* __restrict__ seq, const __grid_constant__ XchgPeers peers) {
    wait_flags(ctrl(peers.base[peers.rank])->consumed, peers.world, peers.rank, *seq);
}
// after the cone trace: this rank's pixels of frame *seq are in rank 0's image
__global__ void k_xchg_image_done(const unsigned* __restrict__ seq, const __grid_constant__ XchgPeers peers) {
    __threadfence_system();
    *(volatile unsigned*)&ctrl(peers.base[0])->image_done[peers.rank] = *seq;
}
__global__ void k_xchg_image_wait(const unsigned* __restrict__ seq, const __grid_constant__ XchgPeers peers) {
    wait_flags(ctrl(peers.base[0])->image_done, peers.world, 0, *seq);
}

// sharded shadow pass: this rank's band of rows -> the same rows of every peer's map; then the flag; then wait for all bands
__global__ void __launch_bounds__(kThreads) k_xchg_push_shadow(const uint4* __restrict__ src, size_t n16, size_t off16, const __grid_constant__ XchgPeers peers, uint4* const* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(src + off16 + i);
        for (int r = 0; r < peers.world; ++r) if (r != peers.rank) dst[r][off16 + i] = v;
    }
    __threadfence_system();
}
__global__ void k_xchg_shadow_sync(unsigned* __restrict__ sseq, const __grid_constant__ XchgPeers peers) {
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) s_seq = *sseq + 1u;
    __syncthreads();
    __threadfence_system();
    if ((int)threadIdx.x < peers.world && (int)threadIdx.x != peers.rank) *(volatile unsigned*)&ctrl(peers.base[threadIdx.x])->shadow_ready[peers.rank] = s_seq;
    __syncthreads();
    if (threadIdx.x == 0) { wait_flags(ctrl(peers.base[peers.rank])->shadow_ready, peers.world, peers.rank, s_seq); *sseq = s_seq; }
}

}  // namespace

size_t vctk_xchg_region_bytes(const vct_ctx* c) {
    const size_t cap = (size_t)c->D * c->D * c->st.T * c->st.count / 8;
    return kHdr + ((cap * 4 + 127) & ~(size_t)127) + cap * 32;
}
int vctk_xchg_setup(vct_ctx* c) {
    if (c->d_xchg) return 0;
    const int ws = c->cfg.world_size;
    if (ws < 2 || ws > VCT_MAX_PEERS) { c->error = "sparse exchange: world_size must be in [2, 8]"; return 1; }
    c->xchg_cap = (unsigned)((size_t)c->D * c->D * c->st.T * c->st.count / 8);
    c->xchg_region_bytes = vctk_xchg_region_bytes(c);
    const size_t bytes = kCtrlBytes + c->xchg_region_bytes * ws;
    VCT_CHECK(c, cudaMalloc(&c->d_xchg, bytes));
    VCT_CHECK(c, cudaMemsetAsync(c->d_xchg, 0, bytes, c->stream));
    VCT_CHECK(c, cudaMalloc(&c->d_xchg_count, 128));                        // [0] record counter, [16] frame sequence number
    VCT_CHECK(c, cudaMemsetAsync(c->d_xchg_count, 0, 128, c->stream));
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < VCT_MAX_PEERS; ++r) c->peer[r] = vct_peer{};
    vct_peer& me = c->peer[c->cfg.rank];
    me.staging = c->d_xchg; me.radiance = c->d_radiance; me.color = c->d_color; me.image = c->d_image; me.shadow = c->d_shadow_base;
    c->peers_attached = 1;
    return 0;
}
void vctk_xchg_free(vct_ctx* c) {
    for (int r = 0; r < VCT_MAX_PEERS; ++r) {
        if (r != c->cfg.rank && c->peer_ipc[r]) {
            for (void* p : {c->peer[r].staging, c->peer[r].radiance, c->peer[r].color, c->peer[r].image, c->peer[r].shadow}) if (p) cudaIpcCloseMemHandle(p);
        }
        c->peer[r] = vct_peer{}; c->peer_ipc[r] = false;
    }
    cudaFree(c->d_xchg); cudaFree(c->d_xchg_count); cudaFree(c->d_xchg_dst);
    c->d_xchg = nullptr; c->d_xchg_count = nullptr; c->d_xchg_dst = nullptr; c->peers_attached = 0;
}
bool vctk_xchg_ready(const vct_ctx* c) { return c->cfg.world_size > 1 && c->d_xchg && c->peers_attached == c->cfg.world_size; }
static int xchg_peers(vct_ctx* c, XchgPeers& p) {
    const bool rad = c->h_fc.p.draw_radiance != 0;
    p.world = c->cfg.world_size; p.rank = c->cfg.rank; p.region_bytes = c->xchg_region_bytes; p.cap = c->xchg_cap;
    for (int r = 0; r < p.world; ++r) {
        if (!c->peer[r].staging) { c->error = "slab exchange: a peer's buffers were not attached (vct_exchange_import / vct_exchange_attach)"; return 1; }
        p.base[r] = reinterpret_cast<unsigned char*>(c->peer[r].staging);
        p.pyramid[r] = reinterpret_cast<uint32_t*>(rad ? c->peer[r].radiance : c->peer[r].color);
    }
    return 0;
}
int vctk_xchg_mip_peers(vct_ctx* c, uint32_t** peer_pyramid) {
    XchgPeers p{};
    if (xchg_peers(c, p)) return 1;
    k_xchg_wait_consumed<<<1, 1, 0, c->stream>>>(c->d_xchg_count + 16, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_wait_consumed");
    for (int r = 0; r < VCT_MAX_PEERS; ++r) peer_pyramid[r] = r < p.world && r != p.rank ? p.pyramid[r] : nullptr;
    return 0;
}
// the exchange of one frame: every rank ends up with the whole traced pyramid in its 3D texture
int vctk_xchg_frame(vct_ctx* c, bool dense) {
    XchgPeers p{};
    if (xchg_peers(c, p)) return 1;
    const bool rad = c->h_fc.p.draw_radiance != 0;
    uint32_t* pyr = rad ? c->d_radiance : c->d_color;
    const cudaSurfaceObject_t surf = rad ? c->radiance_surf[0] : c->color_surf[0];
    uint8_t* mask = rad ? c->d_pub_mask_radiance : c->d_pub_mask_color;
    if (!surf || !mask) { c->error = "slab exchange: the traced pyramid has no texture array"; return 1; }
    unsigned* counter = c->d_xchg_count; unsigned* seq = c->d_xchg_count + 16;
    const int par = c->image_parity ^ 1;                     // the image half this frame's traces write (on every rank)
    if (c->copy_pending[par]) {                              // still being read back on rank 0 (two frames old): the peers may not store into it yet
        VCT_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done[par], 0));
        c->copy_pending[par] = false;
    }
    // gi_body swapped the masks at its end: this frame's is seg_cur ^ 1, last frame's is seg_cur
    const size_t w_stripe = (size_t)c->D * c->D / 32 * c->st.T, n_words = w_stripe * c->st.count;
    const size_t blocks = std::min<size_t>((n_words + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 16);
    k_xchg_push<<<(unsigned)std::max<size_t>(blocks, 1), kThreads, 0, c->stream>>>(reinterpret_cast<const uint4*>(pyr), reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur ^ 1]),
                                                                                  reinterpret_cast<const uint32_t*>(c->d_seg[c->seg_cur]), dense ? 1 : 0, c->st, w_stripe, n_words, counter, seq, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_push");
    const int top = vctk_mip_top_sharded_level(c);          // levels 1..top: this rank's stripes, dense; the levels above follow from level `top` on every rank
    const int first_upper = c->mip_pushed_upto + 1;         // levels below it went to the peers from inside the mip chain
    c->mip_pushed_upto = 0;
    if (top >= first_upper) {
        UpperLevels lv{};
        unsigned long long n = 0;
        for (int l = first_upper; l <= top; ++l) {
            const unsigned long long d = level_dim(c->D, l), run = d * d * (unsigned long long)(c->st.T >> l);
            for (int k = 0; k < c->st.count; ++k) {
                if (lv.n >= kMaxRuns || c->level_off[l] + d * d * d > 0xFFFFFFFFull) { c->error = "slab exchange: too many (level, stripe) runs — use a larger slab_stripe"; return 1; }
                lv.first[lv.n] = (unsigned)n; lv.off[lv.n] = (unsigned)(c->level_off[l] + d * d * (unsigned long long)(stripe_z(c->st, k) >> l)); lv.n++; n += run;
            }
        }
        lv.first[lv.n] = (unsigned)n;
        k_xchg_push_upper<<<(unsigned)std::min<unsigned long long>((n + kThreads - 1) / kThreads, (unsigned long long)VCT_SM_COUNT * 8), kThreads, 0, c->stream>>>(pyr, lv, p);
        VCT_LAUNCH_CHECK(c, "k_xchg_push_upper");
    }
    k_xchg_finish<<<1, 32, 0, c->stream>>>(counter, seq, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_finish");
    dim3 grid(VCT_SM_COUNT, c->cfg.world_size);
    k_xchg_unpack<<<grid, kThreads, 0, c->stream>>>(reinterpret_cast<unsigned char*>(c->d_xchg), c->xchg_region_bytes, c->xchg_cap, c->cfg.rank, c->cfg.world_size, seq,
                                                    reinterpret_cast<uint4*>(pyr), surf, mask, c->D);
    VCT_LAUNCH_CHECK(c, "k_xchg_unpack");
    if (top + 1 < c->L && vctk_mip_tail(c, rad ? VCT_VOL_RADIANCE : VCT_VOL_COLOR)) return 1;
    if (vctk_publish_upper(c, rad ? VCT_VOL_RADIANCE : VCT_VOL_COLOR)) return 1;
    k_xchg_ack<<<1, 32, 0, c->stream>>>(seq, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_ack");
    return 0;
}
// Sharded shadow pass (raster_passes.cu): the caller has rasterised rows [row_lo, row_hi) of the CURRENT shadow map (c->d_shadow, the
// half selected by shadow_parity); store them into the same half of every peer's map and wait until every peer's band has arrived.
int vctk_xchg_shadow(vct_ctx* c, int row_lo, int row_hi) {
    XchgPeers p{};
    if (xchg_peers(c, p)) return 1;
    const size_t half = (size_t)c->S * c->S;                // floats per map
    if (!c->d_xchg_dst) {                                    // device table of the peers' map bases (both halves follow each other)
        VCT_CHECK(c, cudaMalloc(&c->d_xchg_dst, sizeof(void*) * VCT_MAX_PEERS));
        void* h[VCT_MAX_PEERS] = {};
        for (int r = 0; r < p.world; ++r) h[r] = c->peer[r].shadow;
        VCT_CHECK(c, cudaMemcpyAsync(c->d_xchg_dst, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    }
    unsigned* sseq = c->d_xchg_count + 24;
    const size_t off16 = ((size_t)c->shadow_parity * half + (size_t)row_lo * c->S) / 4, n16 = (size_t)(row_hi - row_lo) * c->S / 4;
    if (n16) {
        k_xchg_push_shadow<<<(unsigned)std::min<size_t>((n16 + kThreads - 1) / kThreads, (size_t)VCT_SM_COUNT * 8), kThreads, 0, c->stream>>>(
            reinterpret_cast<const uint4*>(c->d_shadow_base), n16, off16, p, reinterpret_cast<uint4* const*>(c->d_xchg_dst));
        VCT_LAUNCH_CHECK(c, "k_xchg_push_shadow");
    }
    k_xchg_shadow_sync<<<1, 32, 0, c->stream>>>(sseq, p);
    VCT_LAUNCH_CHECK(c, "k_xchg_shadow_sync");
    return 0;
}
// after the cone trace of a sharded frame: rank 0's stream continues only when every rank's pixels are in its image
int vctk_xchg_image_sync(vct_ctx* c) {
    XchgPeers p{};
    if (xchg_peers(c, p)) return 1;
    const unsigned* seq = c->d_xchg_count + 16;
    if (c->cfg.rank != 0) { k_xchg_image_done<<<1, 1, 0, c->stream>>>(seq, p); VCT_LAUNCH_CHECK(c, "k_xchg_image_done"); }
    else { k_xchg_image_wait<<<1, 1, 0, c->stream>>>(seq, p); VCT_LAUNCH_CHECK(c, "k_xchg_image_wait"); }
    return 0;
}
