// common.cuh — context, error handling and device helpers shared by every kernel file of libvct_b200.
//
// Exactness contract (DESIGN.md "Canonical GL semantics"): the translation units that decide WHICH voxel /
// texel / pixel a value lands in (raster.cu, voxelize.cu, volume_passes.cu, warpmap.cu) are compiled with
// -fmad=false so that every float operation is a separately rounded IEEE-754 op in source order; division and
// sqrt are the correctly rounded defaults (-prec-div/-prec-sqrt).  cone_trace.cu is tolerance-gated (PSNR) and
// is compiled with FMA contraction on.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vct_b200.h"

#define VCT_SM_COUNT 148
#define VCT_MAX_PEERS 8

// ------------------------------------------------------------------------------------------- small vectors
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
__host__ __device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__host__ __device__ __forceinline__ float length3(V3 a) { return sqrtf(dot3(a, a)); }
__host__ __device__ __forceinline__ V3 normalize3(V3 a) { float l = sqrtf(dot3(a, a)); return mk3(a.x / l, a.y / l, a.z / l); }
__host__ __device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__host__ __device__ __forceinline__ float maxsel(float a, float b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

struct Mat4 { float m[16]; };                       // column-major, as uploaded by the host
__host__ __device__ __forceinline__ V4 mul44(const Mat4& M, V4 v) {
    const float* m = M.m;
    V4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}

// ------------------------------------------------------------------------------------ unorm8 conversions
__device__ __forceinline__ uint32_t f2u_trunc(float v) { return __float2uint_rz(v); }       // saturating, NaN -> 0
__device__ __forceinline__ uint32_t unorm8(float v) {                                       // round(clamp(v,0,1)*255), RNE
    if (!(v > 0.0f)) return 0u;
    if (v > 1.0f) v = 1.0f;
    return __float2uint_rn(v * 255.0f);
}
__device__ __forceinline__ uint32_t pack_unorm(V4 c) { return unorm8(c.x) | unorm8(c.y) << 8 | unorm8(c.z) << 16 | unorm8(c.w) << 24; }
__device__ __forceinline__ V4 unpack_unorm(uint32_t w) {
    return mk4((float)(w & 255u) / 255.0f, (float)((w >> 8) & 255u) / 255.0f, (float)((w >> 16) & 255u) / 255.0f, (float)(w >> 24) / 255.0f);
}

// ------------------------------------------------------------------------------------------ device scene
struct DevTexture { const uint8_t* level[16]; int w, h, ch, levels; };
struct DevMaterial { int diffuse_tex, specular_tex, normal_tex, roughness_tex, metallic_tex, alpha_tex; float shininess; float diffuse[3]; };

// Per-frame constants every pass reads (uploaded once per frame into __constant__-like global struct).
// Step schedule of a cone (phong.frag:145-177): height h_i and lambda_i = lod_i + lodOffset depend only on the cone
// settings, so the host builds them once per frame.  lambda is non-decreasing: [0,n_point) NEAREST on level 0,
// [n_point,n_last) trilinear + mip-linear, [n_last,steps) last level only.
#define VCT_MAX_CONE_STEPS 64
struct ConeSchedule { float h[VCT_MAX_CONE_STEPS]; float lambda[VCT_MAX_CONE_STEPS]; int n_point, n_last, steps, pad; };

// How the z layers of the volume are dealt to the ranks (multi-GPU): layer z belongs to rank (z / T) mod N.  T = D / N gives one contiguous slab
// per rank, T = 16 interleaves stripes.  A rank owns `count` = D / (T N) stripes; its k-th stripe starts at z = (k N + rank) T.  One GPU: N = 1, T = D.
struct Stripes { int T, N, rank, count; };
__host__ __device__ __forceinline__ bool owns_z(const Stripes& s, int z) { return s.N <= 1 || (z / s.T) % s.N == s.rank; }
__host__ __device__ __forceinline__ int stripe_z(const Stripes& s, int k) { return (k * s.N + s.rank) * s.T; }
// does this rank own a layer of [zlo, zhi] (inclusive, already clamped to the volume)?
__host__ __device__ __forceinline__ bool owns_any_z(const Stripes& s, int zlo, int zhi) {
    if (s.N <= 1) return true;
    const int a = zlo / s.T, b = zhi / s.T;
    if (b - a + 1 >= s.N) return true;
    for (int k = a; k <= b; ++k) if (k % s.N == s.rank) return true;
    return false;
}
// index i of an array laid out over this rank's stripes back to back (`per` items per stripe) -> index in the whole volume's array
__host__ __device__ __forceinline__ size_t stripe_index(const Stripes& s, size_t i, size_t per) { const size_t k = i / per; return (k * s.N + s.rank) * per + (i - k * per); }

struct FrameConst {
    Mat4 projection, view, lp, lv, ls, ls_inverse, mvp_x, mvp_y, mvp_z;
    vct_frame_params p;          // scalar settings (matrices inside are unused on the device)
    int D, L, S, W, H;
    int n_lights;
    vct_light lights[8];
    Stripes st;                  // which z layers this rank owns
    ConeSchedule sched_diffuse, sched_specular;
};

struct Counters {                // device-resident, zeroed per frame
    unsigned total_fragments, unique_voxels, max_fragments_per_voxel;
    unsigned n_frag_slots;       // fragments emitted (== total_fragments within this slab)
    unsigned tile_queue_count;   // raster work queue (fine tiles)
    unsigned expand_count;       // raster work queue (bands of tile rows still to be enumerated)
    unsigned pixel_count;        // raster work queue (single pixels of tiny triangles)
    unsigned setup_count;        // big-triangle setups written by k_raster_bin
    unsigned overflow;           // set when a fixed-capacity buffer was too small (reset with the other per-frame counters)
    unsigned mip_ticket;         // k_mip_chain last-CTA detection (self-resetting)
    unsigned long_count, huge_count, huge_items, huge_tickets;   // deterministic voxeliser: queued long per-voxel lists (voxelize.cu)
    unsigned long long cone_steps;
    unsigned* overflow_host;     // the same flag in mapped pinned host memory: the frame entry points poll it without a device sync
};
#ifdef __CUDACC__
__device__ __forceinline__ void vct_flag_overflow(Counters* c) { c->overflow = 1u; if (c->overflow_host) *c->overflow_host = 1u; }
#endif

struct HostMesh {
    int actor; size_t n_vertices, n_tris; size_t vbase, tbase;   // offsets into the concatenated device arrays
    Mat4 model;
};

#define VCT_MAX_TEXTURES 256
#define VCT_MAX_MATERIALS 256

struct vct_ctx {
    vct_config cfg{};
    int D = 0, L = 0, S = 0, W = 0, H = 0;
    Stripes st{0, 1, 0, 1};          // z layers owned by this rank (make_volumes)
    cudaStream_t stream = nullptr;
    std::string error;
    unsigned long long launches = 0;

    // scene
    std::vector<HostMesh> meshes;
    std::vector<float> h_vertices; std::vector<uint32_t> h_indices; std::vector<int32_t> h_trimat; std::vector<int32_t> h_vactor;
    bool scene_dirty = true, tables_dirty = true;
    size_t n_vertices = 0, n_tris = 0;
    float* d_vertices = nullptr;      // n x 14
    int32_t* d_vactor = nullptr;
    uint32_t* d_indices = nullptr; int32_t* d_trimat = nullptr;
    Mat4* d_models = nullptr; float* d_nmats = nullptr; int n_actors = 0;
    float4 *d_wpos = nullptr, *d_wnrm = nullptr, *d_wT = nullptr, *d_wB = nullptr;   // per-vertex world-space attributes
    DevTexture h_tex[VCT_MAX_TEXTURES]{}; DevTexture* d_tex = nullptr; void* tex_alloc[VCT_MAX_TEXTURES]{}; int n_textures = 0;
    DevMaterial h_mat[VCT_MAX_MATERIALS]{}; DevMaterial* d_mat = nullptr; int n_materials = 0;
    vct_light h_lights[8]{}; int n_lights = 0;

    // volumes (linear, x fastest): one allocation per pyramid, level offsets in voxels
    uint32_t *d_color = nullptr, *d_radiance = nullptr, *d_normal = nullptr, *d_scratch = nullptr;
    size_t level_off[VCT_MAX_LEVELS + 1]{};
    cudaMipmappedArray_t radiance_arr = nullptr, color_arr = nullptr;
    cudaTextureObject_t radiance_tex = 0, radiance_tex_point = 0, radiance_tex_last = 0, color_tex = 0, color_tex_point = 0, color_tex_last = 0;
    cudaSurfaceObject_t radiance_surf[VCT_MAX_LEVELS]{}, color_surf[VCT_MAX_LEVELS]{};
    uint8_t *d_pub_mask_radiance = nullptr, *d_pub_mask_color = nullptr;   // see k_mip_chain
    // Segment masks (sparse frames): one byte per 8 level-0 voxels of an x-row.  d_seg[seg_cur] is written by this frame's
    // voxeliser ("a fragment landed here"), d_seg[seg_cur ^ 1] is last frame's.  While seg_valid, every non-zero word of
    // voxelColor / voxelNormal / voxelRadiance level 0 lies in a segment flagged in the previous mask, so clear, transfer
    // and the mip chain touch only flagged segments (the volume is ~97 % empty).  Anything that writes a volume outside
    // vct_frame / vct_gi_passes invalidates the masks; the next frame then runs the dense kernels once.
    // sparse slab exchange (exchange.cu): local staging, the peers' staging mapped through cudaIpc, record counter
    void* d_xchg = nullptr; size_t xchg_region_bytes = 0; unsigned xchg_cap = 0; unsigned* d_xchg_count = nullptr;
    vct_peer peer[VCT_MAX_PEERS]{}; bool peer_ipc[VCT_MAX_PEERS]{}; int peers_attached = 0; void* d_xchg_dst = nullptr;
    int mip_pushed_upto = 0;     // sharded frame: the mip chain of this frame stored levels 1..mip_pushed_upto into the peers' pyramids itself
    void* d_dbg = nullptr;                                         // vct_debug_voxels: depth | order buffer, voxel list, big-triangle queue (allocated on first use)
    uint32_t* d_trace_tiles = nullptr; int n_trace_tiles = 0;      // own 64x64 screen tiles (x0 | y0 << 16) of the sharded cone trace
    std::vector<vct_ctx*> group;                                    // single-process multi-GPU: the other ranks' contexts (this one is rank 0)
    void* group_state = nullptr; bool in_fan = false;               // worker threads of the group (api.cu); true while a call is being fanned out
    bool last_frame_sparse = false;
    uint8_t* d_seg[2] = {nullptr, nullptr}; int seg_cur = 0, seg_key = -1; bool seg_valid = false, seg_disabled = false, sparse_off = false;
    // warp
    uint32_t* d_occ = nullptr; uint16_t *d_warpmap = nullptr, *d_wlo = nullptr, *d_whi = nullptr;
    // shadow map / visibility / image
    float* d_shadow = nullptr; unsigned long long* d_vis = nullptr; uint32_t* d_image = nullptr;
    // two shadow maps: sharded frames rasterise a band of shadow rows per rank and store it into every peer's map; alternating maps
    // keep a peer that is already on the next frame from overwriting the one this rank still reads (d_shadow = the current one)
    float* d_shadow_base = nullptr; int shadow_parity = 0;
    uint32_t* d_inject_list = nullptr;                              // k_inject_cull: [0] = count, [1..] = active 64x16 shadow-map blocks
    unsigned char inject_key[256]{}; bool inject_list_valid = false; unsigned shadow_gen = 0, inject_gen = 0;   // what the list was built from
    void* d_shadow_mm = nullptr; bool shadow_mm_valid = false;   // (S/4)^2 x float2: min / max filtered depth per 4x4 texel block (k_shadow_minmax)
    // voxel fragments (per-voxel linked lists of the deterministic running average)
    size_t frag_cap = 0; void* d_frags = nullptr; uint8_t* d_displaced = nullptr;
    void* d_long_queue = nullptr; size_t long_cap = 65536; void* d_huge_items = nullptr; void* d_huge_aux = nullptr;   // long per-voxel lists: queue (+ 8 huge entries behind it), scan buffer
    uint32_t* d_warp_scratch = nullptr;
    // raster work queue: 8-byte tile items + one setup record per queued (sub-)triangle
    void* d_tile_queue = nullptr; size_t tile_queue_cap = 0; void* d_expand_queue = nullptr; size_t expand_cap = 0;
    void* d_pixel_queue = nullptr; size_t pixel_cap = 0;
    void* d_setup = nullptr; size_t setup_cap = 0;
    // per-frame constants + counters
    FrameConst* d_fc = nullptr; FrameConst h_fc{};
    // One host->device copy per pass call: [FrameConst | Mat4 models[n_actors] | float nmats[9 n_actors]] is staged in a
    // ring of pinned slots (a slot is reused only after the copy that read it has completed) and lands in one device blob.
    void* d_frame_blob = nullptr; size_t frame_blob_bytes = 0;
    unsigned char* h_stage = nullptr; cudaEvent_t stage_ev[4]{}; unsigned stage_next = 0;
    Counters* d_counters = nullptr; Counters h_counters{};
    unsigned* h_overflow = nullptr;  // mapped pinned word written by vct_flag_overflow
    // timing
    cudaEvent_t ev[32]{}; vct_timings timings{};
    int profiling = 1;               // 0 none, 1 pass-level events (reference GLTimer semantics), 2 + one event per kernel
    bool own_stream = true;
    // The image is double-buffered: d_image holds two images and consecutive cone traces alternate between them (image_parity = the
    // half the LAST trace wrote; all ranks of a sharded frame flip together).  An asynchronous read-back (vct_read_image_async) of
    // frame k runs on its own stream while frame k + 1 is rendered into the other half; only the trace — and, sharded, the exchange
    // that lets the peers store pixels into rank 0 — of frame k + 2 waits for it, i.e. practically never.
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_image_ready = nullptr, ev_copy_done[2] = {nullptr, nullptr}; bool copy_pending[2] = {false, false};
    int image_parity = 0;
    uint32_t* image_of(int parity) const;
    std::vector<cudaEvent_t> prof_pool; size_t prof_used = 0;
    std::vector<std::pair<const char*, cudaEvent_t>> prof_marks;
};

extern thread_local std::string g_create_error;

#define VCT_CHECK(ctx, call)                                                                    \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char b__[512];                                                                      \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            (ctx)->error = b__;                                                                 \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

// Per-kernel profiling (vct_set_profiling level 2): one event is recorded after every launch; the time between
// consecutive marks is attributed to the kernel named by the later mark (launches on one stream run back to back).
static inline void vct_prof_mark(vct_ctx* c, const char* name) {
    if (c->profiling < 2) return;
    if (c->prof_used == c->prof_pool.size()) { cudaEvent_t e = nullptr; if (cudaEventCreate(&e) != cudaSuccess) return; c->prof_pool.push_back(e); }
    cudaEvent_t e = c->prof_pool[c->prof_used++];
    cudaEventRecord(e, c->stream);
    c->prof_marks.push_back(std::make_pair(name, e));
}
static inline void vct_prof_begin(vct_ctx* c) { c->prof_used = 0; c->prof_marks.clear(); vct_prof_mark(c, "<begin>"); }

#define VCT_LAUNCH_CHECK(ctx, name)                                                             \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            char b__[512];                                                                      \
            snprintf(b__, sizeof b__, "%s:%d launch of %s -> %s", __FILE__, __LINE__, name, cudaGetErrorString(e__)); \
            (ctx)->error = b__;                                                                 \
            return 1;                                                                           \
        }                                                                                       \
        vct_prof_mark((ctx), name);                                                             \
    } while (0)

static inline int level_dim(int D, int l) { int d = D >> l; return d < 1 ? 1 : d; }

// pass entry points implemented across the .cu files (all enqueue on ctx->stream)
int vctk_transform_vertices(vct_ctx*);
int vctk_clear_voxels(vct_ctx*, bool reset_frame_counters = false);   // true: the same launch zeroes VoxelizeInfo, the raster queues and cone_steps
// fuse_transfer: the deterministic resolve also does transferVoxels for its voxels (sparse frame, temporal filter off); *transfer_done says whether it did
int vctk_voxelize(vct_ctx*, bool occupancy, bool counters_already_reset = false, bool fuse_transfer = false, bool* transfer_done = nullptr);
void vctk_fill_models(vct_ctx*, Mat4* models, float* nmats);
int vctk_transfer(vct_ctx*);
int vctk_clear_masked(vct_ctx*);        // sparse frames: clear flagged segments of all three volumes, reset the new mask and the frame counters
int vctk_transfer_masked(vct_ctx*);
int vctk_frame_begin_masked(vct_ctx*);   // vertex transform + masked clear in one launch
bool vctk_sparse_supported(const vct_ctx*);
int vctk_inject(vct_ctx*);
int vctk_fill_holes(vct_ctx*);
int vctk_shadow_minmax(vct_ctx*);
int vctk_mip(vct_ctx*, int which, int mode, int publish);
int vctk_mip_chains(vct_ctx*, int n, const int* which, const int* publish, int mode, bool masked = false, bool push_to_peers = false);
int vctk_xchg_mip_peers(vct_ctx*, uint32_t** peer_pyramid);   // sharded frame: wait until the peers are done with last frame's pyramid, then their bases (nullptr for this rank)
int vctk_publish(vct_ctx*, int which);
int vctk_publish_upper(vct_ctx*, int which);   // levels 1..L-1 only
int vctk_mip_top_sharded_level(const vct_ctx*);  // sharded frames: last level filtered from a rank's own stripes
int vctk_mip_tail(vct_ctx*, int which);          // sharded frames: the levels above it, from the exchanged level, on every rank
int vctk_xchg_setup(vct_ctx*);
void vctk_xchg_free(vct_ctx*);
bool vctk_xchg_ready(const vct_ctx*);          // multi-GPU with every peer attached: the frame entry points run the whole sharded frame
int vctk_xchg_frame(vct_ctx*, bool dense);
int vctk_xchg_image_sync(vct_ctx*);
int vctk_xchg_shadow(vct_ctx*, int row_lo, int row_hi);   // sharded shadow pass: own rows -> every peer's map, then wait for everybody's
int vctk_shadowmap(vct_ctx*);
int vctk_visibility(vct_ctx*);
int vctk_warpmap(vct_ctx*);
int vctk_warpmap_floats(vct_ctx*);       // float4 copy behind the unorm16 warp map (d_warpmap + 4 * 32^3 ushorts)
int vctk_cone_trace(vct_ctx*);
int vctk_debug_voxels(vct_ctx*);   // debug_voxels.cu
size_t vctk_image_rows(const vct_ctx*);
int vctk_set_voxel_opacity(vct_ctx*, float);
int vctk_temporal_radiance_filter(vct_ctx*, float);
int vctk_filter3d(vct_ctx*, int which, int src_level);
int vctk_normalize_voxels_f16(vct_ctx*, void*, void*, float);
