// api.cu — the C ABI of include/vct_b200.h: resource ownership (reference `VCT` + Application::init), scene
// upload (Mesh VAO/EBO/texture creation), the per-pass entry points that replace the GL dispatch blocks of
// Application::render (src/Application.cpp:196-1085) and the whole-frame graph in the reference's pass order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>

#include <stdlib.h>

#include <cstddef>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

thread_local std::string g_create_error;
uint32_t* vct_ctx::image_of(int parity) const { return d_image + (size_t)parity * vctk_image_rows(this) * (size_t)W; }
size_t vctk_tile_setup_bytes();
size_t vctk_vox_setup_bytes();
size_t vctk_frag_bytes();
size_t vctk_huge_aux_bytes();
size_t vctk_image_rows(const vct_ctx*);

namespace {

int fail(vct_ctx* c, const char* msg) { c->error = msg; return 1; }

// NVTX ranges around the launches of every pass, named like the reference's GLBufferedTimers (src/Application.h:192: voxelize,
// shadowmap, radiance, mipmap, render, total) plus this build's extra passes; free when no tool is attached (header-only NVTX 3).
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };
struct NvtxPhases { bool open = false; void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; } ~NvtxPhases() { if (open) nvtxRangePop(); } };

void free_volumes(vct_ctx* c) {
    for (int l = 0; l < VCT_MAX_LEVELS; ++l) {
        if (c->radiance_surf[l]) cudaDestroySurfaceObject(c->radiance_surf[l]);
        if (c->color_surf[l]) cudaDestroySurfaceObject(c->color_surf[l]);
        c->radiance_surf[l] = c->color_surf[l] = 0;
    }
    for (cudaTextureObject_t* t : {&c->radiance_tex, &c->radiance_tex_point, &c->radiance_tex_last, &c->color_tex, &c->color_tex_point, &c->color_tex_last}) { if (*t) cudaDestroyTextureObject(*t); *t = 0; }
    if (c->radiance_arr) cudaFreeMipmappedArray(c->radiance_arr);
    if (c->color_arr) cudaFreeMipmappedArray(c->color_arr);
    c->radiance_arr = c->color_arr = nullptr;
    for (uint32_t** p : {&c->d_color, &c->d_radiance, &c->d_normal, &c->d_scratch}) { cudaFree(*p); *p = nullptr; }
    for (uint8_t** p : {&c->d_pub_mask_radiance, &c->d_pub_mask_color, &c->d_seg[0], &c->d_seg[1]}) { cudaFree(*p); *p = nullptr; }
    c->seg_valid = false; c->seg_disabled = false; c->seg_cur = 0;
}

int make_pyramid_texture(vct_ctx* c, cudaMipmappedArray_t* arr, cudaTextureObject_t* lin, cudaTextureObject_t* pt, cudaTextureObject_t* last, cudaSurfaceObject_t* surf, uint8_t** mask) {
    // array contents are undefined until the first publish: mask = everything may be non-zero
    VCT_CHECK(c, cudaMalloc(mask, (size_t)c->D * c->D * c->D / 8));
    VCT_CHECK(c, cudaMemsetAsync(*mask, 0xFF, (size_t)c->D * c->D * c->D / 8, c->stream));
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    VCT_CHECK(c, cudaMallocMipmappedArray(arr, &fd, make_cudaExtent(c->D, c->D, c->D), c->L, cudaArraySurfaceLoadStore));
    for (int l = 0; l < c->L; ++l) {
        cudaArray_t lev; VCT_CHECK(c, cudaGetMipmappedArrayLevel(&lev, *arr, l));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = lev;
        VCT_CHECK(c, cudaCreateSurfaceObject(&surf[l], &rd));
    }
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeMipmappedArray; rd.res.mipmap.mipmap = *arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;      // CLAMP_TO_BORDER, border 0
    td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = (float)(c->L - 1);
    td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModeLinear;        // LINEAR_MIPMAP_LINEAR
    VCT_CHECK(c, cudaCreateTextureObject(lin, &rd, &td, nullptr));
    td.filterMode = cudaFilterModePoint; td.mipmapFilterMode = cudaFilterModePoint;          // mag NEAREST
    VCT_CHECK(c, cudaCreateTextureObject(pt, &rd, &td, nullptr));
    td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModePoint;         // one level, trilinear: lambda >= L-1
    td.minMipmapLevelClamp = td.maxMipmapLevelClamp = (float)(c->L - 1);
    VCT_CHECK(c, cudaCreateTextureObject(last, &rd, &td, nullptr));
    return 0;
}

// `VCT::make` (reference src/Application.h:143-148): voxelColor (L levels), voxelNormal (1), voxelRadiance (L)
int make_volumes(vct_ctx* c) {
    size_t off = 0;
    for (int l = 0; l <= c->L; ++l) {
        c->level_off[l] = off;
        if (l < c->L) { const size_t d = level_dim(c->D, l); off += (d * d * d + 63) & ~(size_t)63; }
    }
    const size_t n0 = (size_t)c->D * c->D * c->D;
    VCT_CHECK(c, cudaMalloc(&c->d_color, off * 4)); VCT_CHECK(c, cudaMalloc(&c->d_radiance, off * 4)); VCT_CHECK(c, cudaMalloc(&c->d_normal, n0 * 4));
    VCT_CHECK(c, cudaMemsetAsync(c->d_color, 0, off * 4, c->stream)); VCT_CHECK(c, cudaMemsetAsync(c->d_radiance, 0, off * 4, c->stream));
    VCT_CHECK(c, cudaMemsetAsync(c->d_normal, 0, n0 * 4, c->stream));
    for (int i = 0; i < 2; ++i) { const size_t nb = std::max<size_t>(n0 / 8, 64); VCT_CHECK(c, cudaMalloc(&c->d_seg[i], nb)); VCT_CHECK(c, cudaMemsetAsync(c->d_seg[i], 0, nb, c->stream)); }
    c->seg_valid = false; c->seg_cur = 0;
    if (make_pyramid_texture(c, &c->radiance_arr, &c->radiance_tex, &c->radiance_tex_point, &c->radiance_tex_last, c->radiance_surf, &c->d_pub_mask_radiance)) return 1;
    const int ws = c->cfg.world_size > 1 ? c->cfg.world_size : 1, r = c->cfg.world_size > 1 ? c->cfg.rank : 0;
    int stripe = c->cfg.slab_stripe;
    if (stripe == 0 && ws > 1) { const char* e = getenv("VCT_SLAB_STRIPE"); if (e && *e) stripe = atoi(e); }        // tuning override for hosts that leave the field 0
    // default: stripes of 16 layers from 4 ranks on (measured on 8 B200s, Sponza 256^3: 0.336 ms against 0.361 ms with contiguous slabs — the scene
    // fills the middle half of z; at 2 ranks the contiguous halves are balanced already and the stripes' shared triangles cost 3 %)
    if (stripe == 0 && ws >= 4) stripe = 16;
    if (stripe < 0 || (stripe > 0 && ws > 1 && (c->D / ws) % stripe)) stripe = 0;                                    // -1, or does not fit this volume: contiguous slabs
    const int T = stripe > 0 && ws > 1 ? stripe : c->D / ws;
    if (T < 1 || c->D % ws) { c->error = "dim must be divisible by world_size"; return 1; }
    if (ws > 1 && stripe > 0 && (T % 16 || (T & (T - 1)))) { c->error = "slab_stripe must be a power of two >= 16"; return 1; }
    c->st.T = T; c->st.N = ws; c->st.rank = r; c->st.count = c->D / (T * ws);
    return 0;
}

int clamp_levels(int dim, int levels) {
    int lg = 0; while ((1 << (lg + 1)) <= dim) lg++;
    return std::min(std::max(levels, 1), std::min(lg + 1, VCT_MAX_LEVELS));
}

// Frame constants travel as a KERNEL PARAMETER when they fit (<= 8 actors: 3972 bytes): a copy-engine transfer in front
// of the first kernel of a frame costs a copy->compute dependency (~10 us on B200), a one-block kernel does not.
constexpr int kBlobActors = 8;
struct FrameBlob { FrameConst fc; Mat4 models[kBlobActors]; float nmats[9 * kBlobActors]; };
static_assert(sizeof(FrameBlob) <= 4096, "kernel parameter space");
static_assert(sizeof(FrameConst) % 4 == 0 && sizeof(FrameBlob) % 4 == 0, "blob layout");
__global__ void k_upload_frame(const __grid_constant__ FrameBlob blob, uint32_t* __restrict__ dst, int n_words) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&blob);
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
}

// device blob [FrameConst | models | nmats] + its pinned staging ring, sized for the current actor count
int make_frame_blob(vct_ctx* c) {
    const size_t bytes = sizeof(FrameConst) + (size_t)c->n_actors * (sizeof(Mat4) + 36);
    if (c->d_frame_blob && bytes <= c->frame_blob_bytes) {
        c->d_models = reinterpret_cast<Mat4*>((char*)c->d_frame_blob + sizeof(FrameConst)); c->d_nmats = reinterpret_cast<float*>(c->d_models + c->n_actors);
        return 0;
    }
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_frame_blob); c->d_frame_blob = nullptr;
    if (c->h_stage) { cudaFreeHost(c->h_stage); c->h_stage = nullptr; }
    VCT_CHECK(c, cudaMalloc(&c->d_frame_blob, bytes));
    VCT_CHECK(c, cudaMallocHost((void**)&c->h_stage, 4 * bytes));
    c->frame_blob_bytes = bytes;
    c->d_fc = reinterpret_cast<FrameConst*>(c->d_frame_blob);
    c->d_models = reinterpret_cast<Mat4*>((char*)c->d_frame_blob + sizeof(FrameConst)); c->d_nmats = reinterpret_cast<float*>(c->d_models + c->n_actors);
    return 0;
}

// concatenate the per-actor meshes into the device arrays the kernels index
int finalize_scene(vct_ctx* c) {
    if (!c->scene_dirty) return 0;
    std::sort(c->meshes.begin(), c->meshes.end(), [](const HostMesh& a, const HostMesh& b) { return a.actor < b.actor; });
    for (void** p : {(void**)&c->d_vertices, (void**)&c->d_vactor, (void**)&c->d_indices, (void**)&c->d_trimat, (void**)&c->d_wpos,
                     (void**)&c->d_wnrm, (void**)&c->d_wT, (void**)&c->d_wB, (void**)&c->d_setup, (void**)&c->d_tile_queue, (void**)&c->d_expand_queue, (void**)&c->d_pixel_queue}) { cudaFree(*p); *p = nullptr; }
    c->n_vertices = c->h_vertices.size() / 14; c->n_tris = c->h_trimat.size();
    c->n_actors = 0; for (auto& m : c->meshes) c->n_actors = std::max(c->n_actors, m.actor + 1);
    if (make_frame_blob(c)) return 1;
    if (!c->n_vertices || !c->n_tris) { c->scene_dirty = false; return 0; }
    const size_t nv = c->n_vertices, nt = c->n_tris;
    VCT_CHECK(c, cudaMalloc(&c->d_vertices, nv * 56)); VCT_CHECK(c, cudaMalloc(&c->d_vactor, nv * 4));
    VCT_CHECK(c, cudaMalloc(&c->d_indices, nt * 12)); VCT_CHECK(c, cudaMalloc(&c->d_trimat, nt * 4));
    VCT_CHECK(c, cudaMalloc(&c->d_wpos, nv * 16)); VCT_CHECK(c, cudaMalloc(&c->d_wnrm, nv * 16)); VCT_CHECK(c, cudaMalloc(&c->d_wT, nv * 16)); VCT_CHECK(c, cudaMalloc(&c->d_wB, nv * 16));
    // one setup per queued (sub-)triangle (camera pass: up to two after near clipping); tile queue sized for the
    // scene: every triangle may push one tile, plus headroom for the multi-tile ones
    c->setup_cap = 2 * nt + 64;
    c->tile_queue_cap = 6 * nt + ((size_t)8 << 20);      // (config 5, 64 Mi triangles: ~3 shadow-map tiles per triangle)
    VCT_CHECK(c, cudaMalloc(&c->d_setup, c->setup_cap * std::max(vctk_tile_setup_bytes(), vctk_vox_setup_bytes())));
    VCT_CHECK(c, cudaMalloc(&c->d_tile_queue, c->tile_queue_cap * 8));
    c->expand_cap = 2 * nt + ((size_t)1 << 20);
    VCT_CHECK(c, cudaMalloc(&c->d_expand_queue, c->expand_cap * 8));
    c->pixel_cap = 2 * nt + ((size_t)1 << 20);
    VCT_CHECK(c, cudaMalloc(&c->d_pixel_queue, c->pixel_cap * 8));
    VCT_CHECK(c, cudaMemcpy(c->d_vertices, c->h_vertices.data(), nv * 56, cudaMemcpyHostToDevice));
    VCT_CHECK(c, cudaMemcpy(c->d_vactor, c->h_vactor.data(), nv * 4, cudaMemcpyHostToDevice));
    VCT_CHECK(c, cudaMemcpy(c->d_indices, c->h_indices.data(), nt * 12, cudaMemcpyHostToDevice));
    VCT_CHECK(c, cudaMemcpy(c->d_trimat, c->h_trimat.data(), nt * 4, cudaMemcpyHostToDevice));
    c->scene_dirty = false;
    return 0;
}

// phong.frag:145-177 in fp32, one rounding per operation: r = h*tan(theta/2); lod = log2(max(1, 2r)); h += r
void build_schedule(ConeSchedule& t, const vct_cone_settings& cs, int levels) {
    const float tan_half = std::tan(cs.cone_angle / 2.0f), max_lod = (float)(levels - 1);
    const int steps = std::min(std::max(cs.steps, 0), VCT_MAX_CONE_STEPS);
    int np = 0, nl = steps;
    volatile float h = cs.cone_initial_height;
    for (int i = 0; i < steps; ++i) {
        volatile float r = h * tan_half;
        volatile float two_r = 2.0f * r;
        volatile float lod = std::log2(two_r > 1.0f ? two_r : 1.0f);
        volatile float lambda = lod + cs.lod_offset;
        t.h[i] = h; t.lambda[i] = lambda;
        if (!(lambda > 0.5f) && np == i) np = i + 1;
        if (lambda >= max_lod && nl == steps) nl = i;
        h = h + r;
    }
    if (nl < np) nl = np;
    t.n_point = np; t.n_last = nl; t.steps = steps; t.pad = 0;
}

int upload_frame(vct_ctx* c, const vct_frame_params* p) {
    if (!p) return fail(c, "null frame params");
    if (p->diffuse_cone.steps > 64 || p->specular_cone.steps > 64 || p->diffuse_cone.steps < 0 || p->specular_cone.steps < 0) return fail(c, "cone steps must be in [0,64]");
    if (p->voxel_fill_holes && c->cfg.world_size > 1)
        return fail(c, "voxelFillHoles reads the 3x3x3 neighbourhood across the borders of a rank's z layers (voxelFillHoles.comp:8-36): not supported with world_size > 1");
    if (!(p->voxelize_multiplier >= 0.0f && p->voxelize_multiplier <= 8.0f) || (p->voxelize_multiplier > 0.0f && (int)(p->voxelize_multiplier * (float)c->D) < 1))
        return fail(c, "voxelize_multiplier: 0 (= 1) or a factor up to 8 that leaves a viewport of at least one pixel");
    if (p->conservative_raster != VCT_RASTER_CENTER && p->conservative_raster != VCT_RASTER_MSAA)
        return fail(c, "conservative_raster: VCT_RASTER_CENTER or VCT_RASTER_MSAA (GL_CONSERVATIVE_RASTERIZATION_NV is not built)");
    for (int i = 0; i < 8; ++i) if (!(p->msaa_samples[i] >= 0.0f && p->msaa_samples[i] < 1.0f)) return fail(c, "msaa_samples: positions are pixel fractions in [0, 1)");
    if (p->radiance_dilate) return fail(c, "radianceDilate is malformed in the reference (injectRadiance.comp:59-64) and is not supported");
    if (finalize_scene(c)) return 1;
    FrameConst& f = c->h_fc;
    auto cp = [](Mat4& d, const float* s) { std::memcpy(d.m, s, 64); };
    cp(f.projection, p->projection); cp(f.view, p->view); cp(f.lp, p->lp); cp(f.lv, p->lv); cp(f.ls, p->ls); cp(f.ls_inverse, p->ls_inverse);
    cp(f.mvp_x, p->mvp_x); cp(f.mvp_y, p->mvp_y); cp(f.mvp_z, p->mvp_z);
    f.p = *p; f.D = c->D; f.L = c->L; f.S = c->S; f.W = c->W; f.H = c->H;
    f.n_lights = c->n_lights; std::memcpy(f.lights, c->h_lights, sizeof f.lights);
    f.st = c->st;
    build_schedule(f.sched_diffuse, p->diffuse_cone, c->L); build_schedule(f.sched_specular, p->specular_cone, c->L);
    if (c->n_actors <= kBlobActors) {                    // by value through the launch: no copy engine on the frame's critical path
        static_assert(offsetof(FrameBlob, models) == sizeof(FrameConst), "blob layout");
        FrameBlob b; b.fc = f;
        std::vector<Mat4> models(std::max(c->n_actors, 1)); std::vector<float> nm((size_t)std::max(c->n_actors, 1) * 9);
        vctk_fill_models(c, models.data(), nm.data());
        // device layout is [FrameConst | models[n_actors] | nmats[9 n_actors]] (make_frame_blob): pack the same way
        unsigned char* raw = reinterpret_cast<unsigned char*>(&b) + sizeof(FrameConst);
        std::memcpy(raw, models.data(), (size_t)c->n_actors * sizeof(Mat4));
        std::memcpy(raw + (size_t)c->n_actors * sizeof(Mat4), nm.data(), (size_t)c->n_actors * 36);
        const int n_words = (int)((sizeof(FrameConst) + (size_t)c->n_actors * (sizeof(Mat4) + 36)) / 4);
        k_upload_frame<<<1, 256, 0, c->stream>>>(b, reinterpret_cast<uint32_t*>(c->d_frame_blob), n_words);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) return fail(c, "k_upload_frame launch failed");
    } else {   // one pinned, asynchronous copy: frame constants + per-actor matrices
        const unsigned slot = c->stage_next++ & 3u;
        VCT_CHECK(c, cudaEventSynchronize(c->stage_ev[slot]));               // the copy that last read this slot (4 calls ago) is done
        unsigned char* h = c->h_stage + (size_t)slot * c->frame_blob_bytes;
        std::memcpy(h, &f, sizeof f);
        vctk_fill_models(c, reinterpret_cast<Mat4*>(h + sizeof f), reinterpret_cast<float*>(h + sizeof f + (size_t)c->n_actors * sizeof(Mat4)));
        VCT_CHECK(c, cudaMemcpyAsync(c->d_frame_blob, h, sizeof f + (size_t)c->n_actors * (sizeof(Mat4) + 36), cudaMemcpyHostToDevice, c->stream));
        VCT_CHECK(c, cudaEventRecord(c->stage_ev[slot], c->stream));
    }
    if (c->tables_dirty) {                      // texture / material tables change only on upload
        for (int i = 0; i < c->n_materials; ++i) {
            const DevMaterial& m = c->h_mat[i];
            for (int t : {m.diffuse_tex, m.specular_tex, m.normal_tex, m.roughness_tex, m.metallic_tex, m.alpha_tex})
                if (t >= 0 && !c->h_tex[t].level[0]) return fail(c, "a material names a texture that was never uploaded (vct_upload_texture before the first pass)");
        }
        VCT_CHECK(c, cudaMemcpyAsync(c->d_tex, c->h_tex, sizeof(DevTexture) * VCT_MAX_TEXTURES, cudaMemcpyHostToDevice, c->stream));
        VCT_CHECK(c, cudaMemcpyAsync(c->d_mat, c->h_mat, sizeof(DevMaterial) * VCT_MAX_MATERIALS, cudaMemcpyHostToDevice, c->stream));
        c->tables_dirty = false;
    }
    return 0;
}

enum { EV_START, EV_SHADOW, EV_WARP, EV_CLEAR, EV_VOXEL, EV_TRANSFER, EV_INJECT, EV_MIP, EV_XCHG, EV_GBUF, EV_TRACE, EV_COUNT };

struct Graph { vct_ctx* c; bool timed; int rec(int e) { if (timed && c->profiling >= 1) { cudaError_t r = cudaEventRecord(c->ev[e], c->stream); if (r != cudaSuccess) { c->error = cudaGetErrorString(r); return 1; } } return 0; } };

int ensure_color_texture(vct_ctx* c) {
    if (c->color_arr) return 0;
    return make_pyramid_texture(c, &c->color_arr, &c->color_tex, &c->color_tex_point, &c->color_tex_last, c->color_surf, &c->d_pub_mask_color);
}
int ensure_scratch(vct_ctx* c) {
    if (c->d_scratch) return 0;
    VCT_CHECK(c, cudaMalloc(&c->d_scratch, (size_t)c->D * c->D * c->D * 4));
    return 0;
}
int zero_info(vct_ctx* c) {
    VCT_CHECK(c, cudaMemsetAsync(c->d_counters, 0, 3 * sizeof(unsigned), c->stream));     // glClearNamedBufferData, Application.cpp:581
    vct_prof_mark(c, "memset");
    return 0;
}
// The GI passes of one frame.  Sparse frame (DESIGN.md "segment masks"): clear, transfer and the mip chain visit only the
// x-row segments that hold (or held last frame) a fragment.  The first frame, and any frame after something wrote a
// volume behind the library's back, runs the dense kernels and (re)builds the masks.
struct FramePlan { int n_chains, key; bool maskable, sparse; };
FramePlan plan_frame(const vct_ctx* c) {
    const vct_frame_params& p = c->h_fc.p;
    FramePlan f;
    f.n_chains = (p.mip_color_chain || !p.draw_radiance) ? 2 : 1;
    f.key = f.n_chains * 2 + (p.draw_radiance ? 1 : 0);                   // which pyramids are filtered / published
    f.maskable = vctk_sparse_supported(c) && !c->sparse_off && !p.voxel_fill_holes;
    f.sparse = f.maskable && c->seg_valid && c->seg_key == f.key;
    return f;
}
int gi_body(vct_ctx* c, Graph& g, bool cleared_by_frame_begin = false) {
    const vct_frame_params& p = c->h_fc.p;
    const bool single = c->cfg.world_size <= 1;
    const FramePlan plan = plan_frame(c);
    const int n_chains = plan.n_chains, key = plan.key;
    const bool maskable = plan.maskable, sparse = plan.sparse;
    NvtxPhases nv; nv.next("voxelize");                      // clear + voxelise, like the reference's voxelizeTimer (Application.cpp:583-755)
    auto next_range = [&](const char* name) { nv.next(name); };
    if (cleared_by_frame_begin) {
        if (g.rec(EV_CLEAR)) return 1;
    } else if (sparse) {
        if (vctk_clear_masked(c) || g.rec(EV_CLEAR)) return 1;              // also zeroes VoxelizeInfo, the raster queues, cone_steps
    } else {
        const size_t nb = std::max<size_t>((size_t)c->D * c->D * c->D / 8, 64);
        VCT_CHECK(c, cudaMemsetAsync(c->d_seg[c->seg_cur], 0, nb, c->stream));
        vct_prof_mark(c, "memset");
        if (vctk_clear_voxels(c, true) || g.rec(EV_CLEAR)) return 1;
    }
    // sparse frame, temporal filter off, deterministic raster voxeliser: the resolve kernel does transferVoxels for its voxels
    bool transfer_done = false;
    if (vctk_voxelize(c, false, true, sparse && !p.temporal_filter_radiance, &transfer_done) || g.rec(EV_VOXEL)) return 1;
    next_range("transfer");
    if (!transfer_done && (sparse ? vctk_transfer_masked(c) : vctk_transfer(c))) return 1;
    if (g.rec(EV_TRANSFER)) return 1;
    next_range("radiance");
    if (vctk_inject(c)) return 1;
    if (p.voxel_fill_holes) { if (ensure_scratch(c) || vctk_fill_holes(c)) return 1; }
    if (g.rec(EV_INJECT)) return 1;
    next_range("mipmap");
    // single GPU: the chain that the cone tracer samples is written straight into its texture array
    // (sharded with attached peers: likewise for the own slab; the exchange publishes the remote slabs)
    const bool direct = single || vctk_xchg_ready(c);
    if (direct && !p.draw_radiance && ensure_color_texture(c)) return 1;
    const int which[2] = {VCT_VOL_RADIANCE, VCT_VOL_COLOR};
    const int publish[2] = {direct && p.draw_radiance, direct && !p.draw_radiance};
    // (sharded with attached peers: the chain also stores its levels >= 1 into the peers' pyramids — compute and exchange in one kernel)
    if (vctk_mip_chains(c, n_chains, which, publish, 0, sparse, !single && vctk_xchg_ready(c))) return 1;   // both pyramids, one launch
    // this frame's mask bounds the support of all three level-0 volumes (dense temporal frames: k_transfer flagged the history)
    c->seg_valid = maskable;
    c->last_frame_sparse = sparse;
    c->seg_key = key;
    c->seg_cur ^= 1;
    return g.rec(EV_MIP);
}


// ---- single-process multi-GPU (vct_config.n_devices > 1): the handle the host holds is rank 0's context; `group` lists the other
// ranks' contexts.  Every call that changes state or enqueues work runs on all of them, each on its own worker thread (a frame is
// ~20 launches per device: issued one device after the other they would cost more host time than the sharded frame takes on the
// devices).  Nothing in a frame call blocks the host on a device, which matters: a rank's unpack kernel spins until its peers'
// push kernels have run, and those are launched by the other workers.
struct GroupWorker {
    std::thread th; std::mutex m; std::condition_variable cv;
    std::function<int()> job; bool has_job = false, quit = false, done = true; int rc = 0;
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job); has_job = false;
            lk.unlock(); const int r = j(); lk.lock();
            rc = r; done = true; cv.notify_all();
        }
    }
};
struct Group { std::vector<GroupWorker*> workers; };
template <class F> int group_run(vct_ctx* leader, F f) {
    Group* g = reinterpret_cast<Group*>(leader->group_state);
    leader->in_fan = true;
    for (size_t i = 0; i < leader->group.size(); ++i) {
        GroupWorker* w = g->workers[i]; vct_ctx* m = leader->group[i];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [f, m]() { return f(m); }; w->has_job = true; w->done = false; w->cv.notify_all();
    }
    int rc = f(leader);
    for (size_t i = 0; i < leader->group.size(); ++i) {
        GroupWorker* w = g->workers[i];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc && !rc) { rc = w->rc; leader->error = "rank " + std::to_string(i + 1) + ": " + leader->group[i]->error; }
    }
    leader->in_fan = false;
    return rc;
}
#define VCT_FAN(c, expr) do { if ((c) && !(c)->group.empty() && !(c)->in_fan) return group_run((c), [=](vct_ctx* c) { return (expr); }); } while (0)
#define VCT_NO_GROUP(c, what) do { if ((c) && !(c)->group.empty()) return fail((c), what ": not available on a multi-device handle (vct_config.n_devices > 1)"); } while (0)

}  // namespace

extern "C" {

const char* vct_last_error(const vct_ctx* c) { return c ? c->error.c_str() : g_create_error.c_str(); }

static int create_group(const vct_config* cfg, vct_ctx** out);
int vct_create(const vct_config* cfg, vct_ctx** out) {
    if (!cfg || !out) { g_create_error = "vct_create: null argument"; return 1; }
    *out = nullptr;
    if (cfg->n_devices > 1) return create_group(cfg, out);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_create_error = std::string("vct_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback"; return 1; }
    if (cfg->dim < 4 || (cfg->dim & (cfg->dim - 1)) || cfg->dim > 1024) { g_create_error = "vct_create: dim must be a power of two in [4,1024]"; return 1; }
    if (cfg->shadow_size < 1 || cfg->width < 1 || cfg->height < 1) { g_create_error = "vct_create: bad sizes"; return 1; }
    if (cfg->world_size > 1 && (cfg->rank < 0 || cfg->rank >= cfg->world_size || cfg->dim % cfg->world_size)) { g_create_error = "vct_create: bad rank/world_size"; return 1; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = std::string("vct_create: cudaSetDevice: ") + cudaGetErrorString(e); return 1; }
    vct_ctx* c = new vct_ctx();
    c->cfg = *cfg; c->D = cfg->dim; c->L = clamp_levels(cfg->dim, cfg->levels); c->S = cfg->shadow_size; c->W = cfg->width; c->H = cfg->height;
    auto bail = [&](const char* what) { g_create_error = std::string("vct_create: ") + what + ": " + c->error; vct_destroy(c); return 1; };
    if (const char* v = getenv("VCT_SPARSE")) c->sparse_off = atoi(v) == 0;   // VCT_SPARSE=0: dense kernels every frame (A/B and tests)
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
    for (auto& ev : c->ev) if (cudaEventCreate(&ev) != cudaSuccess) return bail("event");
    for (auto& ev : c->stage_ev) if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail("copy stream");
    if (cudaEventCreateWithFlags(&c->ev_image_ready, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_copy_done[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_copy_done[1], cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (make_frame_blob(c)) return bail("frame constants");
    if (make_volumes(c)) return bail("volumes");
    const int N = VCT_WARP_DIM;
    auto alloc = [&](void** p, size_t bytes) { cudaError_t r = cudaMalloc(p, bytes); if (r != cudaSuccess) { c->error = cudaGetErrorString(r); return 1; } cudaMemsetAsync(*p, 0, bytes, c->stream); return 0; };
    c->frag_cap = cfg->max_fragments > 0 ? (size_t)cfg->max_fragments : ((size_t)8 << 20);
    if (alloc((void**)&c->d_occ, N * N * N * 4) || alloc((void**)&c->d_warpmap, N * N * N * (8 + 32)) || alloc((void**)&c->d_wlo, N * N * N * 8) || alloc((void**)&c->d_whi, N * N * N * 8) ||
        alloc((void**)&c->d_shadow_base, 2 * (size_t)c->S * c->S * 4) || alloc(&c->d_shadow_mm, ((size_t)(c->S / 4 + 1) * (c->S / 4 + 1) + (size_t)(c->S / 64 + 1) * (c->S / 16 + 1)) * 8) || alloc((void**)&c->d_inject_list, ((size_t)(c->S / 64 + 1) * (c->S / 16 + 1) + 1) * 4) || alloc((void**)&c->d_vis, (size_t)c->W * c->H * 8) || alloc((void**)&c->d_image, 2 * vctk_image_rows(c) * (size_t)c->W * 4) ||
        alloc((void**)&c->d_frags, c->frag_cap * vctk_frag_bytes()) || alloc((void**)&c->d_displaced, c->frag_cap) || alloc(&c->d_long_queue, (c->long_cap + 16) * 16) || alloc(&c->d_huge_items, c->frag_cap * 16) || alloc(&c->d_huge_aux, vctk_huge_aux_bytes()) || alloc((void**)&c->d_warp_scratch, N * N * N * 4) ||
        alloc((void**)&c->d_counters, sizeof(Counters)) ||
        alloc((void**)&c->d_tex, sizeof(DevTexture) * VCT_MAX_TEXTURES) || alloc((void**)&c->d_mat, sizeof(DevMaterial) * VCT_MAX_MATERIALS))
        return bail("cudaMalloc");
    // overflow of a fixed-capacity buffer is also flagged in mapped host memory, so that the next entry point sees it without a sync
    if (cudaHostAlloc((void**)&c->h_overflow, 4 * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) { c->error = "cudaHostAlloc"; return bail("overflow flag"); }
    cudaMemsetAsync(c->d_vis, 0xFF, (size_t)c->W * c->H * 8, c->stream);   // "no geometry" everywhere: a cone trace before the first visibility pass shades background
    for (int i = 0; i < 4; ++i) c->h_overflow[i] = 0u;       // [0] overflow, [1] a long per-voxel list was met (voxelize.cu), [2] [3] inject block count + generation
    { unsigned* dp = nullptr; if (cudaHostGetDevicePointer((void**)&dp, c->h_overflow, 0) != cudaSuccess || cudaMemcpyAsync(&c->d_counters->overflow_host, &dp, sizeof dp, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { c->error = "mapped overflow flag"; return bail("overflow flag"); } }
    for (int i = 0; i < VCT_MAX_MATERIALS; ++i) { DevMaterial& m = c->h_mat[i]; m.diffuse_tex = m.specular_tex = m.normal_tex = m.roughness_tex = m.metallic_tex = m.alpha_tex = -1; m.shininess = 32.0f; }
    c->d_shadow = c->d_shadow_base;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return bail("sync");
    *out = c;
    return 0;
}

// vct_config.n_devices > 1: one context per device (rank i on devices[i]), peer access both ways, slab exchange attached
static int create_group(const vct_config* cfg, vct_ctx** out) {
    const int n = cfg->n_devices;
    if (n > VCT_MAX_PEERS) { g_create_error = "vct_create: n_devices must be <= 8"; return 1; }
    std::vector<vct_ctx*> ctx((size_t)n, nullptr);
    auto undo = [&](const std::string& why) { for (vct_ctx* m : ctx) if (m) vct_destroy(m); g_create_error = why; return 1; };
    for (int r = 0; r < n; ++r) {
        vct_config one = *cfg;
        one.n_devices = 0; one.devices = nullptr; one.device = cfg->devices ? cfg->devices[r] : r; one.rank = r; one.world_size = n;
        if (vct_create(&one, &ctx[r])) return undo("vct_create: device " + std::to_string(one.device) + ": " + g_create_error);
    }
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            cudaSetDevice(ctx[a]->cfg.device);
            int can = 0; cudaDeviceCanAccessPeer(&can, ctx[a]->cfg.device, ctx[b]->cfg.device);
            if (!can) return undo("vct_create: devices " + std::to_string(ctx[a]->cfg.device) + " and " + std::to_string(ctx[b]->cfg.device) + " have no peer access (NVLink / PCIe P2P)");
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctx[b]->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return undo(std::string("vct_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    for (int r = 0; r < n; ++r) if (vct_exchange_setup(ctx[r])) return undo("vct_create: slab exchange: " + ctx[r]->error);
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            vct_peer pb; vct_exchange_local(ctx[b], &pb);
            if (vct_exchange_attach(ctx[a], b, &pb)) return undo("vct_create: slab exchange: " + ctx[a]->error);
        }
    vct_ctx* leader = ctx[0];
    Group* g = new Group();
    for (int r = 1; r < n; ++r) {
        leader->group.push_back(ctx[r]);
        GroupWorker* w = new GroupWorker();
        w->th = std::thread([w] { w->loop(); });
        g->workers.push_back(w);
    }
    leader->group_state = g;
    *out = leader;
    return 0;
}

int vct_destroy(vct_ctx* c) {
    if (!c) return 0;
    if (c->group_state) {
        Group* g = reinterpret_cast<Group*>(c->group_state);
        for (GroupWorker* w : g->workers) { { std::lock_guard<std::mutex> lk(w->m); w->quit = true; w->cv.notify_all(); } w->th.join(); delete w; }
        delete g; c->group_state = nullptr;
        for (vct_ctx* m : c->group) { cudaSetDevice(m->cfg.device); if (m->stream) cudaStreamSynchronize(m->stream); }
        cudaSetDevice(c->cfg.device); if (c->stream) cudaStreamSynchronize(c->stream);
        for (vct_ctx* m : c->group) vct_destroy(m);
        c->group.clear();
    }
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_volumes(c);
    vctk_xchg_free(c);
    for (void* p : {(void*)c->d_occ, (void*)c->d_warpmap, (void*)c->d_wlo, (void*)c->d_whi, (void*)c->d_shadow_base, c->d_shadow_mm, (void*)c->d_inject_list, (void*)c->d_vis, (void*)c->d_image, c->d_frags, (void*)c->d_displaced, c->d_long_queue, c->d_huge_items, c->d_huge_aux, (void*)c->d_warp_scratch,
                    c->d_tile_queue, c->d_expand_queue, c->d_pixel_queue, c->d_frame_blob, (void*)c->d_counters, (void*)c->d_trace_tiles, c->d_dbg,
                    (void*)c->d_tex, (void*)c->d_mat, (void*)c->d_vertices, (void*)c->d_vactor, (void*)c->d_indices, (void*)c->d_trimat, (void*)c->d_wpos,
                    (void*)c->d_wnrm, (void*)c->d_wT, (void*)c->d_wB, c->d_setup})
        cudaFree(p);
    for (void* p : c->tex_alloc) cudaFree(p);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->stage_ev) if (ev) cudaEventDestroy(ev);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_image_ready) cudaEventDestroy(c->ev_image_ready);
    for (auto& ev : c->ev_copy_done) if (ev) cudaEventDestroy(ev);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_overflow) cudaFreeHost(c->h_overflow);
    for (auto& ev : c->prof_pool) cudaEventDestroy(ev);
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

// VCT::remake (reference src/Application.h:118-129): destroy + create with clamped level count
int vct_remake(vct_ctx* c, int dim, int levels) { VCT_NO_GROUP(c, "vct_remake");
    if (!c) return 1;
    if (dim < 4 || (dim & (dim - 1)) || dim > 1024) return fail(c, "vct_remake: dim must be a power of two in [4,1024]");
    if (c->cfg.world_size > 1) {
        if (dim % c->cfg.world_size) return fail(c, "vct_remake: dim must be divisible by world_size");
        // the staging buffer of the sparse exchange is mapped by the peers (cudaIpc): freeing it here would leave them storing into freed memory
        if (c->d_xchg) return fail(c, "vct_remake: not supported once the sparse slab exchange is set up (vct_exchange_setup); create new contexts on every rank");
    }
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    free_volumes(c);
    vctk_xchg_free(c);
    cudaFree(c->d_dbg); c->d_dbg = nullptr;                      // sized by dim
    c->D = dim; c->L = clamp_levels(dim, levels); c->cfg.dim = dim; c->cfg.levels = c->L;
    return make_volumes(c);
}

int vct_upload_mesh(vct_ctx* c, int actor, const void* vertices, size_t n_vertices, size_t stride, const uint32_t* indices, size_t n_indices, const int32_t* material_of_triangle) { VCT_FAN(c, vct_upload_mesh(c, actor, vertices, n_vertices, stride, indices, n_indices, material_of_triangle));
    if (!c) return 1;
    if (stride < 56 || !vertices || !indices || n_indices % 3 || actor < 0) return fail(c, "vct_upload_mesh: bad arguments (Vertex stride is 56 bytes, src/Graphics/Mesh.h:72-76)");
    for (auto& m : c->meshes) if (m.actor == actor) return fail(c, "vct_upload_mesh: actor already has a mesh");
    if (actor >= 65536) return fail(c, "vct_upload_mesh: actor ids are 0..65535");
    for (size_t i = 0; i < n_indices; ++i) if (indices[i] >= n_vertices) return fail(c, "vct_upload_mesh: index out of range");
    if (material_of_triangle)
        for (size_t t = 0; t < n_indices / 3; ++t)
            if (material_of_triangle[t] < 0 || material_of_triangle[t] >= VCT_MAX_MATERIALS) return fail(c, "vct_upload_mesh: material id out of range");
    const size_t nv0 = c->h_vertices.size(), na0 = c->h_vactor.size(), ni0 = c->h_indices.size(), nt0 = c->h_trimat.size();
    try {
    HostMesh m{}; m.actor = actor; m.n_vertices = n_vertices; m.n_tris = n_indices / 3; m.vbase = c->h_vertices.size() / 14; m.tbase = c->h_trimat.size();
    for (int i = 0; i < 16; ++i) m.model.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    const size_t v0 = c->h_vertices.size();
    c->h_vertices.resize(v0 + n_vertices * 14);
    for (size_t i = 0; i < n_vertices; ++i) std::memcpy(&c->h_vertices[v0 + 14 * i], (const char*)vertices + i * stride, 56);
    c->h_vactor.insert(c->h_vactor.end(), n_vertices, actor);
    for (size_t i = 0; i < n_indices; ++i) c->h_indices.push_back(indices[i] + (uint32_t)m.vbase);
    for (size_t t = 0; t < n_indices / 3; ++t) c->h_trimat.push_back(material_of_triangle ? material_of_triangle[t] : 0);
    c->meshes.push_back(m);
    c->scene_dirty = true;
    return 0;
    } catch (const std::exception&) {                           // nothing may cross the C ABI; leave the scene as it was
        c->h_vertices.resize(nv0); c->h_vactor.resize(na0); c->h_indices.resize(ni0); c->h_trimat.resize(nt0);
        return fail(c, "vct_upload_mesh: out of host memory");
    }
}

int vct_upload_texture(vct_ctx* c, int tex, int width, int height, int channels, int levels, const void* pixels) { VCT_FAN(c, vct_upload_texture(c, tex, width, height, channels, levels, pixels));
    if (!c) return 1;
    if (tex < 0 || tex >= VCT_MAX_TEXTURES || width < 1 || height < 1 || !(channels == 1 || channels == 3 || channels == 4) || levels < 1 || levels > 16 || !pixels)
        return fail(c, "vct_upload_texture: bad arguments");
    size_t total = 0;
    for (int l = 0; l < levels; ++l) total += (size_t)std::max(1, width >> l) * std::max(1, height >> l) * channels;
    uint8_t* d = nullptr;
    VCT_CHECK(c, cudaMalloc(&d, total + 16));
    if (cudaError_t r = cudaMemcpy(d, pixels, total, cudaMemcpyHostToDevice); r != cudaSuccess) { cudaFree(d); return fail(c, cudaGetErrorString(r)); }
    if (c->tex_alloc[tex]) {                                    // re-upload (streaming / reload): the old block may still be read by queued work
        cudaStreamSynchronize(c->stream);
        cudaFree(c->tex_alloc[tex]);
    }
    c->tex_alloc[tex] = d;
    DevTexture& t = c->h_tex[tex]; t.w = width; t.h = height; t.ch = channels; t.levels = levels;
    size_t off = 0;
    for (int l = 0; l < levels; ++l) { t.level[l] = d + off; off += (size_t)std::max(1, width >> l) * std::max(1, height >> l) * channels; }
    c->n_textures = std::max(c->n_textures, tex + 1);
    c->tables_dirty = true;
    return 0;
}

int vct_set_material(vct_ctx* c, int material, const vct_material* m) { VCT_FAN(c, vct_set_material(c, material, m));
    if (!c) return 1;
    if (material < 0 || material >= VCT_MAX_MATERIALS || !m) return fail(c, "vct_set_material: bad arguments");
    for (int t : {m->diffuse_tex, m->specular_tex, m->normal_tex, m->roughness_tex, m->metallic_tex, m->alpha_tex})
        if (t < -1 || t >= VCT_MAX_TEXTURES) return fail(c, "vct_set_material: texture id out of range (-1 = none)");
    DevMaterial& d = c->h_mat[material];
    d.diffuse_tex = m->diffuse_tex; d.specular_tex = m->specular_tex; d.normal_tex = m->normal_tex; d.roughness_tex = m->roughness_tex; d.metallic_tex = m->metallic_tex; d.alpha_tex = m->alpha_tex;
    d.shininess = m->shininess; std::memcpy(d.diffuse, m->diffuse, 12);
    c->n_materials = std::max(c->n_materials, material + 1);
    c->tables_dirty = true;
    return 0;
}

int vct_set_actor_transform(vct_ctx* c, int actor, const float model[16]) { VCT_FAN(c, vct_set_actor_transform(c, actor, model));
    if (!c) return 1;
    if (!model) return fail(c, "vct_set_actor_transform: null matrix");
    for (auto& m : c->meshes) if (m.actor == actor) { std::memcpy(m.model.m, model, 64); return 0; }
    return fail(c, "vct_set_actor_transform: unknown actor");
}

int vct_set_lights(vct_ctx* c, const vct_light* lights, int n) { VCT_FAN(c, vct_set_lights(c, lights, n));
    if (!c) return 1;
    if (n < 0 || n > 8 || (n && !lights)) return fail(c, "vct_set_lights: 0..8 lights");
    std::memset(c->h_lights, 0, sizeof c->h_lights);
    if (n) std::memcpy(c->h_lights, lights, n * sizeof(vct_light));
    c->n_lights = n;
    return 0;
}

// ------------------------------------------------------------------------------------------- passes
#define PASS_PROLOGUE                                   \
    if (!c) return 1;                                   \
    cudaSetDevice(c->cfg.device);                       \
    if (overflow_seen(c)) return 1;                     \
    vct_prof_begin(c);                                  \
    if (upload_frame(c, p)) return 1;                   \
    vct_prof_mark(c, "h2d_params");

// A fixed-capacity buffer (fragment records, raster queues, triangle setups) overflowed in an EARLIER call: kernels flag that in
// mapped host memory, the next pass entry point reports it once (and does nothing else), then the context carries on.
static int overflow_seen(vct_ctx* c) {
    if (!c->h_overflow || !*(volatile unsigned*)c->h_overflow) return 0;
    const unsigned what = *(volatile unsigned*)c->h_overflow;
    *(volatile unsigned*)c->h_overflow = 0u;
    if (what == 2u) {
        c->error = "a voxel received more than 1024 fragments in a frame that ran without the long-list kernels (they start with the first frame that meets a long "
                   "per-voxel list, or at once in the warp modes): that voxel was left unresolved in the previous frame; this call was not executed, the next one "
                   "runs them and is exact";
        return 1;
    }
    c->error = "a fixed-capacity device buffer overflowed during an earlier call and fragments or tiles were dropped (fragment buffer, raster queues or triangle setups): "
               "raise vct_config.max_fragments; this call was not executed, the next one will be";
    return 1;
}
int vct_shadowmap(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_shadowmap(c, p)); PASS_PROLOGUE; return vctk_transform_vertices(c) || vctk_shadowmap(c) || vctk_shadow_minmax(c); }
int vct_occupancy(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_occupancy(c, p)); PASS_PROLOGUE; return vctk_transform_vertices(c) || vctk_voxelize(c, true); }
int vct_warpmap(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_warpmap(c, p)); PASS_PROLOGUE; return vctk_warpmap(c); }
int vct_voxelize(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_voxelize(c, p)); PASS_PROLOGUE; c->seg_valid = false; return zero_info(c) || vctk_transform_vertices(c) || vctk_clear_voxels(c) || vctk_voxelize(c, false); }
int vct_transfer(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_transfer(c, p)); PASS_PROLOGUE; c->seg_valid = false; return vctk_transfer(c); }
int vct_inject(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_inject(c, p)); PASS_PROLOGUE; c->seg_valid = false; return vctk_inject(c); }
int vct_fill_holes(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_fill_holes(c, p)); PASS_PROLOGUE; c->seg_valid = false; return ensure_scratch(c) || vctk_fill_holes(c); }
int vct_gbuffer(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_gbuffer(c, p)); PASS_PROLOGUE; return vctk_transform_vertices(c) || vctk_visibility(c); }
int vct_cone_trace(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_cone_trace(c, p));
    PASS_PROLOGUE;
    if (!p->draw_radiance && !c->color_arr) { if (ensure_color_texture(c) || vctk_publish(c, VCT_VOL_COLOR)) return 1; }
    VCT_CHECK(c, cudaMemsetAsync(&c->d_counters->cone_steps, 0, sizeof(unsigned long long), c->stream));
    vct_prof_mark(c, "memset");
    return vctk_cone_trace(c);
}
// Application::debugVoxels (Application.cpp:1222-1275): the non-empty voxels of the pyramid the frame traces, as cubes, into the image buffer
int vct_debug_voxels(vct_ctx* c, const vct_frame_params* p) { VCT_NO_GROUP(c, "vct_debug_voxels");
    PASS_PROLOGUE;
    if (!p->draw_radiance && !c->color_arr) { if (ensure_color_texture(c) || vctk_publish(c, VCT_VOL_COLOR)) return 1; }
    return vctk_debug_voxels(c);
}
int vct_mip(vct_ctx* c, int which) { VCT_FAN(c, vct_mip(c, which));
    if (!c) return 1;
    cudaSetDevice(c->cfg.device);
    if (which != VCT_VOL_RADIANCE && which != VCT_VOL_COLOR) return fail(c, "vct_mip: radiance or colour volume only");
    c->seg_valid = false;
    const bool publish = c->cfg.world_size <= 1 && !(which == VCT_VOL_COLOR && !c->color_arr);
    return vctk_mip(c, which, 0, publish);
}
// filterRadiance.comp's `kernelMode` uniform (:9-13): 0 = BOX2 (what the host dispatches, = vct_mip), 1 = BOX3 (27 taps x 0.037),
// 2 = CUBE (7 axial taps x 0.143); the reference declares the uniform and never sets it.  BOX3 / CUBE read across 2x2x2 cell borders,
// so they need the whole source level: single GPU only.
int vct_mip_kernel(vct_ctx* c, int which, int kernel_mode) { VCT_FAN(c, vct_mip_kernel(c, which, kernel_mode));
    if (!c) return 1;
    cudaSetDevice(c->cfg.device);
    if (which != VCT_VOL_RADIANCE && which != VCT_VOL_COLOR) return fail(c, "vct_mip_kernel: radiance or colour volume only");
    if (kernel_mode < 0 || kernel_mode > 2) return fail(c, "vct_mip_kernel: kernel_mode is 0 (BOX2), 1 (BOX3) or 2 (CUBE)");
    if (kernel_mode != 0 && c->cfg.world_size > 1) return fail(c, "vct_mip_kernel: BOX3 / CUBE read across the borders of a rank's z layers; world_size must be 1");
    c->seg_valid = false;
    const bool publish = c->cfg.world_size <= 1 && !(which == VCT_VOL_COLOR && !c->color_arr);
    return vctk_mip(c, which, kernel_mode, publish);
}
int vct_exchange(vct_ctx* c) { VCT_NO_GROUP(c, "vct_exchange");
    if (!c) return 1;
    cudaSetDevice(c->cfg.device);
    vct_prof_begin(c);
    const bool rad = c->h_fc.p.draw_radiance != 0;
    if (!rad && ensure_color_texture(c)) return 1;
    // the levels whose texel layers straddle the ranks' stripes are filtered here, on every rank, from the gathered level below them
    if (c->cfg.world_size > 1 && vctk_mip_top_sharded_level(c) + 1 < c->L && vctk_mip_tail(c, rad ? VCT_VOL_RADIANCE : VCT_VOL_COLOR)) return 1;
    return vctk_publish(c, rad ? VCT_VOL_RADIANCE : VCT_VOL_COLOR);
}

// ---- slab exchange over peer memory (exchange.cu); the caller only moves the cudaIpc handle blobs between the ranks once
int vct_exchange_setup(vct_ctx* c) { VCT_NO_GROUP(c, "vct_exchange_setup"); if (!c) return 1; cudaSetDevice(c->cfg.device); return vctk_xchg_setup(c); }
int vct_exchange_export(vct_ctx* c, void* handle) {
    if (!c || !handle) return 1;
    cudaSetDevice(c->cfg.device);
    if (!c->d_xchg) return fail(c, "vct_exchange_export: call vct_exchange_setup first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && VCT_EXCHANGE_HANDLE_BYTES >= 5 * 64, "ipc handle size");
    void* ptrs[5] = {c->d_xchg, c->d_radiance, c->d_color, c->d_image, c->d_shadow_base};
    for (int i = 0; i < 5; ++i) {
        cudaIpcMemHandle_t h;
        VCT_CHECK(c, cudaIpcGetMemHandle(&h, ptrs[i]));
        std::memcpy((char*)handle + 64 * i, &h, 64);
    }
    return 0;
}
int vct_exchange_import(vct_ctx* c, int rank, const void* handle) {
    if (!c || !handle) return 1;
    cudaSetDevice(c->cfg.device);
    if (!c->d_xchg) return fail(c, "vct_exchange_import: call vct_exchange_setup first");
    if (rank < 0 || rank >= c->cfg.world_size) return fail(c, "vct_exchange_import: bad rank");
    if (rank == c->cfg.rank) return 0;
    if (c->peer[rank].staging) return fail(c, "vct_exchange_import: this rank is already attached");
    void* ptrs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 5; ++i) {
        cudaIpcMemHandle_t h; std::memcpy(&h, (const char*)handle + 64 * i, 64);
        VCT_CHECK(c, cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->peer[rank].staging = ptrs[0]; c->peer[rank].radiance = ptrs[1]; c->peer[rank].color = ptrs[2]; c->peer[rank].image = ptrs[3]; c->peer[rank].shadow = ptrs[4];
    c->peer_ipc[rank] = true; c->peers_attached++;
    return 0;
}
int vct_exchange_local(vct_ctx* c, vct_peer* out) {
    if (!c || !out) return 1;
    if (!c->d_xchg) return fail(c, "vct_exchange_local: call vct_exchange_setup first");
    *out = c->peer[c->cfg.rank];
    return 0;
}
int vct_exchange_attach(vct_ctx* c, int rank, const vct_peer* peer) {
    if (!c || !peer) return 1;
    if (!c->d_xchg) return fail(c, "vct_exchange_attach: call vct_exchange_setup first");
    if (rank < 0 || rank >= c->cfg.world_size || !peer->staging || !peer->radiance || !peer->color || !peer->image || !peer->shadow) return fail(c, "vct_exchange_attach: bad arguments");
    if (rank == c->cfg.rank) return 0;
    if (c->peer[rank].staging) return fail(c, "vct_exchange_attach: this rank is already attached");
    c->peer[rank] = *peer; c->peer_ipc[rank] = false; c->peers_attached++;
    return 0;
}
int vct_frame_was_sparse(vct_ctx* c) { return c && c->last_frame_sparse ? 1 : 0; }
int vct_mask_parity(vct_ctx* c) { return c ? c->seg_cur : 0; }
int vct_slab_stripe(vct_ctx* c) { return c ? c->st.T : 0; }

// tail of a sharded frame with attached peers: exchange, (visibility), cone trace of the own tiles, image hand-over to rank 0
static int sharded_tail(vct_ctx* c, Graph& g, bool with_visibility) {
    { NvtxRange r("exchange"); if (vctk_xchg_frame(c, !c->last_frame_sparse) || g.rec(EV_XCHG)) return 1; }
    if (with_visibility) { NvtxRange r("gbuffer"); if (vctk_visibility(c)) return 1; }
    if (g.rec(EV_GBUF)) return 1;
    NvtxRange r("render");
    if (vctk_cone_trace(c) || vctk_xchg_image_sync(c)) return 1;
    return g.rec(EV_TRACE);
}

int vct_gi_passes(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_gi_passes(c, p));
    PASS_PROLOGUE;
    Graph g{c, true};
    if (g.rec(EV_START) || g.rec(EV_SHADOW) || g.rec(EV_WARP)) return 1;
    // sparse frame: vertex transform and masked clear share one launch (kept apart under per-kernel profiling)
    const bool fused_begin = plan_frame(c).sparse && c->profiling < 2 && c->n_vertices > 0;
    if (fused_begin ? vctk_frame_begin_masked(c) : vctk_transform_vertices(c)) return 1;
    if (gi_body(c, g, fused_begin)) return 1;
    if (c->cfg.world_size > 1) return vctk_xchg_ready(c) ? sharded_tail(c, g, false) : 0;   // no peers: caller all-gathers, then vct_exchange + vct_cone_trace
    if (g.rec(EV_XCHG) || g.rec(EV_GBUF)) return 1;
    NvtxRange r("render");
    if (vctk_cone_trace(c)) return 1;                         // cone_steps was zeroed by gi_body's clear
    return g.rec(EV_TRACE);
}

// Application::render in the reference's pass order (src/Application.cpp:196-1085)
int vct_frame(vct_ctx* c, const vct_frame_params* p) { VCT_FAN(c, vct_frame(c, p));
    NvtxRange total("total");
    PASS_PROLOGUE;
    Graph g{c, true};
    if (g.rec(EV_START)) return 1;
    { NvtxRange r("shadowmap"); if (vctk_transform_vertices(c) || vctk_shadowmap(c) || vctk_shadow_minmax(c) || g.rec(EV_SHADOW)) return 1; }
    if (p->warp_texture) { NvtxRange r("warpmap"); if (vctk_voxelize(c, true) || vctk_warpmap(c)) return 1; }
    if (g.rec(EV_WARP)) return 1;
    if (gi_body(c, g)) return 1;
    if (c->cfg.world_size > 1) return vctk_xchg_ready(c) ? sharded_tail(c, g, true) : 0;
    { NvtxRange r("gbuffer"); if (g.rec(EV_XCHG) || vctk_visibility(c) || g.rec(EV_GBUF)) return 1; }
    NvtxRange r("render");
    if (vctk_cone_trace(c)) return 1;                         // cone_steps was zeroed by gi_body's clear
    return g.rec(EV_TRACE);
}

// dead-shader equivalents
int vct_set_voxel_opacity(vct_ctx* c, float o) { VCT_FAN(c, vct_set_voxel_opacity(c, o)); if (!c) return 1; cudaSetDevice(c->cfg.device); c->seg_valid = false; return zero_info(c) || vctk_set_voxel_opacity(c, o); }
int vct_temporal_radiance_filter(vct_ctx* c, float d) { VCT_FAN(c, vct_temporal_radiance_filter(c, d)); if (!c) return 1; cudaSetDevice(c->cfg.device); c->seg_valid = false; return vctk_temporal_radiance_filter(c, d); }
int vct_filter3d(vct_ctx* c, int which, int src_level) { VCT_FAN(c, vct_filter3d(c, which, src_level)); if (!c) return 1; cudaSetDevice(c->cfg.device); c->seg_valid = false; return vctk_filter3d(c, which, src_level); }
int vct_normalize_voxels_f16(vct_ctx* c, void* col, void* nrm, float o) { VCT_NO_GROUP(c, "vct_normalize_voxels_f16"); if (!c) return 1; cudaSetDevice(c->cfg.device); c->seg_valid = false; return zero_info(c) || vctk_normalize_voxels_f16(c, col, nrm, o); }

// ------------------------------------------------------------------------------------------ outputs
static int volume_ptr(vct_ctx* c, int which, int level, void** ptr, size_t* bytes) {
    const int N = VCT_WARP_DIM;
    switch (which) {
        case VCT_VOL_COLOR: case VCT_VOL_RADIANCE: {
            if (level < 0 || level >= c->L) return fail(c, "volume level out of range");
            const size_t d = level_dim(c->D, level);
            *ptr = (which == VCT_VOL_COLOR ? c->d_color : c->d_radiance) + c->level_off[level]; *bytes = d * d * d * 4; return 0;
        }
        case VCT_VOL_NORMAL: if (level) return fail(c, "voxelNormal has one level"); *ptr = c->d_normal; *bytes = (size_t)c->D * c->D * c->D * 4; return 0;
        case VCT_VOL_OCCUPANCY: *ptr = c->d_occ; *bytes = N * N * N * 4; return 0;
        case VCT_VOL_WARPMAP: *ptr = c->d_warpmap; *bytes = N * N * N * 8; return 0;
        case VCT_VOL_WARP_WEIGHTS_LOW: *ptr = c->d_wlo; *bytes = N * N * N * 8; return 0;
        case VCT_VOL_WARP_WEIGHTS_HIGH: *ptr = c->d_whi; *bytes = N * N * N * 8; return 0;
        case VCT_BUF_IMAGE: *ptr = c->image_of(c->image_parity); *bytes = vctk_image_rows(c) * (size_t)c->W * 4; return 0;   // the image of the last trace
    }
    return fail(c, "unknown volume");
}
int vct_sync(vct_ctx* c) { VCT_FAN(c, vct_sync(c)); if (!c) return 1; cudaSetDevice(c->cfg.device); VCT_CHECK(c, cudaStreamSynchronize(c->stream)); return 0; }
int vct_read_volume(vct_ctx* c, int which, int level, void* out) {
    if (!c || !out) return 1;
    if (!c->group.empty() && (which == VCT_VOL_COLOR || which == VCT_VOL_NORMAL || which == VCT_VOL_RADIANCE)) {
        // multi-device handle: every rank holds its z layers (slab or stripes) of a volume (the traced pyramid is complete everywhere after a frame, the
        // others are not): assemble the level from the ranks' slabs
        void* p0; size_t bytes; if (volume_ptr(c, which, level, &p0, &bytes)) return 1;
        const int ws = (int)c->group.size() + 1, d = level_dim(c->D, level);
        // levels above the ranks' stripes exist (whole) only in the traced pyramid, computed on every rank after the exchange: rank 0's copy
        if (level > vctk_mip_top_sharded_level(c)) { cudaSetDevice(c->cfg.device); VCT_CHECK(c, cudaMemcpyAsync(out, p0, bytes, cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream)); return 0; }
        const size_t run = (size_t)d * d * (c->st.T >> level) * 4;          // bytes of one stripe at this level
        for (int r = 0; r < ws; ++r) {
            vct_ctx* m = r ? c->group[r - 1] : c;
            void* pm; size_t bm; if (volume_ptr(m, which, level, &pm, &bm)) return fail(c, m->error.c_str());
            cudaSetDevice(m->cfg.device);
            for (int k = 0; k < m->st.count; ++k) {
                const size_t off = (size_t)d * d * (stripe_z(m->st, k) >> level) * 4;
                VCT_CHECK(c, cudaMemcpyAsync((char*)out + off, (char*)pm + off, run, cudaMemcpyDeviceToHost, m->stream));
            }
            VCT_CHECK(c, cudaStreamSynchronize(m->stream));
        }
        return 0;
    }
    cudaSetDevice(c->cfg.device);
    void* p; size_t b; if (volume_ptr(c, which, level, &p, &b)) return 1;
    VCT_CHECK(c, cudaMemcpyAsync(out, p, b, cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vct_write_volume(vct_ctx* c, int which, int level, const void* in) { VCT_FAN(c, vct_write_volume(c, which, level, in));
    if (!c || !in) return 1;
    cudaSetDevice(c->cfg.device);
    void* p; size_t b; if (volume_ptr(c, which, level, &p, &b)) return 1;
    c->seg_valid = false;
    VCT_CHECK(c, cudaMemcpyAsync(p, in, b, cudaMemcpyHostToDevice, c->stream));
    if (which == VCT_VOL_WARPMAP && vctk_warpmap_floats(c)) return 1;
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vct_read_image(vct_ctx* c, void* rgba8) {
    if (!c || !rgba8) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaMemcpyAsync(rgba8, c->image_of(c->image_parity), (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
// Pipelined read-back: the copy of this frame's image runs on a second stream while the library stream goes on with the next
// frame's voxel passes; only the next cone trace (the next writer of the image) waits for it.
int vct_read_image_async(vct_ctx* c, void* pinned_rgba8) { VCT_NO_GROUP(c, "vct_read_image_async");
    if (!c || !pinned_rgba8) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaEventRecord(c->ev_image_ready, c->stream));
    VCT_CHECK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_image_ready, 0));
    const int par = c->image_parity;
    VCT_CHECK(c, cudaMemcpyAsync(pinned_rgba8, c->image_of(par), (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->copy_stream));
    VCT_CHECK(c, cudaEventRecord(c->ev_copy_done[par], c->copy_stream));
    c->copy_pending[par] = true;
    return 0;
}
int vct_read_image_wait(vct_ctx* c, int block_host) {
    if (!c) return 1;
    cudaSetDevice(c->cfg.device);
    for (int par = 0; par < 2; ++par) if (c->copy_pending[par]) { VCT_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_copy_done[par], 0)); c->copy_pending[par] = false; }
    if (block_host) VCT_CHECK(c, cudaStreamSynchronize(c->copy_stream));
    return 0;
}
int vct_read_shadowmap(vct_ctx* c, float* d) {
    if (!c || !d) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaMemcpyAsync(d, c->d_shadow, (size_t)c->S * c->S * 4, cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vct_write_shadowmap(vct_ctx* c, const float* d) { VCT_FAN(c, vct_write_shadowmap(c, d));
    if (!c || !d) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaMemcpyAsync(c->d_shadow, d, (size_t)c->S * c->S * 4, cudaMemcpyHostToDevice, c->stream));
    if (vctk_shadow_minmax(c)) return 1;
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vct_read_visibility(vct_ctx* c, uint64_t* v) {
    if (!c || !v) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaMemcpyAsync(v, c->d_vis, (size_t)c->W * c->H * 8, cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}
static int fetch_counters(vct_ctx* c) {
    VCT_CHECK(c, cudaMemcpyAsync(&c->h_counters, c->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream)); VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    if (c->h_counters.overflow) return fail(c, "a fixed-capacity device buffer overflowed (fragment buffer or raster tile queue): raise vct_config.max_fragments");
    return 0;
}
int vct_get_counters(vct_ctx* c, vct_voxelize_info* info) {
    if (!c || !info) return 1;
    cudaSetDevice(c->cfg.device);
    if (fetch_counters(c)) return 1;
    info->total_fragments = c->h_counters.total_fragments; info->unique_voxels = c->h_counters.unique_voxels; info->max_fragments_per_voxel = c->h_counters.max_fragments_per_voxel;
    for (vct_ctx* m : c->group) {                              // multi-device handle: a voxel's fragments all land on the rank that owns its slab
        vct_voxelize_info mi;
        if (vct_get_counters(m, &mi)) return fail(c, m->error.c_str());
        info->total_fragments += mi.total_fragments; info->unique_voxels += mi.unique_voxels; info->max_fragments_per_voxel = std::max(info->max_fragments_per_voxel, mi.max_fragments_per_voxel);
    }
    return 0;
}
int vct_get_cone_steps(vct_ctx* c, unsigned long long* steps) {
    if (!c || !steps) return 1;
    cudaSetDevice(c->cfg.device);
    if (fetch_counters(c)) return 1;
    *steps = c->h_counters.cone_steps;
    for (vct_ctx* m : c->group) { unsigned long long ms = 0; if (vct_get_cone_steps(m, &ms)) return fail(c, m->error.c_str()); *steps += ms; }
    return 0;
}
int vct_get_timings(vct_ctx* c, vct_timings* t) {
    if (!c || !t) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    auto ms = [&](int a, int b) { float v = 0.f; if (cudaEventElapsedTime(&v, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.0; } return (double)v * 1e6; };
    std::memset(t, 0, sizeof *t);
    t->shadowmap_ns = ms(EV_START, EV_SHADOW); t->warpmap_ns = ms(EV_SHADOW, EV_WARP); t->clear_ns = ms(EV_WARP, EV_CLEAR);
    t->voxelize_ns = ms(EV_WARP, EV_VOXEL); t->transfer_ns = ms(EV_VOXEL, EV_TRANSFER); t->radiance_ns = ms(EV_TRANSFER, EV_INJECT);
    t->mipmap_ns = ms(EV_INJECT, EV_MIP); t->exchange_ns = ms(EV_MIP, EV_XCHG); t->gbuffer_ns = ms(EV_XCHG, EV_GBUF); t->render_ns = ms(EV_GBUF, EV_TRACE); t->total_ns = ms(EV_START, EV_TRACE);
    return 0;
}
// A raw pointer lets the caller write the volume at any time: sparse frames are switched off for good (until vct_remake).
void* vct_device_ptr(vct_ctx* c, int which, int level) {
    if (!c) return nullptr;
    void* p; size_t b;
    if (volume_ptr(c, which, level, &p, &b)) return nullptr;
    // (multi-GPU: the caller all-gathers the remote slabs into the levels, which the own slab's masks do not describe anyway)
    if (c->cfg.world_size <= 1 && (which == VCT_VOL_COLOR || which == VCT_VOL_NORMAL || which == VCT_VOL_RADIANCE)) { c->seg_disabled = true; c->seg_valid = false; }
    return p;
}
size_t vct_level_bytes(vct_ctx* c, int which, int level) { if (!c) return 0; void* p; size_t b; return volume_ptr(c, which, level, &p, &b) ? 0 : b; }
void* vct_stream(vct_ctx* c) { return c ? (void*)c->stream : nullptr; }
// Enqueue on a caller-owned stream (e.g. the stream torch.distributed's NCCL collectives run on) so that the
// slab exchange needs no host synchronisation.  NULL restores a private stream.
int vct_set_stream(vct_ctx* c, void* stream) { VCT_NO_GROUP(c, "vct_set_stream");
    if (!c) return 1;
    cudaSetDevice(c->cfg.device);
    VCT_CHECK(c, cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { VCT_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    return 0;
}
int vct_set_profiling(vct_ctx* c, int level) { VCT_FAN(c, vct_set_profiling(c, level));
    if (!c) return 1;
    if (level < 0 || level > 2) return fail(c, "vct_set_profiling: level 0 (off), 1 (per pass) or 2 (per kernel)");
    c->profiling = level; c->prof_marks.clear(); c->prof_used = 0;
    return 0;
}
// per-kernel device time of the calls since the last pass entry point (profiling level 2); returns the count
int vct_get_kernel_times(vct_ctx* c, vct_kernel_time* out, int max_entries) {
    if (!c || !out || max_entries < 1) return -1;
    cudaSetDevice(c->cfg.device);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    int n = 0;
    for (size_t i = 1; i < c->prof_marks.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->prof_marks[i - 1].second, c->prof_marks[i].second) != cudaSuccess) { cudaGetLastError(); continue; }
        const char* name = c->prof_marks[i].first;
        int k = 0;
        while (k < n && std::strncmp(out[k].name, name, sizeof out[k].name - 1)) k++;
        if (k == n) { if (n == max_entries) continue; std::memset(&out[n], 0, sizeof out[n]); std::strncpy(out[n].name, name, sizeof out[n].name - 1); n++; }
        out[k].ns += (double)ms * 1e6; out[k].launches++;
    }
    return n;
}
unsigned long long vct_launch_count(vct_ctx* c, int reset) {
    if (!c) return 0;
    unsigned long long n = c->launches; if (reset) c->launches = 0;
    for (vct_ctx* m : c->group) n += vct_launch_count(m, reset);
    return n;
}

}  // extern "C"
