/* vct_b200.h — C ABI of the B200-native per-frame GI pipeline (drop-in for the GL dispatch blocks of
 * sfreed141/vct's Application::render, reference src/Application.cpp:196-1085).
 *
 * The reference has no plugin/FFI interface; the seam is the set of GL pass blocks inside render().  Each
 * export below replaces exactly one of those blocks (file:line cited per function) and takes the same
 * "uniforms" the block uploads, as one POD struct.  Conventions kept from the reference:
 *   - matrices are column-major float[16] exactly as GLM stores them and glUniformMatrix4fv uploads them;
 *   - the host computes every matrix uniform (projection, view, ls, inverse(ls), mvp_x/y/z, pv), as the
 *     reference does on the CPU (src/Application.cpp:200-210, 689-692, 804);
 *   - vertices are the reference's 56-byte `Vertex` (src/Graphics/Mesh.h:72-76), indices are uint32 in draw
 *     order (per-material lists concatenated, src/Graphics/Mesh.cpp:340-371);
 *   - lights are the 80-byte std140 records of src/Scene.cpp:64-76;
 *   - never throws, never aborts: every call returns 0 on success, non-zero on error, and
 *     vct_last_error() returns the message (reference logs and carries on, src/log.h:29-47);
 *   - a context is NOT thread-safe: one host thread, in-order, like the single GL context.
 * All device work is enqueued on the context's CUDA stream; calls that return data to the host synchronise.
 * There is no CPU fallback: without a CUDA device vct_create fails.
 */
#ifndef VCT_B200_H
#define VCT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCT_WARP_DIM 32            /* reference src/Application.h:178, shaders/generateWarpmap.geom:6 */
#define VCT_MAX_LEVELS 12

typedef struct vct_ctx vct_ctx;    /* opaque */

/* Resources created once — reference `VCT` ctor + Application::init (src/Application.h:107-156,
 * src/Application.cpp:37-70). */
typedef struct {
    int dim;            /* voxelDim (default 256)                          Application.h:133 */
    int levels;         /* voxelLevels (default 6), clamped to [1, log2(dim)+1] like VCT::remake */
    int shadow_size;    /* SHADOWMAP_WIDTH/HEIGHT (4096)                   Application.cpp:30-31 */
    int width, height;  /* frame size                                      common.h:9-10 */
    int device;         /* CUDA device ordinal */
    /* z sharding of the voxel volume across ranks (slab or stripes: slab_stripe below); single GPU: rank 0 of 1. */
    int rank, world_size;
    int max_fragments;  /* capacity of the voxel-fragment buffer (0 = default 8 Mi) */
    /* Single-process multi-GPU (SURVEY §8b "multi-GPU fan-out is internal"): n_devices > 1 makes vct_create build one context per
     * listed device (rank i on devices[i]; NULL = devices 0..n-1), enable peer access, attach the slab exchange, and return a handle
     * that fans every call out; rank / world_size / device above are then ignored.  0 or 1: one context on `device`. */
    int n_devices;
    const int* devices;
    /* How z is dealt to the ranks: rank of voxel layer z = (z / slab_stripe) mod world_size.  A power of two >= 16 that divides
     * dim / world_size interleaves stripes, which balances scenes that fill only part of the volume (Sponza occupies the middle half of z: with
     * contiguous slabs half of 8 ranks have nothing to voxelise).  -1 = one contiguous slab of dim / world_size layers per rank.  0 = the
     * library chooses: environment variable VCT_SLAB_STRIPE if set, else 16 from 4 ranks on, else contiguous.  A stripe that does not fit the
     * volume falls back to contiguous slabs; vct_slab_stripe() reports what is in effect. */
    int slab_stripe;
} vct_config;

/* 80-byte std140 light record — reference src/Scene.cpp:64-76, shaders/voxelize.frag:31-45. */
typedef struct {
    float position[3];  float _pad0;
    float direction[3]; float _pad1;
    float color[3];
    float range;
    float intensity;
    int   enabled, selected, shadow_caster;
    unsigned type;      /* 0 = point, 1 = directional */
    float _pad2[3];
} vct_light;

/* Material as Mesh::draw binds it — reference src/Graphics/Mesh.cpp:326-360, Mesh.h:45-57.
 * Texture ids are the ones given to vct_upload_texture; -1 = no map (GL handle 0). */
typedef struct {
    int   diffuse_tex, specular_tex, normal_tex, roughness_tex, metallic_tex, alpha_tex;
    float shininess;
    float diffuse[3];   /* always (0,0,0) in the reference (self-assignment bug, Mesh.h:34-43) */
} vct_material;

/* VoxelizeInfo SSBO — reference src/Application.h:195-198. */
typedef struct { unsigned total_fragments, unique_voxels, max_fragments_per_voxel; } vct_voxelize_info;

/* Cone settings — reference `VCTSettings`, src/Application.h:28-34, defaults :96-97. */
typedef struct { int steps; float cone_angle, bias, cone_initial_height, lod_offset; } vct_cone_settings;

/* Everything Application::render uploads as uniforms in one frame: `Settings` (src/Application.h:36-103),
 * `VCT{min,max,center}` (:139-140) and the per-frame matrices. */
typedef struct {
    /* matrices (column-major) */
    float projection[16];   /* perspective(fov, aspect, near, far)        Application.cpp:200 */
    float view[16];         /* camera.lookAt()                            :201 */
    float pv[16];           /* perspective(fov, aspect, 1, 20) * view     :202 */
    float lp[16], lv[16];   /* light ortho / lookAt                       :208-209 */
    float ls[16];           /* lp * lv                                    :210 */
    float ls_inverse[16];   /* inverse(ls)                                :804 */
    float mvp_x[16], mvp_y[16], mvp_z[16];   /* voxelisation views       :689-692 */
    float eye[3];           /* camera.position */
    float voxel_min[3], voxel_max[3], voxel_center[3];   /* vct.min/max/center (symmetric cube) */
    float clear_color[3];   /* glClearColor                               :41 */
    /* voxelisation — Settings */
    int   voxelize_lighting;      /* true  */
    int   voxelize_atomic_max;    /* reference default true; parity mode false */
    int   axis_override;          /* -1 */
    int   deterministic;          /* this build: 1 = apply running average in canonical draw order
                                     (bit-reproducible), 0 = free-running CAS like the GLSL */
    float voxel_set_opacity;      /* 0.5 */
    int   temporal_filter_radiance; float temporal_decay;   /* false, 0.8 */
    int   radiance_lighting;      /* false */
    int   radiance_dilate;        /* false (malformed in the reference; rejected if set) */
    int   voxel_fill_holes;       /* false */
    int   mip_color_chain;        /* 1 = also filter voxelColor like the reference (:903-917) */
    /* warp modes — common.glsl:44-60 */
    int   warp_voxels, warp_texture, warp_texture_linear, warp_texture_axes[3];
    int   use_warpmap_weights_texture;   /* true */
    float warp_texture_high_resolution, warp_texture_low_resolution;   /* 2.0, 0.5 */
    /* shading — phong.frag uniforms */
    int   draw_radiance, draw_occlusion, cooktorrance, enable_postprocess, enable_normal_map;
    int   enable_indirect, enable_diffuse, enable_specular, enable_reflections;
    float ambient_scale, reflect_scale;
    vct_cone_settings diffuse_cone, specular_cone;
    int   specular_cone_angle_from_roughness;
    /* debug views of phong.frag (Settings::drawVoxels/drawNormals/drawDominantAxis/debug*, Application.cpp:992-1005;
     * Overlay.cpp toggles): one VCT_VIEW_* value instead of nine booleans, in the shader's own priority order. */
    int   debug_view;       /* VCT_VIEW_SHADED (0) = the normal frame */
    float miplevel;         /* Settings::miplevel: lod of VCT_VIEW_VOXELS (phong.frag:350-401) */
    /* Settings::voxelizeTesselation (Application.h:85, reference default TRUE; parity mode false = the raster path):
     * patches through testTesselation.tesc/.tese instead of voxelize.vert/geom/frag (Application.cpp:585-665).  Unlit albedo,
     * atomicMax (voxelize_atomic_max, bit-reproducible) or the free-running running average (order-dependent like the GLSL). */
    int   voxelize_tesselation;
    /* Settings::voxelizeTesselationWarp (Application.h:102, default false): the last-but-one entry of common.glsl's mapping
     * priority (warpVoxels > warpTexture > this > linear, common.glsl:44-60): the voxel grid becomes the camera frustum,
     * position = (pv * P).xyz / w * 0.5 + 0.5 (common.glsl:37-42).  Used by testTesselation.tese, injectRadiance.comp and
     * phong.frag (each cone sample goes back to world space and through pv, phong.frag:158-162); voxelize.frag declares the
     * uniform but never reads it, so the raster voxeliser ignores it — like the reference. */
    int   voxelize_tesselation_warp;
    /* Settings::conservativeRasterization (Application.h:61-62; reference default MSAA, parity mode OFF): how the two voxelisation
     * passes (occupancy :244-249, voxelise :673-678) decide coverage.  VCT_RASTER_CENTER = OFF: a fragment where the pixel centre is
     * covered.  VCT_RASTER_MSAA: GL_MULTISAMPLE on the 4-sample window framebuffer (main.cpp:256) — a fragment where ANY sample is
     * covered by the near/far-clipped triangle (OpenGL 4.5 section 14.6.6), its inputs interpolated at the pixel centre (no `centroid`
     * in voxelize.geom/.frag: extrapolated when the centre itself is outside).  GL leaves the sample positions to the driver
     * (glGetMultisamplefv); msaa_samples holds them as (x, y) pairs in pixels, y up, multiples of 1/256 — all zero selects the
     * standard 4x pattern of NVIDIA's GL and D3D: (0.375, 0.125) (0.875, 0.375) (0.125, 0.625) (0.625, 0.875).  The NV mode
     * (GL_CONSERVATIVE_RASTERIZATION_NV) is not built. */
    int   conservative_raster;
    float msaa_samples[8];
    /* Settings::voxelizeMultiplier (Application.h:87, default 1; the overlay's slider goes from 0.5 to 4): the voxelise pass rasterises into a
     * viewport of (int)(m * dim) pixels squared (Application.cpp:668) while the voxel index still comes from the interpolated position times
     * dim — m = 2 gives every voxel four times the fragments.  0 reads as 1.  (The occupancy pass keeps its 32^2 viewport, :239.  Not
     * replicated: the reference draws into the WINDOW's framebuffer, which cuts a viewport taller than the window.) */
    float voxelize_multiplier;
} vct_frame_params;
enum { VCT_RASTER_CENTER = 0, VCT_RASTER_MSAA = 1 };

enum { VCT_VIEW_SHADED = 0,
       VCT_VIEW_VOXELS = 1,            /* `voxelize`: the traced volume sampled at the fragment's voxel, lod = miplevel   phong.frag:347-404 */
       VCT_VIEW_MATERIAL_DIFFUSE = 2,  /* debugMaterialDiffuse / Roughness / Metallic                                    :405-425 */
       VCT_VIEW_MATERIAL_ROUGHNESS = 3,
       VCT_VIEW_MATERIAL_METALLIC = 4,
       VCT_VIEW_NORMALS = 5,           /* `normals`: the shading normal after normal mapping                              :441-443 */
       VCT_VIEW_DOMINANT_AXIS = 6,     /* `dominant_axis`                                                                 :444-447 */
       VCT_VIEW_INDIRECT = 7,          /* debugIndirect: the six diffuse cones (x occlusion if draw_occlusion)            :489 */
       VCT_VIEW_OCCLUSION = 8,         /* debugOcclusion                                                                  :490 */
       VCT_VIEW_REFLECTIONS = 9,       /* debugReflections: the specular cone                                             :505 */
       /* sub-views of `voxelize` (phong.frag:350-356), tested by the shader before the radiance / colour lookup of VCT_VIEW_VOXELS */
       VCT_VIEW_VOXEL_NORMALS = 10,    /* voxelize && normals: voxelNormal (one level, NEAREST) at the fragment's voxel          :350-353 */
       VCT_VIEW_WARP_TEXTURE = 11,     /* voxelize && debugWarpTexture: texture(warpmap, linear position).xyz                    :354-357 */
       VCT_VIEW_WARP_TEXTURE_TC = 12,  /* ... with `toggle`: the linear position itself                                          :357 */
       /* (drawWarpSlope, :360-389, is not built: its colour lookup indexes colors[min(ceil(slope), colors.length())] — out of bounds
        * whenever the slope reaches the table's length, undefined in GLSL) */
       VCT_VIEW_LAST = VCT_VIEW_WARP_TEXTURE_TC };

/* GLBufferedTimer results in ns, same names as reference src/Application.h:192 (+ producers). */
typedef struct {
    double voxelize_ns, shadowmap_ns, radiance_ns, mipmap_ns, render_ns, total_ns;
    double transfer_ns, gbuffer_ns, warpmap_ns, clear_ns, exchange_ns;
} vct_timings;

typedef struct { char name[40]; double ns; unsigned launches; unsigned _pad; } vct_kernel_time;

enum { VCT_VOL_COLOR = 0, VCT_VOL_NORMAL = 1, VCT_VOL_RADIANCE = 2, VCT_VOL_OCCUPANCY = 3, VCT_VOL_WARPMAP = 4,
       VCT_VOL_WARP_WEIGHTS_LOW = 5, VCT_VOL_WARP_WEIGHTS_HIGH = 6,
       VCT_BUF_IMAGE = 7 /* device pointer only: RGBA8 rows of the LAST rendered image (the library keeps two images and alternates,
                            so that a read-back may overlap the next frame); rows padded to whole 64-row screen tiles */ };

/* ---- lifetime: VCT ctor/dtor/remake, Application::init --------------------------- Application.h:109-156 */
int  vct_create(const vct_config* cfg, vct_ctx** out);
int  vct_destroy(vct_ctx* ctx);
int  vct_remake(vct_ctx* ctx, int dim, int levels);
const char* vct_last_error(const vct_ctx* ctx);          /* ctx may be NULL: error of the last failed create */

/* ---- scene upload: Mesh VAO/EBO/texture creation ------------------ Mesh.cpp:208-266, GLHelper.cpp:165-211 */
int  vct_upload_mesh(vct_ctx*, int actor, const void* vertices, size_t n_vertices, size_t stride /*56*/,
                     const uint32_t* indices, size_t n_indices, const int32_t* material_of_triangle);
/* all mip levels packed back to back, level 0 first (the host generates them, like the DDS path
 * ResourceLoader.h:94-103; glGenerateTextureMipmap's filter is implementation-defined) */
int  vct_upload_texture(vct_ctx*, int tex, int width, int height, int channels, int levels, const void* pixels);
int  vct_set_material(vct_ctx*, int material, const vct_material*);
int  vct_set_actor_transform(vct_ctx*, int actor, const float model[16]);   /* Scene::draw "model" uniform, Scene.cpp:31-36 */
int  vct_set_lights(vct_ctx*, const vct_light* lights, int n);              /* Scene::bindLightSSBO, Scene.cpp:58-62 */

/* ---- one export per GL pass block of Application::render, same order ------------------------------------ */
int  vct_shadowmap(vct_ctx*, const vct_frame_params*);     /* Application.cpp:212-233 */
int  vct_occupancy(vct_ctx*, const vct_frame_params*);     /* :235-301  (32^3, imageAtomicOr) */
int  vct_warpmap(vct_ctx*, const vct_frame_params*);       /* :303-577  (prefix sums + weights + warpmap, on device) */
int  vct_voxelize(vct_ctx*, const vct_frame_params*);      /* :581-755  (clear + raster path) */
int  vct_transfer(vct_ctx*, const vct_frame_params*);      /* :757-785 */
int  vct_inject(vct_ctx*, const vct_frame_params*);        /* :787-837 */
int  vct_fill_holes(vct_ctx*, const vct_frame_params*);    /* :839-875 */
int  vct_mip(vct_ctx*, int which_volume);                  /* :877-921  (VCT_VOL_RADIANCE or VCT_VOL_COLOR) */
int  vct_mip_kernel(vct_ctx*, int which_volume, int kernel_mode);   /* shaders/filterRadiance.comp:9-60 with its `kernelMode` uniform:
                                                              0 BOX2 (= vct_mip), 1 BOX3, 2 CUBE; the host never sets it (dead modes) */
int  vct_exchange(vct_ctx*);                               /* multi-GPU only: publish slab pyramid to the 3D texture
                                                              after the caller's all-gather */
/* Multi-GPU (z sharding): once every rank knows its peers' buffers, vct_frame / vct_gi_passes run the WHOLE sharded
 * frame — voxel passes on the own slab, the slab exchange over NVLink peer memory (exchange.cu: level 0 as the flagged x-row segments
 * pushed into staging regions in every peer's memory, levels >= 1 stored straight into the peers' pyramids, device-side flags instead of
 * collectives), the cone trace of the own screen tiles, pixels stored into rank 0's image — with no collective and no host
 * synchronisation by the caller; rank 0's image is complete when its stream reaches the end of the call.
 * Setup once, either way:
 *   one process per GPU   vct_exchange_setup on every rank; hand every rank's VCT_EXCHANGE_HANDLE_BYTES blob (vct_exchange_export:
 *                         cudaIpc handles) to every other rank (vct_exchange_import)
 *   one process, N GPUs   vct_config.n_devices > 1 does all of it inside vct_create (peer access + vct_exchange_attach)
 * Without attached peers a world_size > 1 context stops after its mip chains (the caller all-gathers the levels itself, then
 * vct_exchange + vct_cone_trace: the round-1 protocol, kept for transports other than peer memory). */
#define VCT_EXCHANGE_HANDLE_BYTES 384
typedef struct { void* staging; void* radiance; void* color; void* image; void* shadow; } vct_peer;   /* device pointers valid on THIS rank's device */
int  vct_exchange_setup(vct_ctx*);
int  vct_exchange_export(vct_ctx*, void* handle /* VCT_EXCHANGE_HANDLE_BYTES */);
int  vct_exchange_import(vct_ctx*, int rank, const void* handle);
int  vct_exchange_local(vct_ctx*, vct_peer* out);          /* this rank's own buffers (to attach on peers of the same process) */
int  vct_exchange_attach(vct_ctx*, int rank, const vct_peer*);
int  vct_frame_was_sparse(vct_ctx*);                       /* 1: the last vct_frame / vct_gi_passes visited flagged segments only */
int  vct_slab_stripe(vct_ctx*);                            /* layers per stripe in effect (dim for one GPU) */
int  vct_mask_parity(vct_ctx*);                            /* which of the two segment masks the NEXT frame writes (0/1): a CUDA graph that
                                                              captured frames must be replayed at the parity it was captured at */
int  vct_gbuffer(vct_ctx*, const vct_frame_params*);       /* :936-965  depth prepass -> visibility buffer */
int  vct_cone_trace(vct_ctx*, const vct_frame_params*);    /* :967-1067 phong + cone tracing */
int  vct_frame(vct_ctx*, const vct_frame_params*);         /* the whole graph, in reference order */
int  vct_gi_passes(vct_ctx*, const vct_frame_params*);     /* the BASELINE metric's passes only:
                                                              voxelize+transfer+inject(+fill)+mip+cone trace */

/* ---- dead-shader equivalents (kernel parity only; the reference host never dispatches them) ------------- */
int  vct_set_voxel_opacity(vct_ctx*, float opacity);       /* shaders/setVoxelOpacity.comp:18-35 */
int  vct_temporal_radiance_filter(vct_ctx*, float decay);  /* shaders/temporalRadianceFilter.comp:9-19 */
int  vct_filter3d(vct_ctx*, int which_volume, int src_level);  /* shaders/filter3d.comp:16-47 */
int  vct_normalize_voxels_f16(vct_ctx*, void* color_rgba16f, void* normal_rgba16f, float opacity);
                                                           /* shaders/normalizeVoxels.comp:19-42 on caller-provided
                                                              DEVICE buffers of dim^3 half4 (USE_RGBA16F build) */

/* ---- outputs ------------------------------------------------------------------------------------------- */
/* Application::debugVoxels (src/Application.cpp:1222-1275, shaders/debugVoxels.vert/.geom/.frag; Settings::debugVoxels replaces the frame's
 * render passes, :923-928): every voxel of the base grid whose colour — the traced pyramid sampled at its centre with lod = miplevel — has
 * alpha > 0 is drawn as a cube of one voxel's size with depth test and back-face culling, in that colour, over the clear colour; uses
 * projection, view, miplevel, draw_radiance, clear_color and the volume extents of `p`.  The result is the current image (vct_read_image).
 * One GPU, dim <= 512.  The vertex shader's float(gl_InstanceID) arithmetic is reproduced (ids collapse beyond 2^24 instances, i.e. at 512^3). */
int  vct_debug_voxels(vct_ctx*, const vct_frame_params*);
int  vct_read_image(vct_ctx*, void* rgba8 /* width*height*4, row 0 = bottom like glReadPixels */);
/* Pipelined read-back for render loops (the reference's glfwSwapBuffers never blocks on the frame either, main.cpp:229-240):
 * enqueue the copy of the current image into PINNED host memory on the library's copy stream and return at once; the library
 * stream carries on with the next frame and only the next cone trace waits for the copy.  vct_read_image_wait orders the
 * library stream behind the last copy (block_host != 0: also blocks the host until the pixels are in `pinned_rgba8`).
 * Not for use while the library stream is being captured into a CUDA graph. */
int  vct_read_image_async(vct_ctx*, void* pinned_rgba8);
int  vct_read_image_wait(vct_ctx*, int block_host);
int  vct_read_volume(vct_ctx*, int which, int level, void* out);   /* RGBA8 words / u32 occupancy / u16x4 warpmap / f16x4 */
/* (multi-device handle: a level is assembled from the devices' own z layers; the levels above the sharded ones — see vct_slab_stripe — exist whole
 * only in the TRACED pyramid, computed on every device after the exchange, and are read from device 0) */
int  vct_write_volume(vct_ctx*, int which, int level, const void* in);   /* test hook: seed a volume */
int  vct_read_shadowmap(vct_ctx*, float* depth /* S*S */);
int  vct_write_shadowmap(vct_ctx*, const float* depth);
int  vct_read_visibility(vct_ctx*, uint64_t* vis /* W*H: depth bits << 32 | ~draw index */);
int  vct_get_counters(vct_ctx*, vct_voxelize_info*);               /* Overlay.cpp:104-110 */
int  vct_get_timings(vct_ctx*, vct_timings*);                      /* GLTimer.h:49-87 */
int  vct_get_cone_steps(vct_ctx*, unsigned long long* steps);      /* texture fetches of the last cone trace */
int  vct_sync(vct_ctx*);

/* device pointers for zero-copy plumbing (torch.distributed all-gather of the pyramid, CUDA-GL interop) */
void*  vct_device_ptr(vct_ctx*, int which, int level);
size_t vct_level_bytes(vct_ctx*, int which, int level);
void*  vct_stream(vct_ctx*);                                        /* cudaStream_t */
int    vct_set_stream(vct_ctx*, void* cuda_stream);                 /* enqueue on a caller-owned stream (NULL: private) */
/* kernels of this library launched since the last reset (bench.py's gpu_launches) */
unsigned long long vct_launch_count(vct_ctx*, int reset);
/* GLTimer granularity (reference src/Graphics/GLTimer.h:6-87 brackets passes): 0 = no events, 1 = one event per
 * pass (default; feeds vct_get_timings), 2 = additionally one event per kernel (feeds vct_get_kernel_times). */
int    vct_set_profiling(vct_ctx*, int level);
int    vct_get_kernel_times(vct_ctx*, vct_kernel_time* out, int max_entries);   /* returns the entry count, <0 on error */

/* ---- scene ingest (host only, no CUDA call): what the reference does when it constructs a Mesh ------------------
 * Mesh::loadMesh (src/Graphics/Mesh.cpp:42-206: tinyobjloader OBJ/MTL -> de-duplicated Vertex array, one index list per
 * material, tangent frames, trailing "default" material), GLHelper::createTextureFromImage (src/Graphics/GLHelper.cpp:
 * 165-211: stb_image PNG decode -> R8/RGB8/RGBA8 + generated mips) and ResourceLoader::loadDDS (src/ResourceLoader.h:
 * 26-108: DXT1/DXT3/DXT5 with the file's mips).  Implemented in vct_b200/host/vct_ingest{,_image}.hpp; byte-for-byte
 * parity with the reference's own tinyobjloader and stb_image is tested in tests/test_ingest.py. */
typedef struct vct_ingest vct_ingest;    /* opaque: one loaded OBJ (or one decoded image) */
typedef struct {
    const float*    vertices;            /* n_vertices x 14 floats = the 56-byte Vertex of Mesh.h:72-76 */
    size_t          n_vertices;
    const uint32_t* indices;             /* draw order: material by material (Mesh.cpp:340-371) */
    size_t          n_indices;
    const int32_t*  material_of_triangle;
    int             n_materials;         /* the file's materials + the trailing default */
    int             n_textures;
    float           bounds_min[3], bounds_max[3], radius;   /* Mesh.cpp:127-128, 197-204 */
} vct_ingest_mesh;
typedef struct {
    int width, height, channels, levels; /* as vct_upload_texture takes them */
    const void* pixels;                  /* all levels packed, level 0 first */
    size_t bytes;
    const char* name;                    /* as written in the MTL ("@default_texture.png" for the default material) */
} vct_ingest_texture;
enum { VCT_INGEST_NO_TEXTURES = 1 /* parse names only, decode nothing */ };
/* resource_dir: where default_texture.png lives (RESOURCE_DIR, src/common.h:13-15).  Returns 0 and a handle, or 1 and
 * a handle that only carries the log (free it too); a missing MTL or texture is a log line, not an error. */
int  vct_ingest_obj(const char* obj_path, const char* resource_dir, int flags, vct_ingest** out);
int  vct_ingest_image(const char* path, int generate_mips, vct_ingest** out);   /* one PNG / DDS file as texture 0 */
void vct_ingest_free(vct_ingest*);
const char* vct_ingest_log(const vct_ingest*);
int  vct_ingest_get_mesh(const vct_ingest*, vct_ingest_mesh* out);
/* texture ids inside `out` are local to this ingest (0 .. n_textures-1, -1 = no map) */
int  vct_ingest_get_material(const vct_ingest*, int material, vct_material* out, const char** name);
int  vct_ingest_get_texture(const vct_ingest*, int texture, vct_ingest_texture* out);
/* Mesh VAO/EBO/texture creation for one actor: vct_upload_texture for every texture at texture_base + local id,
 * vct_set_material at material_base + local id (texture ids rebased), vct_upload_mesh with material ids rebased, then
 * vct_set_actor_transform(model) (NULL: identity). */
int  vct_ingest_upload(vct_ctx*, const vct_ingest*, int actor, int material_base, int texture_base, const float model[16]);

#ifdef __cplusplus
}
#endif
#endif /* VCT_B200_H */
