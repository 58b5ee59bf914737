/* vct_oracle.h — CPU ORACLE for the per-frame GI pipeline.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (vct_b200/) never links, imports or calls it.
 *
 * PINNING.  The reference (sfreed141/vct @ c5c763d) has no tests and no golden vectors, and no OpenGL stack exists in the
 * build container or on the GPU box (no libGL/EGL/Mesa, no GLM/GLFW), so its binary cannot run.  Its SOURCES for the hot
 * path can be compiled, though, and the oracle is pinned against them bit for bit (oracle/Makefile builds everything FROM
 * WHERE THE SOURCES LIE into oracle/_ref/; nothing of the reference is stored in this repository):
 *   - host C++: src/main.cpp:21-128 (the `#if 0` CPU warp example) and src/Application.cpp:311-370 (per-frame warp-map
 *     tables) -> _ref/{warp_rig,warpmap_cpu}; fixtures tests/golden/warp_rig_ref.json, warpmap_cpu_ref.npz;
 *     tests/test_oracle_kat.py checks orc_warp_rig / orc_warp_partials / orc_warp_weight_table against them;
 *   - GLSL: transferVoxels.comp, filterRadiance.comp (BOX2/BOX3/CUBE), voxelFillHoles.comp, injectRadiance.comp,
 *     setVoxelOpacity.comp, normalizeVoxels.comp, temporalRadianceFilter.comp, filter3d.comp, voxelize.frag, phong.frag (+ common.glsl) and
 *     generateWarpmapWeights.frag / generateWarpmap.frag, testTesselation.tesc / .tese
 *     are mapped to C++ SYNTAX by ref_rig/glsl2cpp.py (bodies untouched), compiled against ref_rig/glsl_shim.h into
 *     _ref/libvct_glsl_ref.so and run on the same seeded inputs as the orc_* functions (compute shaders: whole dispatches;
 *     fragment shaders: replayed on the fragment-stage inputs recorded by orc_voxelize_trace / orc_shade_trace);
 *     tests/test_glsl_ref.py requires identical voxel words, counters, RGBA8 pixels and cone-step counts, live and against
 *     the committed outputs tests/golden/glsl_ref.npz / glsl_ref_fragment.npz.
 * What this pins: every expression, branch, constant and evaluation order of the shaders.  What it cannot pin, because it
 * is the OpenGL IMPLEMENTATION's and not the reference's: rasteriser coverage/interpolation, texture filtering and LOD
 * selection, unorm conversion, built-in function precision.  Those follow the OpenGL 4.5 specification with the canonical
 * choices listed in DESIGN.md §2 ("Canonical GL semantics"), stated once in glsl_shim.h and once here.
 *   - vertex / geometry stages (voxelize.vert/.geom, simple.vert, phong.vert) are compiled too: the voxelisation path is
 *     identical; the camera / light clip transforms differ by association order only (the shaders multiply the matrices first,
 *     the passes here transform the vector step by step): <= 2e-7 relative, bounded by the test, see DESIGN.md section 2.
 * STILL UNPINNED (restated from the source, known-answer tests only, tests/golden/kat.json): dither.frag /
 * reflectiveShadowMap.frag (one alpha test each).
 *
 * POD parameter structs are shared with the product's public header (the boundary spec); no code is.
 */
#ifndef VCT_ORACLE_H
#define VCT_ORACLE_H
#include "../include/vct_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int width, height, channels, levels; long long offset[16]; } orc_texture;

typedef struct {
    const float* vertices;            /* n_vertices x 14 (pos3 nrm3 uv2 tan3 bitan3) — Mesh.h:72-76 */
    const int* vertex_actor;          /* n_vertices */
    int n_vertices;
    const unsigned* indices;          /* 3 x n_tris, global vertex ids, DRAW ORDER */
    const int* tri_material;          /* n_tris */
    int n_tris;
    const float* actor_model;         /* n_actors x 16, column-major */
    int n_actors;
    const vct_material* materials; int n_materials;
    const orc_texture* textures; int n_textures;
    const unsigned char* texels;      /* blob the texture offsets index */
    const vct_light* lights; int n_lights;
} orc_scene;

/* a0  shadow map — Application.cpp:212-233, simple.vert, reflectiveShadowMap.frag:34-38 */
void orc_shadowmap(const orc_scene*, const vct_frame_params*, int S, float* depth);
/* a1+a2 voxelise (raster path) — voxelize.vert/geom/frag.  warpmap = 32^3 x 4 u16 or NULL. */
void orc_voxelize(const orc_scene*, const vct_frame_params*, int D, const float* shadow, int S,
                  const unsigned short* warpmap, unsigned* color, unsigned* normal, vct_voxelize_info* info);
/* a1'/a2' tessellation voxeliser (reference default) — simpleTesselated.vert, testTesselation.tesc/.tese,
 * Application.cpp:585-665; canonical fixed-function tessellation stated next to the definition */
void orc_voxelize_tess(const orc_scene*, const vct_frame_params*, int D, unsigned* color, unsigned* normal, vct_voxelize_info* info);
void orc_voxelize_tess_trace(const orc_scene*, const vct_frame_params*, int D, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                             float* rec, long long rec_cap, long long* rec_count);
void orc_world_vertices(const orc_scene*, float* wpos3, float* wnrm3);
/* tests only: evaluate the camera / light clip transforms with the shaders' own association ((P*V)*M)*v instead of P*(V*(M*v)) */
void orc_set_literal_vertex_transforms(int on);
/* tests only: record the alpha-tested fragments of orc_shadowmap (pass 0) / orc_visibility (pass 1): 6 floats each — pass, material,
   u, v, rho^2 of the alpha map, kept; rec = NULL stops recording; the count keeps running past cap */
void orc_alpha_trace(float* rec6, long long cap);
long long orc_alpha_trace_count(void);
void orc_vertex_stage(const orc_scene*, const vct_frame_params*, float* voxel13, float* light4, float* cam4, float* phong16);
long long orc_tess_patch(const float* wpos9, const vct_frame_params*, int D, float* levels4, float* uvw, long long cap);
/* a9(1) occupancy voxelise at 32^3 — voxelize.frag:187-193 */
void orc_occupancy(const orc_scene*, const vct_frame_params*, unsigned* occ);
/* a9(2-4) warpmap — Application.cpp:303-370 + generateWarpmap{Weights}.frag.  outputs 32^3 x 4 */
void orc_warpmap(const unsigned* occ, const vct_frame_params*, unsigned short* warpmap,
                 unsigned short* weights_low_f16, unsigned short* weights_high_f16);
/* a3  transferVoxels.comp:29-70 (+ the radiance clear of Application.cpp:762-764) */
void orc_transfer(const vct_frame_params*, int D, unsigned* color, unsigned* radiance, vct_voxelize_info* info);
/* a5  injectRadiance.comp:40-103 */
void orc_inject(const vct_frame_params*, int D, const unsigned* color, const unsigned* normal, const float* shadow,
                int S, const unsigned short* warpmap, const float light_pos[3], const float light_int[3],
                unsigned* radiance);
/* a4  voxelFillHoles.comp:8-36 + copy back (Application.cpp:867-872) */
void orc_fill_holes(int D, unsigned* radiance);
/* a6  filterRadiance.comp:14-65; mode 0 BOX2, 1 BOX3, 2 CUBE.  src is Ds^3, dst is (Ds/2)^3 */
void orc_mip(int Ds, const unsigned* src, unsigned* dst, int mode);
/* dead variants */
void orc_set_voxel_opacity(int D, float opacity, unsigned* color, unsigned* radiance, vct_voxelize_info* info);
void orc_temporal_radiance_filter(int D, float decay, unsigned* vol);
void orc_filter3d(int Ds, const unsigned* src, unsigned* dst);
void orc_normalize_voxels_f16(int D, float opacity, unsigned short* color_f16, unsigned short* normal_f16,
                              unsigned* radiance, vct_voxelize_info* info);
/* a0' visibility — depth prepass + GL_EQUAL, Application.cpp:936-977, dither.frag */
void orc_visibility(const orc_scene*, const vct_frame_params*, int W, int H, unsigned long long* vis);
/* a7  phong.vert/frag incl. traceCone.  pyramids: levels packed back to back (level 0 first). */
void orc_shade(const orc_scene*, const vct_frame_params*, int W, int H, const unsigned long long* vis,
               int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr, const float* shadow, int S,
               const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps);

void orc_set_normal_volume(const unsigned* normal);   /* voxelNormal for VCT_VIEW_VOXEL_NORMALS; D^3 words, must outlive the shade call */
void orc_shade_rows(const orc_scene*, const vct_frame_params*, int W, int H, int y_lo, int y_hi, int y_stride,
                    const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                    const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps);

/* Fragment-stage taps (tests/test_glsl_ref.py): the same passes, additionally recording what the rasteriser hands the
 * fragment shader — 16 floats per voxel fragment in canonical order, 28 floats per pixel (layouts next to the definitions). */
void orc_voxelize_trace(const orc_scene*, const vct_frame_params*, int D, const float* shadow, int S,
                        const unsigned short* warpmap, unsigned* color, unsigned* normal, vct_voxelize_info* info,
                        float* frag_rec, long long frag_cap, long long* frag_count);
void orc_shade_trace(const orc_scene*, const vct_frame_params*, int W, int H, const unsigned long long* vis, int D, int L,
                     const unsigned* radiance_pyr, const unsigned* color_pyr, const float* shadow, int S,
                     const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps, float* frag_rec);

void orc_shade_trace_rows(const orc_scene*, const vct_frame_params*, int W, int H, int y_lo, int y_hi, int y_stride,
                          const unsigned long long* vis, int D, int L, const unsigned* radiance_pyr, const unsigned* color_pyr,
                          const float* shadow, int S, const unsigned short* warpmap, unsigned* image, unsigned long long* cone_steps, float* frag_rec);
/* In the three *_trace functions the output volume / image may be NULL: only the fixed-function stage runs and records. */

/* N3  Application::debugVoxels (Application.cpp:1222-1275; debugVoxels.vert/.geom/.frag): the non-empty voxels of the base grid as depth-tested cubes */
void orc_debug_voxels(const vct_frame_params*, int W, int H, int D, int L, const unsigned* pyramid, unsigned* image);
void orc_debug_voxel_color(const vct_frame_params*, int D, int L, const unsigned* pyramid, unsigned instance, float* rgba);
void orc_debug_voxel_vertices(const vct_frame_params*, int D, unsigned instance, float* world3, float* clip21x4);   /* vertex + geometry stage of one instance */

/* KAT helpers */
unsigned orc_rgba8_avg(unsigned stored, float r, float g, float b);     /* voxelize.frag:111-139, one insertion */
unsigned orc_pack_unorm4x8(float r, float g, float b, float a);
void orc_warp_weight_table(int dim, float high, float low, float* low_out, float* high_out); /* Application.cpp:346-370 */
void orc_warp_rig(int n, const float* cells, float fixed_low, float high, float low, const float* tc, int npts,
                  float* out, int* part_x, int* part_y, float* wl, float* wh);                 /* src/main.cpp:20-127 */
float orc_cone_trace_const(int D, int L, unsigned voxel_word, const vct_cone_settings* cs, int* steps_out);
void orc_warp_partials(const unsigned* occ, int* xyz);                                          /* Application.cpp:311-343 */
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
