// vct_host.hpp — C++ host side above the C ABI: the reference's `Settings` / `VCT` / `Application::render` for the
// GI hot path with every OpenGL dispatch block replaced by the vct_* call that stands in for it.
//
// This is what a maintainer's patched src/Application.cpp looks like (INTEGRATION.md shows the diff): the same
// members (`vct`, `settings`, the six GLBufferedTimer names), the same per-frame matrix set-up
// (reference src/Application.cpp:196-210, 689-692, 804), the same pass order and the same `settings` guards around
// each pass (:212-1067).  It is header-only C++17 with no dependency beyond include/vct_b200.h; GLM is restated by
// the few functions the path uses (right-handed, depth -1..1, column-major, like GLM 0.9.9).
// Errors follow the reference's convention (src/log.h:29-47): log to stderr and carry on, never throw, never abort;
// `render` returns false if any pass reported an error.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vct_b200.h"

namespace vct_host {

// ------------------------------------------------------------------------------------------- GLM subset
struct vec3 { float x = 0, y = 0, z = 0; };
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 normalize(vec3 a) { const float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

struct mat4 {                      // m[col*4 + row]
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float& at(int col, int row) { return m[col * 4 + row]; }
    float at(int col, int row) const { return m[col * 4 + row]; }
};
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a.at(k, row) * b.at(c, k);
            r.at(c, row) = s;
        }
    return r;
}
inline mat4 perspective(float fovy, float aspect, float zn, float zf) {       // glm::perspective, RH_NO
    const float t = std::tan(fovy / 2.0f);
    mat4 r; std::memset(r.m, 0, sizeof r.m);
    r.at(0, 0) = 1.0f / (aspect * t); r.at(1, 1) = 1.0f / t;
    r.at(2, 2) = -(zf + zn) / (zf - zn); r.at(2, 3) = -1.0f; r.at(3, 2) = -(2.0f * zf * zn) / (zf - zn);
    return r;
}
inline mat4 ortho(float l, float r_, float b, float t, float zn, float zf) {  // glm::ortho, RH_NO
    mat4 r;
    r.at(0, 0) = 2.0f / (r_ - l); r.at(1, 1) = 2.0f / (t - b); r.at(2, 2) = -2.0f / (zf - zn);
    r.at(3, 0) = -(r_ + l) / (r_ - l); r.at(3, 1) = -(t + b) / (t - b); r.at(3, 2) = -(zf + zn) / (zf - zn);
    return r;
}
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {                           // glm::lookAt, RH
    const vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r;
    r.at(0, 0) = s.x; r.at(1, 0) = s.y; r.at(2, 0) = s.z;
    r.at(0, 1) = u.x; r.at(1, 1) = u.y; r.at(2, 1) = u.z;
    r.at(0, 2) = -f.x; r.at(1, 2) = -f.y; r.at(2, 2) = -f.z;
    r.at(3, 0) = -dot(s, eye); r.at(3, 1) = -dot(u, eye); r.at(3, 2) = dot(f, eye);
    return r;
}
inline mat4 inverse(const mat4& a) {                                           // cofactor expansion, double accumulators
    const float* m = a.m; double inv[16];
    inv[0] = (double)m[5] * m[10] * m[15] - (double)m[5] * m[11] * m[14] - (double)m[9] * m[6] * m[15] + (double)m[9] * m[7] * m[14] + (double)m[13] * m[6] * m[11] - (double)m[13] * m[7] * m[10];
    inv[4] = -(double)m[4] * m[10] * m[15] + (double)m[4] * m[11] * m[14] + (double)m[8] * m[6] * m[15] - (double)m[8] * m[7] * m[14] - (double)m[12] * m[6] * m[11] + (double)m[12] * m[7] * m[10];
    inv[8] = (double)m[4] * m[9] * m[15] - (double)m[4] * m[11] * m[13] - (double)m[8] * m[5] * m[15] + (double)m[8] * m[7] * m[13] + (double)m[12] * m[5] * m[11] - (double)m[12] * m[7] * m[9];
    inv[12] = -(double)m[4] * m[9] * m[14] + (double)m[4] * m[10] * m[13] + (double)m[8] * m[5] * m[14] - (double)m[8] * m[6] * m[13] - (double)m[12] * m[5] * m[10] + (double)m[12] * m[6] * m[9];
    inv[1] = -(double)m[1] * m[10] * m[15] + (double)m[1] * m[11] * m[14] + (double)m[9] * m[2] * m[15] - (double)m[9] * m[3] * m[14] - (double)m[13] * m[2] * m[11] + (double)m[13] * m[3] * m[10];
    inv[5] = (double)m[0] * m[10] * m[15] - (double)m[0] * m[11] * m[14] - (double)m[8] * m[2] * m[15] + (double)m[8] * m[3] * m[14] + (double)m[12] * m[2] * m[11] - (double)m[12] * m[3] * m[10];
    inv[9] = -(double)m[0] * m[9] * m[15] + (double)m[0] * m[11] * m[13] + (double)m[8] * m[1] * m[15] - (double)m[8] * m[3] * m[13] - (double)m[12] * m[1] * m[11] + (double)m[12] * m[3] * m[9];
    inv[13] = (double)m[0] * m[9] * m[14] - (double)m[0] * m[10] * m[13] - (double)m[8] * m[1] * m[14] + (double)m[8] * m[2] * m[13] + (double)m[12] * m[1] * m[10] - (double)m[12] * m[2] * m[9];
    inv[2] = (double)m[1] * m[6] * m[15] - (double)m[1] * m[7] * m[14] - (double)m[5] * m[2] * m[15] + (double)m[5] * m[3] * m[14] + (double)m[13] * m[2] * m[7] - (double)m[13] * m[3] * m[6];
    inv[6] = -(double)m[0] * m[6] * m[15] + (double)m[0] * m[7] * m[14] + (double)m[4] * m[2] * m[15] - (double)m[4] * m[3] * m[14] - (double)m[12] * m[2] * m[7] + (double)m[12] * m[3] * m[6];
    inv[10] = (double)m[0] * m[5] * m[15] - (double)m[0] * m[7] * m[13] - (double)m[4] * m[1] * m[15] + (double)m[4] * m[3] * m[13] + (double)m[12] * m[1] * m[7] - (double)m[12] * m[3] * m[5];
    inv[14] = -(double)m[0] * m[5] * m[14] + (double)m[0] * m[6] * m[13] + (double)m[4] * m[1] * m[14] - (double)m[4] * m[2] * m[13] - (double)m[12] * m[1] * m[6] + (double)m[12] * m[2] * m[5];
    inv[3] = -(double)m[1] * m[6] * m[11] + (double)m[1] * m[7] * m[10] + (double)m[5] * m[2] * m[11] - (double)m[5] * m[3] * m[10] - (double)m[9] * m[2] * m[7] + (double)m[9] * m[3] * m[6];
    inv[7] = (double)m[0] * m[6] * m[11] - (double)m[0] * m[7] * m[10] - (double)m[4] * m[2] * m[11] + (double)m[4] * m[3] * m[10] + (double)m[8] * m[2] * m[7] - (double)m[8] * m[3] * m[6];
    inv[11] = -(double)m[0] * m[5] * m[11] + (double)m[0] * m[7] * m[9] + (double)m[4] * m[1] * m[11] - (double)m[4] * m[3] * m[9] - (double)m[8] * m[1] * m[7] + (double)m[8] * m[3] * m[5];
    inv[15] = (double)m[0] * m[5] * m[10] - (double)m[0] * m[6] * m[9] - (double)m[4] * m[1] * m[10] + (double)m[4] * m[2] * m[9] + (double)m[8] * m[1] * m[6] - (double)m[8] * m[2] * m[5];
    const double det = (double)m[0] * inv[0] + (double)m[1] * inv[4] + (double)m[2] * inv[8] + (double)m[3] * inv[12];
    mat4 r;
    for (int i = 0; i < 16; ++i) r.m[i] = (float)(inv[i] / det);
    return r;
}

// ------------------------------------------------------------- reference src/Application.h:28-103 (hot-path fields)
struct VCTSettings { int steps; float coneAngle, bias, coneInitialHeight, lodOffset; };
struct Settings {
    int drawRadiance = true, axisOverride = -1, drawOcclusion = true;
    // debug views, Application.h:40-60 (all false by default); frameParams() folds them into vct_frame_params::debug_view
    int drawVoxels = false, drawNormals = false, drawDominantAxis = false, debugOcclusion = false, debugIndirect = false, debugReflections = false;
    int debugMaterialDiffuse = false, debugMaterialRoughness = false, debugMaterialMetallic = false;
    int debugVoxels = false;                        // Application.h:50: the voxels as cubes instead of the frame's render passes (Application.cpp:923-928)
    int debugWarpTexture = false, toggle = false;   // Application.h:54, :180: sub-views of drawVoxels (phong.frag:354-357)
    float miplevel = 0.0f;
    int voxelizeTesselation = false;    // Application.h:85 (reference default true; this host defaults to the north star's raster path)
    int voxelizeTesselationWarp = false;   // Application.h:102: the camera frustum as voxel grid (common.glsl:37-42)
    // Application.h:61-62 (reference default MSAA; this host defaults to OFF = the north star's parity mode); NV is not built
    enum ConservativeRasterizeMode { OFF, MSAA, NV };
    int conservativeRasterization = OFF;
    float voxelizeMultiplier = 1.0f;       // Application.h:87: viewport of the voxelise pass = multiplier * voxelDim (Application.cpp:668)
    int voxelTrackCamera = false;          // Application.h:86: the volume follows the camera, snapped to the coarsest mip cell (Application.cpp:187-191)
    int cooktorrance = true, enablePostprocess = true, enableNormalMap = true;
    int enableIndirect = true, enableDiffuse = true, enableSpecular = true, enableReflections = true;
    float ambientScale = 1.0f, reflectScale = 1.0f;
    int radianceLighting = false, radianceDilate = false, temporalFilterRadiance = false;
    float temporalDecay = 0.8f, voxelSetOpacity = 0.5f;
    float warpTextureHighResolution = 2.0f, warpTextureLowResolution = 0.5f;
    int voxelizeLighting = true;
    int voxelizeAtomicMax = false;      // reference default true; north-star parity mode = running-average atomics
    int warpVoxels = false, warpTexture = false, warpTextureLinear = false;
    int warpTextureAxes[3] = {true, true, true};
    int useWarpmapWeightsTexture = true, voxelFillHoles = false;
    VCTSettings diffuseConeSettings{16, radians(60.f), 1.0f, 1.0f, 0.5f};
    VCTSettings specularConeSettings{32, radians(30.f), 1.7f, 0.5f, 0.1f};
    int specularConeAngleFromRoughness = true;
    // this build only
    int deterministicAverage = true;    // canonical-order running average (bit-reproducible) vs free-running CAS
    int mipColorChain = true;           // also filter voxelColor like Application.cpp:903-917
};

struct Camera {                         // reference src/Camera.h: fov = 45.0f handed to glm::perspective as RADIANS [sic]
    vec3 position{0, 0, 0}; float yaw = -90.0f, pitch = 0.0f, fov = 45.0f; vec3 up{0, 1, 0};
    bool hasFront = false; vec3 frontOverride{0, 0, -1};          // headless configs that give a look direction instead of yaw/pitch
    vec3 front() const {
        if (hasFront) return normalize(frontOverride);
        const float cy = std::cos(radians(yaw)), sy = std::sin(radians(yaw)), cp = std::cos(radians(pitch)), sp = std::sin(radians(pitch));
        return {cp * cy, sp, cp * sy};
    }
    mat4 lookAt() const { return vct_host::lookAt(position, position + front(), up); }
};

// ---------------------------------------------------------------------- reference `VCT`, Application.h:107-156
class VCT {
public:
    int voxelDim = 256, voxelLevels = 6;
    vec3 center{0, 0, 0}, min{-20, -20, -20}, max{20, 20, 20};
    vct_ctx* ctx = nullptr;

    // n_devices > 1: one handle that shards every frame over devices 0..n-1 of this process (vct_config.n_devices)
    bool make(int shadow_size, int width, int height, int device = 0, int rank = 0, int world_size = 1, int n_devices = 0) {
        vct_config cfg{};
        cfg.dim = voxelDim; cfg.levels = voxelLevels; cfg.shadow_size = shadow_size; cfg.width = width; cfg.height = height;
        cfg.device = device; cfg.rank = rank; cfg.world_size = world_size; cfg.n_devices = n_devices;
        if (vct_create(&cfg, &ctx)) { std::fprintf(stderr, "[ERROR] %s\n", vct_last_error(nullptr)); ctx = nullptr; return false; }
        return true;
    }
    void remake(int dim, int levels) {
        voxelDim = dim;
        int lg = 0; while ((1 << (lg + 1)) <= dim) lg++;
        voxelLevels = levels < 1 ? 1 : (levels > lg + 1 ? lg + 1 : levels);
        if (voxelLevels != levels) std::fprintf(stderr, "[WARN] Attempted remaking VCT with invalid number of levels, clamped %d to %d\n", levels, voxelLevels);
        if (ctx && vct_remake(ctx, dim, voxelLevels)) std::fprintf(stderr, "[ERROR] %s\n", vct_last_error(ctx));
    }
    ~VCT() { if (ctx) vct_destroy(ctx); }
};

// ------------------------------------------------------------------------------------------------- scene
struct Actor { std::vector<float> vertices; std::vector<uint32_t> indices; std::vector<int32_t> tri_material; mat4 model; };
struct Texture { int width = 0, height = 0, channels = 0, levels = 0; std::vector<uint8_t> pixels; };   // all mips, level 0 first
struct Scene {
    std::vector<Actor> actors; std::vector<Texture> textures; std::vector<vct_material> materials; std::vector<vct_light> lights;

    // flat scene file written by tools/pack_scene.py ("VCTS" little-endian; layout documented there)
    bool load(const char* path) {
        FILE* f = std::fopen(path, "rb");
        if (!f) { std::fprintf(stderr, "[ERROR] cannot open %s\n", path); return false; }
        auto rd = [&](void* p, size_t n) { return std::fread(p, 1, n, f) == n; };
        uint32_t hdr[6];
        bool ok = rd(hdr, sizeof hdr) && hdr[0] == 0x53544356u /* "VCTS" */ && hdr[1] == 1u;
        if (ok) {
            textures.resize(hdr[2]); materials.resize(hdr[3]); actors.resize(hdr[4]); lights.resize(hdr[5]);
            for (auto& t : textures) {
                uint32_t d[5]; ok = ok && rd(d, sizeof d);
                if (!ok) break;
                t.width = (int)d[0]; t.height = (int)d[1]; t.channels = (int)d[2]; t.levels = (int)d[3]; t.pixels.resize(d[4]);
                ok = rd(t.pixels.data(), d[4]);
            }
            for (auto& m : materials) ok = ok && rd(&m, sizeof m);
            for (auto& a : actors) {
                uint32_t d[2]; ok = ok && rd(d, sizeof d);
                if (!ok) break;
                a.vertices.resize((size_t)d[0] * 14); a.indices.resize((size_t)d[1] * 3); a.tri_material.resize(d[1]);
                ok = rd(a.vertices.data(), a.vertices.size() * 4) && rd(a.indices.data(), a.indices.size() * 4) && rd(a.tri_material.data(), a.tri_material.size() * 4) && rd(a.model.m, 64);
            }
            for (auto& l : lights) ok = ok && rd(&l, sizeof l);
        }
        std::fclose(f);
        if (!ok) std::fprintf(stderr, "[ERROR] %s is not a version-1 VCTS scene file\n", path);
        return ok;
    }

    // StaticMeshActor{path} + scene->addActor (Application.cpp:96-99): the OBJ goes through the library's own ingest
    // (vct_ingest_obj: tinyobjloader/stb_image/DDS behaviour restated, include/vct_b200.h) and is appended as one actor
    // with its materials and textures; `resource_dir` is RESOURCE_DIR (common.h:13-15, home of default_texture.png).
    bool addObj(const char* path, const char* resource_dir, float uniform_scale = 1.0f) {
        vct_ingest* g = nullptr;
        const int rc = vct_ingest_obj(path, resource_dir, 0, &g);
        if (g && vct_ingest_log(g)[0]) std::fprintf(stderr, "[%s] %s", rc ? "ERROR" : "WARN", vct_ingest_log(g));
        if (rc) { if (g) vct_ingest_free(g); std::fprintf(stderr, "[ERROR] Failed to load mesh: %s\n", path); return false; }
        vct_ingest_mesh m;
        vct_ingest_get_mesh(g, &m);
        const int tex_base = (int)textures.size(), mat_base = (int)materials.size();
        for (int t = 0; t < m.n_textures; ++t) {
            vct_ingest_texture it;
            vct_ingest_get_texture(g, t, &it);
            Texture tx; tx.width = it.width; tx.height = it.height; tx.channels = it.channels; tx.levels = it.levels;
            tx.pixels.assign((const uint8_t*)it.pixels, (const uint8_t*)it.pixels + it.bytes);
            textures.push_back(std::move(tx));
        }
        for (int k = 0; k < m.n_materials; ++k) {
            vct_material mat;
            vct_ingest_get_material(g, k, &mat, nullptr);
            int* ids[6] = {&mat.diffuse_tex, &mat.specular_tex, &mat.normal_tex, &mat.roughness_tex, &mat.metallic_tex, &mat.alpha_tex};
            for (int* id : ids) if (*id >= 0) *id += tex_base;
            materials.push_back(mat);
        }
        Actor a;
        a.vertices.assign(m.vertices, m.vertices + m.n_vertices * 14);
        a.indices.assign(m.indices, m.indices + m.n_indices);
        a.tri_material.assign(m.material_of_triangle, m.material_of_triangle + m.n_indices / 3);
        for (int32_t& id : a.tri_material) id += mat_base;
        a.model.at(0, 0) = a.model.at(1, 1) = a.model.at(2, 2) = uniform_scale;      // transform.setScale(vec3(s)), Application.cpp:98
        actors.push_back(std::move(a));
        vct_ingest_free(g);
        return true;
    }

    // the two lights Application::init adds (Application.cpp:125-136); Light defaults from Scene.h:13-29
    void addReferenceLights() {
        vct_light main{}; main.type = 1; main.shadow_caster = 1; main.enabled = 1; main.range = 5.0f; main.intensity = 1.0f;
        main.position[0] = 12.0f; main.position[1] = 40.0f; main.position[2] = -7.0f;
        main.direction[0] = -0.38f; main.direction[1] = -0.88f; main.direction[2] = 0.2f;
        main.color[0] = main.color[1] = main.color[2] = 1.0f;
        vct_light test{}; test.type = 0; test.enabled = 1; test.range = 5.0f; test.intensity = 1.0f;
        test.position[1] = 10.0f; test.direction[2] = -1.0f; test.color[0] = 1.0f; test.color[2] = 1.0f;
        lights.push_back(main); lights.push_back(test);
    }
};

struct Timers { double voxelize = 0, shadowmap = 0, radiance = 0, mipmap = 0, render = 0, total = 0; };   // ms, GLBufferedTimer names

// ------------------------------------------------------------- reference `Application`, hot-path members only
class Application {
public:
    int width = 1280, height = 720;                    // common.h:9-10
    float near_ = 0.1f, far_ = 100.0f;                 // Application.h:172
    static constexpr int SHADOWMAP_WIDTH = 4096;       // Application.cpp:30-31
    Camera camera; Settings settings; VCT vct; Scene* scene = nullptr;
    Timers timers; vct_voxelize_info voxelizeInfo{};
    float clearColor[3] = {0.5294f, 0.8078f, 0.9216f}; // Application.cpp:41

    // Application::init: create the GPU resources and upload the scene (Mesh VAO/EBO/texture creation)
    bool init(Scene* s, int shadow_size = SHADOWMAP_WIDTH, int device = 0, int n_devices = 0) {
        scene = s; shadow_size_ = shadow_size;
        if (!vct.make(shadow_size, width, height, device, 0, 1, n_devices)) return false;
        bool ok = true;
        for (size_t i = 0; i < s->textures.size(); ++i) {
            const Texture& t = s->textures[i];
            ok &= ck(vct_upload_texture(vct.ctx, (int)i, t.width, t.height, t.channels, t.levels, t.pixels.data()));
        }
        for (size_t i = 0; i < s->materials.size(); ++i) ok &= ck(vct_set_material(vct.ctx, (int)i, &s->materials[i]));
        for (size_t a = 0; a < s->actors.size(); ++a) {
            const Actor& A = s->actors[a];
            ok &= ck(vct_upload_mesh(vct.ctx, (int)a, A.vertices.data(), A.vertices.size() / 14, 56, A.indices.data(), A.indices.size(), A.tri_material.data()));
        }
        return ok;
    }

    // Application::render(dt), src/Application.cpp:196-1085 — GL blocks replaced one for one
    // the uniforms of one frame: every matrix the reference computes on the CPU + the Settings scalars
    vct_frame_params frameParams() const {
        vct_frame_params p{};
        // :200-210 — camera and light matrices
        const mat4 projection = perspective(camera.fov, (float)width / (float)height, near_, far_);
        const mat4 view = camera.lookAt();
        const mat4 pv = perspective(camera.fov, (float)width / (float)height, 1.f, 20.f) * view;
        const vct_light& mainlight = scene->lights.at(0);
        const vec3 lpos{mainlight.position[0], mainlight.position[1], mainlight.position[2]}, ldir{mainlight.direction[0], mainlight.direction[1], mainlight.direction[2]};
        const mat4 lp = ortho(-25.f, 25.f, -25.f, 25.f, 0.f, 100.f);
        const mat4 lv = lookAt(lpos, lpos + ldir, {0, 1, 0});
        const mat4 ls = lp * lv;
        // :689-692 — the three voxelisation views
        const mat4 vproj = ortho(vct.min.x, vct.max.x, vct.min.y, vct.max.y, 0.0f, vct.max.z - vct.min.z);
        const mat4 mvp_x = vproj * lookAt(vct.center + vec3{vct.max.x, 0, 0}, vct.center, {0, 1, 0});
        const mat4 mvp_y = vproj * lookAt(vct.center + vec3{0, vct.max.y, 0}, vct.center, {0, 0, -1});
        const mat4 mvp_z = vproj * lookAt(vct.center + vec3{0, 0, vct.max.z}, vct.center, {0, 1, 0});
        auto put = [](float* d, const mat4& m) { std::memcpy(d, m.m, 64); };
        put(p.projection, projection); put(p.view, view); put(p.pv, pv); put(p.lp, lp); put(p.lv, lv); put(p.ls, ls);
        put(p.ls_inverse, inverse(ls)); put(p.mvp_x, mvp_x); put(p.mvp_y, mvp_y); put(p.mvp_z, mvp_z);
        const float* vs[4] = {&camera.position.x, &vct.min.x, &vct.max.x, &vct.center.x}; float* vd[4] = {p.eye, p.voxel_min, p.voxel_max, p.voxel_center};
        for (int i = 0; i < 4; ++i) std::memcpy(vd[i], vs[i], 12);
        std::memcpy(p.clear_color, clearColor, 12);
        // every glUniform of :695-714, :805-822, :983-1042 in one POD
        const Settings& s = settings;
        p.voxelize_lighting = s.voxelizeLighting; p.voxelize_atomic_max = s.voxelizeAtomicMax; p.axis_override = s.axisOverride;
        p.deterministic = s.deterministicAverage; p.voxel_set_opacity = s.voxelSetOpacity;
        p.temporal_filter_radiance = s.temporalFilterRadiance; p.temporal_decay = s.temporalDecay;
        p.radiance_lighting = s.radianceLighting; p.radiance_dilate = s.radianceDilate; p.voxel_fill_holes = s.voxelFillHoles;
        p.mip_color_chain = s.mipColorChain;
        p.warp_voxels = s.warpVoxels; p.warp_texture = s.warpTexture; p.warp_texture_linear = s.warpTextureLinear;
        for (int i = 0; i < 3; ++i) p.warp_texture_axes[i] = s.warpTextureAxes[i];
        p.use_warpmap_weights_texture = s.useWarpmapWeightsTexture;
        p.warp_texture_high_resolution = s.warpTextureHighResolution; p.warp_texture_low_resolution = s.warpTextureLowResolution;
        p.draw_radiance = s.drawRadiance; p.draw_occlusion = s.drawOcclusion; p.cooktorrance = s.cooktorrance;
        p.enable_postprocess = s.enablePostprocess; p.enable_normal_map = s.enableNormalMap;
        p.enable_indirect = s.enableIndirect; p.enable_diffuse = s.enableDiffuse; p.enable_specular = s.enableSpecular; p.enable_reflections = s.enableReflections;
        p.ambient_scale = s.ambientScale; p.reflect_scale = s.reflectScale;
        auto cone = [](const VCTSettings& c) { vct_cone_settings o; o.steps = c.steps; o.cone_angle = c.coneAngle; o.bias = c.bias; o.cone_initial_height = c.coneInitialHeight; o.lod_offset = c.lodOffset; return o; };
        p.diffuse_cone = cone(s.diffuseConeSettings); p.specular_cone = cone(s.specularConeSettings);
        p.specular_cone_angle_from_roughness = s.specularConeAngleFromRoughness;
        // phong.frag tests its debug uniforms in this order (:346-447, 489-505)
        p.debug_view = s.drawVoxels ? (s.drawNormals ? VCT_VIEW_VOXEL_NORMALS : s.debugWarpTexture ? (s.toggle ? VCT_VIEW_WARP_TEXTURE_TC : VCT_VIEW_WARP_TEXTURE) : VCT_VIEW_VOXELS)
                     : s.debugMaterialDiffuse ? VCT_VIEW_MATERIAL_DIFFUSE : s.debugMaterialRoughness ? VCT_VIEW_MATERIAL_ROUGHNESS
                     : s.debugMaterialMetallic ? VCT_VIEW_MATERIAL_METALLIC : s.drawNormals ? VCT_VIEW_NORMALS : s.drawDominantAxis ? VCT_VIEW_DOMINANT_AXIS
                     : s.debugIndirect ? VCT_VIEW_INDIRECT : s.debugOcclusion ? VCT_VIEW_OCCLUSION : s.debugReflections ? VCT_VIEW_REFLECTIONS : VCT_VIEW_SHADED;
        p.miplevel = s.miplevel;
        p.voxelize_tesselation = s.voxelizeTesselation;
        p.voxelize_tesselation_warp = s.voxelizeTesselationWarp;
        p.voxelize_multiplier = s.voxelizeMultiplier;
        p.conservative_raster = s.conservativeRasterization == Settings::MSAA ? VCT_RASTER_MSAA : VCT_RASTER_CENTER;   // msaa_samples all zero: standard 4x pattern
        return p;
    }

    // Application::update's voxelTrackCamera block (src/Application.cpp:187-191): the volume centre follows the camera in steps of one
    // coarsest-level cell, so that the voxelisation does not swim (glm::floor, glm::pow(2.f, levels) component-wise)
    void trackCamera() {
        if (!settings.voxelTrackCamera) return;
        const float k = std::pow(2.0f, (float)vct.voxelLevels);
        const vec3 cell{k * (vct.max.x - vct.min.x) / (float)vct.voxelDim, k * (vct.max.y - vct.min.y) / (float)vct.voxelDim, k * (vct.max.z - vct.min.z) / (float)vct.voxelDim};
        vct.center = {std::floor(camera.position.x / cell.x) * cell.x, std::floor(camera.position.y / cell.y) * cell.y, std::floor(camera.position.z / cell.z) * cell.z};
    }
    bool render(float /*dt*/) {
        if (!vct.ctx || !scene) return false;
        trackCamera();
        bool ok = true;
        const Settings& s = settings;
        const vct_frame_params p = frameParams();
        // Scene::draw's per-actor "model" uniform (Scene.cpp:31-36) and the light SSBO (Scene.cpp:58-62)
        for (size_t a = 0; a < scene->actors.size(); ++a) ok &= ck(vct_set_actor_transform(vct.ctx, (int)a, scene->actors[a].model.m));
        ok &= ck(vct_set_lights(vct.ctx, scene->lights.data(), (int)scene->lights.size()));

        ok &= ck(vct_shadowmap(vct.ctx, &p));                          // :212-233  shadowmapTimer
        if (s.warpTexture) {
            ok &= ck(vct_occupancy(vct.ctx, &p));                      // :235-301
            ok &= ck(vct_warpmap(vct.ctx, &p));                        // :303-577  (no CPU read-back any more)
        }
        ok &= ck(vct_voxelize(vct.ctx, &p));                           // :581-755  voxelizeTimer
        ok &= ck(vct_transfer(vct.ctx, &p));                           // :757-785
        ok &= ck(vct_inject(vct.ctx, &p));                             // :787-837  radianceTimer
        if (s.voxelFillHoles) ok &= ck(vct_fill_holes(vct.ctx, &p));   // :839-875
        ok &= ck(vct_mip(vct.ctx, VCT_VOL_RADIANCE));                  // :877-902  mipmapTimer
        if (s.mipColorChain || !s.drawRadiance) ok &= ck(vct_mip(vct.ctx, VCT_VOL_COLOR));   // :903-917
        if (s.debugVoxels) ok &= ck(vct_debug_voxels(vct.ctx, &p));     // :923-928  instead of the prepass and the shading pass
        else {
            ok &= ck(vct_gbuffer(vct.ctx, &p));                        // :936-965  depth prepass
            ok &= ck(vct_cone_trace(vct.ctx, &p));                     // :967-1067 renderTimer
        }
        ok &= ck(vct_get_counters(vct.ctx, &voxelizeInfo));            // Overlay.cpp:104-110
        return ok;
    }

    // the whole graph as ONE library call (same passes, pass-level timers filled like GLBufferedTimer::getTime)
    bool renderFused(float /*dt*/) {
        if (!vct.ctx || !scene) return false;
        trackCamera();
        const vct_frame_params p = frameParams();
        bool ok = true;
        for (size_t a = 0; a < scene->actors.size(); ++a) ok &= ck(vct_set_actor_transform(vct.ctx, (int)a, scene->actors[a].model.m));
        ok &= ck(vct_set_lights(vct.ctx, scene->lights.data(), (int)scene->lights.size()));
        ok &= ck(vct_frame(vct.ctx, &p));
        ok &= ck(vct_get_counters(vct.ctx, &voxelizeInfo));
        vct_timings t{};
        if (ok && ck(vct_get_timings(vct.ctx, &t))) {
            timers.voxelize = t.voxelize_ns * 1e-6; timers.shadowmap = t.shadowmap_ns * 1e-6; timers.radiance = t.radiance_ns * 1e-6;
            timers.mipmap = t.mipmap_ns * 1e-6; timers.render = t.render_ns * 1e-6; timers.total = t.total_ns * 1e-6;
        }
        return ok;
    }

    // glReadPixels of the default framebuffer (row 0 = bottom)
    bool readPixels(std::vector<uint8_t>& rgba) { rgba.resize((size_t)width * height * 4); return ck(vct_read_image(vct.ctx, rgba.data())); }

private:
    int shadow_size_ = SHADOWMAP_WIDTH;
    bool ck(int status) {
        if (status) std::fprintf(stderr, "[ERROR] %s\n", vct_last_error(vct.ctx));      // log and carry on, like src/log.h
        return status == 0;
    }
};

}  // namespace vct_host
