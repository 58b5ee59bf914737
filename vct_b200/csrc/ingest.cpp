// ingest.cpp — C ABI over vct_b200/host/vct_ingest.hpp (scene ingest).  Host code only: no CUDA call is
// made here except through the public vct_upload_* entry points in vct_ingest_upload.
#include <new>

#include "../../include/vct_b200.h"
#include "../host/vct_ingest.hpp"

struct vct_ingest {
    vct::IngestScene scene;
    std::vector<int32_t> scratch;
};

extern "C" {

int vct_ingest_obj(const char* obj_path, const char* resource_dir, int flags, vct_ingest** out) {
    if (!out) return 1;
    *out = nullptr;
    if (!obj_path) return 1;
    vct_ingest* g = new (std::nothrow) vct_ingest();
    if (!g) return 1;
    *out = g;
    try {
        return vct::load_obj_scene(obj_path, resource_dir ? resource_dir : "", g->scene, !(flags & VCT_INGEST_NO_TEXTURES)) ? 0 : 1;
    } catch (const std::exception& e) { g->scene = vct::IngestScene(); g->scene.log = std::string("vct_ingest_obj: ") + e.what(); return 1; }
}

int vct_ingest_image(const char* path, int generate_mips, vct_ingest** out) {
    if (!out) return 1;
    *out = nullptr;
    if (!path) return 1;
    vct_ingest* g = new (std::nothrow) vct_ingest();
    if (!g) return 1;
    *out = g;
    try {
        vct::Image im = vct::load_texture_file(path, generate_mips != 0);
        if (!im.error.empty()) { g->scene.log = im.error; return 1; }
        g->scene.textures.push_back(std::move(im));
        g->scene.texture_names.push_back(path);
        return 0;
    } catch (const std::exception& e) { g->scene.log = std::string("vct_ingest_image: ") + e.what(); return 1; }
}

void vct_ingest_free(vct_ingest* g) { delete g; }
const char* vct_ingest_log(const vct_ingest* g) { return g ? g->scene.log.c_str() : "null ingest handle"; }

int vct_ingest_get_mesh(const vct_ingest* g, vct_ingest_mesh* out) {
    if (!g || !out) return 1;
    const vct::IngestMesh& m = g->scene.mesh;
    out->vertices = m.vertices.data(); out->n_vertices = m.vertices.size() / 14;
    out->indices = m.indices.data(); out->n_indices = m.indices.size();
    out->material_of_triangle = m.tri_material.data();
    out->n_materials = (int)m.materials.size(); out->n_textures = (int)g->scene.textures.size();
    for (int a = 0; a < 3; ++a) { out->bounds_min[a] = m.bounds_min[a]; out->bounds_max[a] = m.bounds_max[a]; }
    out->radius = m.radius;
    return 0;
}

int vct_ingest_get_material(const vct_ingest* g, int material, vct_material* out, const char** name) {
    if (!g || !out || material < 0 || material >= (int)g->scene.slots.size()) return 1;
    const vct::IngestSlots& s = g->scene.slots[material];
    out->diffuse_tex = s.tex[0]; out->specular_tex = s.tex[1]; out->normal_tex = s.tex[2];
    out->roughness_tex = s.tex[3]; out->metallic_tex = s.tex[4]; out->alpha_tex = s.tex[5];
    out->shininess = s.shininess; out->diffuse[0] = out->diffuse[1] = out->diffuse[2] = 0.0f;
    if (name) *name = g->scene.mesh.materials[material].name.c_str();
    return 0;
}

int vct_ingest_get_texture(const vct_ingest* g, int texture, vct_ingest_texture* out) {
    if (!g || !out || texture < 0 || texture >= (int)g->scene.textures.size()) return 1;
    const vct::Image& im = g->scene.textures[texture];
    out->width = im.width; out->height = im.height; out->channels = im.channels; out->levels = im.levels;
    out->pixels = im.pixels.data(); out->bytes = im.pixels.size();
    out->name = g->scene.texture_names[texture].c_str();
    return 0;
}

int vct_ingest_upload(vct_ctx* c, const vct_ingest* g, int actor, int material_base, int texture_base, const float model[16]) {
    if (!c || !g) return 1;
    try {
    const vct::IngestScene& s = g->scene;
    for (size_t t = 0; t < s.textures.size(); ++t) {
        const vct::Image& im = s.textures[t];
        if (im.levels < 1) continue;                                   // VCT_INGEST_NO_TEXTURES: nothing decoded
        if (vct_upload_texture(c, texture_base + (int)t, im.width, im.height, im.channels, im.levels, im.pixels.data())) return 1;
    }
    for (size_t m = 0; m < s.slots.size(); ++m) {
        vct_material mat;
        vct_ingest_get_material(g, (int)m, &mat, nullptr);
        int* ids[6] = {&mat.diffuse_tex, &mat.specular_tex, &mat.normal_tex, &mat.roughness_tex, &mat.metallic_tex, &mat.alpha_tex};
        for (int* id : ids) if (*id >= 0) *id = s.textures[(size_t)*id].levels < 1 ? -1 : *id + texture_base;
        if (vct_set_material(c, material_base + (int)m, &mat)) return 1;
    }
    std::vector<int32_t> rebased(s.mesh.tri_material);
    for (int32_t& m : rebased) m += material_base;
    if (vct_upload_mesh(c, actor, s.mesh.vertices.data(), s.mesh.vertices.size() / 14, 56, s.mesh.indices.data(), s.mesh.indices.size(), rebased.data())) return 1;
    static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    return vct_set_actor_transform(c, actor, model ? model : identity);
    } catch (const std::exception&) { return 1; }                      // (bad_alloc of the rebased material list: nothing may cross the C ABI)
}

}  // extern "C"
