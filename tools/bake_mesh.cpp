// bake_mesh — harness tool (NOT part of the product path, NOT shipped to the GPU box as source of truth).
//
// Turns an OBJ/MTL into the flat arrays the hot path consumes, following the *behaviour* of the
// reference's loader so the inputs on both sides of the parity tests are the reference's inputs:
//   - OBJ parse + fan triangulation: the reference's vendored tinyobjloader 1.0.7, compiled IN PLACE from
//     /root/reference/ext/include (never copied into this repo)
//   - vertex de-duplication on (v, vn, vt) index triples, v-flip of texcoords, un-weighted accumulation
//     of per-face tangent/bitangent then normalisation (NaN for meshes without UVs)
//     → reference src/Graphics/Mesh.cpp:120-206
//   - one index list ("drawable") per material, drawn in material order, plus an extra trailing
//     "default" material for faces without one → reference src/Graphics/Mesh.cpp:91-118, 340-371
//   - only diffuse/specular/normal/roughness/metallic/alpha texture names are honoured
//     → reference src/Graphics/Mesh.cpp:81-86
//
// Output (directory given as argv[2]): vertices.f32 (n×14), indices.u32 (3T, draw order),
// tri_material.i32 (T), materials.txt (one line per material: name|diffuse|specular|normal|roughness|metallic|alpha)
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <vector>

struct V14 { float p[3], n[3], uv[2], t[3], b[3]; };

static void write_file(const std::string& path, const void* data, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(1); }
    fwrite(data, 1, bytes, f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: bake_mesh in.obj outdir\n"); return 2; }
    std::string obj = argv[1], out = argv[2];
    std::string base = obj.substr(0, obj.find_last_of('/') + 1);
    tinyobj::attrib_t attrib; std::vector<tinyobj::shape_t> shapes; std::vector<tinyobj::material_t> mtl;
    std::string err;
    if (!tinyobj::LoadObj(&attrib, &shapes, &mtl, &err, obj.c_str(), base.c_str())) {
        fprintf(stderr, "load failed: %s\n", err.c_str()); return 1;
    }
    const size_t n_mat = mtl.size() + 1;               // + trailing default material
    std::vector<std::vector<uint32_t>> per_material(n_mat);
    std::vector<V14> verts;
    std::map<std::tuple<int,int,int>, uint32_t> seen;

    for (const auto& sh : shapes) {
        size_t cursor = 0;
        for (size_t f = 0; f < sh.mesh.num_face_vertices.size(); ++f) {
            const int fv = sh.mesh.num_face_vertices[f];
            if (fv != 3) { fprintf(stderr, "non-triangle face\n"); return 1; }
            int m = sh.mesh.material_ids[f];
            auto& list = per_material[m < 0 ? n_mat - 1 : (size_t)m];
            uint32_t id[3];
            for (int k = 0; k < 3; ++k) {
                const tinyobj::index_t ix = sh.mesh.indices[cursor + k];
                auto key = std::make_tuple(ix.vertex_index, ix.normal_index, ix.texcoord_index);
                auto it = seen.find(key);
                if (it == seen.end()) {
                    V14 v{};                                  // zero-initialised like Vertex{}
                    for (int c = 0; c < 3; ++c) v.p[c] = attrib.vertices[3 * ix.vertex_index + c];
                    if (ix.normal_index >= 0)
                        for (int c = 0; c < 3; ++c) v.n[c] = attrib.normals[3 * ix.normal_index + c];
                    if (ix.texcoord_index >= 0) {
                        v.uv[0] = attrib.texcoords[2 * ix.texcoord_index];
                        v.uv[1] = 1.f - attrib.texcoords[2 * ix.texcoord_index + 1];
                    }
                    id[k] = (uint32_t)verts.size();
                    seen.emplace(key, id[k]);
                    verts.push_back(v);
                } else id[k] = it->second;
                list.push_back(id[k]);
            }
            // per-face tangent frame, summed un-weighted into the three corners
            const V14 &a = verts[id[0]], &b = verts[id[1]], &c = verts[id[2]];
            float e1[3], e2[3];
            for (int k = 0; k < 3; ++k) { e1[k] = b.p[k] - a.p[k]; e2[k] = c.p[k] - a.p[k]; }
            const float du1 = b.uv[0] - a.uv[0], dv1 = b.uv[1] - a.uv[1];
            const float du2 = c.uv[0] - a.uv[0], dv2 = c.uv[1] - a.uv[1];
            const float inv = 1.0f / (du1 * dv2 - du2 * dv1);
            float tg[3], bt[3];
            for (int k = 0; k < 3; ++k) {
                tg[k] = inv * (dv2 * e1[k] - dv1 * e2[k]);
                bt[k] = inv * (du2 * e1[k] - du1 * e2[k]);
            }
            for (int k = 0; k < 3; ++k)
                for (int c2 = 0; c2 < 3; ++c2) { verts[id[k]].t[c2] += tg[c2]; verts[id[k]].b[c2] += bt[c2]; }
            cursor += 3;
        }
    }
    for (auto& v : verts) {
        // glm::normalize = v * inversesqrt(dot(v,v))
        float lt = 1.0f / std::sqrt(v.t[0]*v.t[0] + v.t[1]*v.t[1] + v.t[2]*v.t[2]);
        float lb = 1.0f / std::sqrt(v.b[0]*v.b[0] + v.b[1]*v.b[1] + v.b[2]*v.b[2]);
        for (int k = 0; k < 3; ++k) { v.t[k] *= lt; v.b[k] *= lb; }
    }
    std::vector<uint32_t> indices; std::vector<int32_t> tri_mat;
    for (size_t m = 0; m < n_mat; ++m) {
        indices.insert(indices.end(), per_material[m].begin(), per_material[m].end());
        tri_mat.insert(tri_mat.end(), per_material[m].size() / 3, (int32_t)m);
    }
    write_file(out + "/vertices.f32", verts.data(), verts.size() * sizeof(V14));
    write_file(out + "/indices.u32", indices.data(), indices.size() * 4);
    write_file(out + "/tri_material.i32", tri_mat.data(), tri_mat.size() * 4);
    FILE* f = fopen((out + "/materials.txt").c_str(), "w");
    for (size_t m = 0; m + 1 < n_mat; ++m) {
        const auto& k = mtl[m];
        fprintf(f, "%s|%s|%s|%s|%s|%s|%s\n", k.name.c_str(), k.diffuse_texname.c_str(), k.specular_texname.c_str(),
                k.normal_texname.c_str(), k.roughness_texname.c_str(), k.metallic_texname.c_str(), k.alpha_texname.c_str());
    }
    fprintf(f, "default|@default_texture.png|||||\n");
    fclose(f);
    fprintf(stderr, "%s: %zu vertices, %zu triangles, %zu materials\n", obj.c_str(), verts.size(), indices.size() / 3, n_mat);
    return 0;
}
