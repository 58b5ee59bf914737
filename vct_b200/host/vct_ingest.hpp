// vct_ingest.hpp — scene ingest for the host side of libvct_b200: OBJ/MTL -> the Vertex / index /
// per-triangle material arrays that vct_upload_mesh and vct_set_material take.  Header-only C++17, no dependencies.
//
// It restates, in this repository's own code, what the reference does when it loads a mesh:
//   * the OBJ/MTL reading of its vendored tinyobjloader 1.0.7 (reference ext/include/tiny_obj_loader.h, called from
//     src/Graphics/Mesh.cpp:45) as far as the hot path's inputs depend on it: line splitting on \n, \r\n and \r; `v`, `vn`,
//     `vt`, `f`, `usemtl`, `mtllib`; 1-based and negative (relative) indices resolved against the counts read so far;
//     i, i/j, i//k, i/j/k corners; fan triangulation (c0, c(k-1), c(k)); the material of a face is the one named by the
//     last `usemtl` (unknown name: none); `newmtl` / `map_Kd` / `map_Ks` / `norm` / `map_Pr` / `map_Pm` / `map_d` with
//     their `-option value` prefixes; and its NUMBER GRAMMAR — digits accumulated in a double, decimals added as
//     digit * 10^-k (table for k < 8, pow beyond), exponent applied as ldexp(m * 5^e, e) — so that every float is the
//     float tinyobjloader produces, not merely close to it;
//   * Mesh::loadMesh (src/Graphics/Mesh.cpp:120-206): corner de-duplication on the (v, vn, vt) triple in first-use order,
//     v-flip of the texture coordinate, per-face tangent/bitangent summed un-weighted into the corners and normalised
//     as v * (1/sqrt(dot)) (NaN for meshes without UVs: the reference's behaviour), one index list per material in
//     material order with a trailing "default" material for faces without one (:91-118), bounds and radius (:197-204).
// tests/test_ingest.py checks the result byte for byte against tools/bake_mesh.cpp, which runs the reference's own
// tinyobjloader compiled in place, on every OBJ the reference ships.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "vct_ingest_image.hpp"

namespace vct {

struct IngestMaterial { std::string name, diffuse, specular, normal, roughness, metallic, alpha; float shininess = 1.0f; };
struct IngestMesh {
    std::vector<float> vertices;            // 14 floats per vertex: position, normal, uv, tangent, bitangent (Mesh.h:72-76)
    std::vector<uint32_t> indices;          // 3 per triangle, in draw order (material by material)
    std::vector<int32_t> tri_material;      // per triangle
    std::vector<IngestMaterial> materials;  // file materials + the trailing "default"
    float bounds_min[3] = {0, 0, 0}, bounds_max[3] = {0, 0, 0}, radius = 0.0f;
    std::string warnings;
};

namespace ingest_detail {

inline bool is_space(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }

// lines end at \n, \r\n or \r (and at end of file)
inline std::vector<std::string> split_lines(const std::string& text) {
    std::vector<std::string> lines;
    size_t i = 0, n = text.size();
    while (i < n) {
        size_t j = i;
        while (j < n && text[j] != '\n' && text[j] != '\r') ++j;
        lines.emplace_back(text, i, j - i);
        if (j < n && text[j] == '\r' && j + 1 < n && text[j + 1] == '\n') ++j;
        i = j + 1;
    }
    return lines;
}
inline bool read_file(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16]; size_t got;
    out.clear();
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    std::fclose(f);
    return true;
}

// The number grammar described in the header comment: [sign] digits [. digits] [(e|E) [sign] digits].
// Returns false (value untouched) when the text does not start like a number.
inline bool parse_number(const char* s, const char* end, double& value) {
    if (s >= end) return false;
    const char* p = s;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    else if (!is_digit(*p)) return false;
    double m = 0.0;
    int count = 0;
    while (p != end && is_digit(*p)) { m *= 10; m += (int)(*p - '0'); ++p; ++count; }
    if (count == 0) return false;
    int e10 = 0;
    if (p != end) {
        bool want_exponent = false;
        if (*p == '.') {
            ++p;
            static const double tenth[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            int k = 1;
            while (p != end && is_digit(*p)) { m += (int)(*p - '0') * (k < 8 ? tenth[k] : std::pow(10.0, -k)); ++k; ++p; }
            want_exponent = p != end;
        } else if (*p == 'e' || *p == 'E') want_exponent = true;
        if (want_exponent && (*p == 'e' || *p == 'E')) {
            ++p;
            bool eneg = false;
            if (p != end && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
            else if (!is_digit(*p)) return false;                  // "1e" is not a number
            int digits = 0;
            while (p != end && is_digit(*p)) { e10 = e10 * 10 + (int)(*p - '0'); ++p; ++digits; }
            if (eneg) e10 = -e10;
            if (digits == 0) return false;
        }
    }
    value = (neg ? -1 : 1) * (e10 ? std::ldexp(m * std::pow(5.0, e10), e10) : m);
    return true;
}
// next blank-delimited field as a float; a field that is not a number yields `fallback`
inline float next_real(const char*& t, double fallback = 0.0) {
    t += std::strspn(t, " \t");
    const char* end = t + std::strcspn(t, " \t\r");
    double v = fallback;
    parse_number(t, end, v);
    t = end;
    return (float)v;
}
inline int resolve_index(int idx, int count) { return idx > 0 ? idx - 1 : idx == 0 ? 0 : count + idx; }

struct Corner { int v = -1, vn = -1, vt = -1; };
inline Corner next_corner(const char*& t, int nv, int nvn, int nvt) {
    Corner c;
    c.v = resolve_index(std::atoi(t), nv);
    t += std::strcspn(t, "/ \t\r");
    if (*t != '/') return c;
    ++t;
    if (*t == '/') {                                                   // i//k
        ++t;
        c.vn = resolve_index(std::atoi(t), nvn);
        t += std::strcspn(t, "/ \t\r");
        return c;
    }
    c.vt = resolve_index(std::atoi(t), nvt);                           // i/j or i/j/k
    t += std::strcspn(t, "/ \t\r");
    if (*t != '/') return c;
    ++t;
    c.vn = resolve_index(std::atoi(t), nvn);
    t += std::strcspn(t, "/ \t\r");
    return c;
}
inline std::string first_word(const char* t) {                         // sscanf("%s")
    while (*t == ' ' || *t == '\t' || *t == '\n' || *t == '\r' || *t == '\f' || *t == '\v') ++t;
    const char* e = t;
    while (*e && !(*e == ' ' || *e == '\t' || *e == '\n' || *e == '\r' || *e == '\f' || *e == '\v')) ++e;
    return std::string(t, e);
}
inline bool keyword(const char* t, const char* word) { const size_t n = std::strlen(word); return std::strncmp(t, word, n) == 0 && is_space(t[n]); }

// `map_* [-option value ...] file`: the options are skipped with their arguments, the first bare field is the file name
inline bool texture_name(const char* t, std::string& name) {
    bool found = false; std::string result;
    auto skip_fields = [&](int n) { for (int i = 0; i < n; ++i) next_real(t); };
    while (!(*t == '\r' || *t == '\n' || *t == '\0')) {
        t += std::strspn(t, " \t");
        if (keyword(t, "-blendu") || keyword(t, "-blendv")) { t += 8; t += std::strspn(t, " \t"); t += std::strcspn(t, " \t\r"); }
        else if (keyword(t, "-clamp")) { t += 7; t += std::strspn(t, " \t"); t += std::strcspn(t, " \t\r"); }
        else if (keyword(t, "-boost")) { t += 7; skip_fields(1); }
        else if (keyword(t, "-bm")) { t += 4; skip_fields(1); }
        else if (keyword(t, "-o")) { t += 3; skip_fields(3); }
        else if (keyword(t, "-s")) { t += 3; skip_fields(3); }
        else if (keyword(t, "-t")) { t += 3; skip_fields(3); }
        else if (keyword(t, "-type")) { t += 5; t += std::strspn(t, " \t"); t += std::strcspn(t, " \t\r"); }
        else if (keyword(t, "-imfchan")) { t += 9; t += std::strspn(t, " \t"); t += std::strcspn(t, " \t\r"); }
        else if (keyword(t, "-mm")) { t += 4; skip_fields(2); }
        else {
            const size_t len = std::strcspn(t, " \t\r");
            result.assign(t, len);
            t += len;
            t += std::strspn(t, " \t");
            found = true;
        }
    }
    if (found) name = result;
    return found;
}

inline void load_mtl(const std::string& text, std::vector<IngestMaterial>& materials, std::map<std::string, int>& by_name) {
    IngestMaterial cur;
    for (std::string line : split_lines(text)) {
        const size_t last = line.find_last_not_of(" \t");
        line = last == std::string::npos ? std::string() : line.substr(0, last + 1);
        if (line.empty()) continue;
        const char* t = line.c_str();
        t += std::strspn(t, " \t");
        if (*t == '\0' || *t == '#') continue;
        if (keyword(t, "newmtl")) {
            if (!cur.name.empty()) { by_name.insert({cur.name, (int)materials.size()}); materials.push_back(cur); }
            cur = IngestMaterial();
            cur.name = first_word(t + 7);
            continue;
        }
        if (keyword(t, "Ns")) { t += 2; cur.shininess = next_real(t); continue; }
        if (keyword(t, "map_Kd")) { texture_name(t + 7, cur.diffuse); continue; }
        if (keyword(t, "map_Ks")) { texture_name(t + 7, cur.specular); continue; }
        if (keyword(t, "map_d")) { cur.alpha = t + 6; texture_name(t + 6, cur.alpha); continue; }
        if (keyword(t, "map_Pr")) { texture_name(t + 7, cur.roughness); continue; }
        if (keyword(t, "map_Pm")) { texture_name(t + 7, cur.metallic); continue; }
        if (keyword(t, "norm")) { texture_name(t + 5, cur.normal); continue; }
    }
    by_name.insert({cur.name, (int)materials.size()});                  // the last material is always kept, named or not
    materials.push_back(cur);
}

}  // namespace ingest_detail

// Returns false (with `out.warnings` set) only when the OBJ itself cannot be read; a missing MTL is a warning.
inline bool load_obj(const std::string& path, IngestMesh& out) {
    using namespace ingest_detail;
    out = IngestMesh();
    std::string text;
    if (!read_file(path, text)) { out.warnings = "cannot open " + path; return false; }
    const std::string base = path.substr(0, path.find_last_of('/') + 1);

    std::vector<float> v, vn, vt;
    std::vector<IngestMaterial> file_materials;
    std::map<std::string, int> by_name;
    struct Tri { Corner c[3]; int material; };
    std::vector<Tri> tris;
    int material = -1;
    for (const std::string& line : split_lines(text)) {
        if (line.empty()) continue;
        const char* t = line.c_str();
        t += std::strspn(t, " \t");
        if (*t == '\0' || *t == '#') continue;
        if (t[0] == 'v' && is_space(t[1])) { t += 2; const float x = next_real(t), y = next_real(t), z = next_real(t); v.insert(v.end(), {x, y, z}); continue; }
        if (t[0] == 'v' && t[1] == 'n' && is_space(t[2])) { t += 3; const float x = next_real(t), y = next_real(t), z = next_real(t); vn.insert(vn.end(), {x, y, z}); continue; }
        if (t[0] == 'v' && t[1] == 't' && is_space(t[2])) { t += 3; const float x = next_real(t), y = next_real(t); vt.insert(vt.end(), {x, y}); continue; }
        if (t[0] == 'f' && is_space(t[1])) {
            t += 2; t += std::strspn(t, " \t");
            std::vector<Corner> face;
            while (!(*t == '\r' || *t == '\n' || *t == '\0')) {
                face.push_back(next_corner(t, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2)));
                t += std::strspn(t, " \t\r");
            }
            for (size_t k = 2; k < face.size(); ++k) tris.push_back(Tri{{face[0], face[k - 1], face[k]}, material});   // fan
            continue;
        }
        if (keyword(t, "usemtl")) {
            const std::string name = first_word(t + 7);
            const auto it = by_name.find(name);
            material = it == by_name.end() ? -1 : it->second;
            continue;
        }
        if (keyword(t, "mtllib")) {
            // file names separated by single blanks; the first one that can be read is used
            std::string rest(t + 7);
            bool found = false;
            size_t i = 0;
            while (i <= rest.size() && !found) {
                size_t j = rest.find(' ', i);
                if (j == std::string::npos) j = rest.size();
                const std::string name = rest.substr(i, j - i);
                std::string mtl;
                if (read_file(base + name, mtl)) { load_mtl(mtl, file_materials, by_name); found = true; }
                else out.warnings += "material file " + base + name + " not found\n";
                i = j + 1;
            }
            continue;
        }
    }

    // ---- Mesh::loadMesh
    out.materials = file_materials;
    IngestMaterial def; def.name = "default"; def.diffuse = "@default_texture.png"; def.shininess = 1.0f;   // Mesh.cpp:91-107 (resources/default_texture.png)
    out.materials.push_back(def);
    const size_t n_mat = out.materials.size();
    std::vector<std::vector<uint32_t>> per_material(n_mat);
    std::map<std::tuple<int, int, int>, uint32_t> seen;
    std::vector<float>& V = out.vertices;
    const int nv = (int)(v.size() / 3), nvn = (int)(vn.size() / 3), nvt = (int)(vt.size() / 2);
    size_t dropped = 0;
    for (Tri tr : tris) {
        if (tr.material >= (int)n_mat - 1) continue;                   // cannot happen: ids come from this file's materials
        // A corner that names a vertex the file does not have is undefined behaviour in the reference (tinyobjloader and
        // Mesh::loadMesh index their arrays unchecked).  Here: the triangle is dropped, a missing normal / uv reads as absent.
        bool bad = false;
        for (Corner& c : tr.c) {
            if (c.v < 0 || c.v >= nv) bad = true;
            if (c.vn >= nvn || c.vn < -1) c.vn = -1;
            if (c.vt >= nvt || c.vt < -1) c.vt = -1;
        }
        if (bad) { ++dropped; continue; }
        std::vector<uint32_t>& list = per_material[tr.material < 0 ? n_mat - 1 : (size_t)tr.material];
        uint32_t id[3];
        for (int k = 0; k < 3; ++k) {
            const Corner& c = tr.c[k];
            const auto key = std::make_tuple(c.v, c.vn, c.vt);
            const auto it = seen.find(key);
            if (it == seen.end()) {
                id[k] = (uint32_t)(V.size() / 14);
                seen.emplace(key, id[k]);
                V.resize(V.size() + 14, 0.0f);                         // Vertex{}: all zero
                float* o = &V[14 * (size_t)id[k]];
                for (int a = 0; a < 3; ++a) o[a] = v[3 * (size_t)c.v + a];
                if (c.vn >= 0) for (int a = 0; a < 3; ++a) o[3 + a] = vn[3 * (size_t)c.vn + a];
                if (c.vt >= 0) { o[6] = vt[2 * (size_t)c.vt]; o[7] = 1.f - vt[2 * (size_t)c.vt + 1]; }
            } else id[k] = it->second;
            list.push_back(id[k]);
        }
        // per-face tangent frame (Mesh.cpp:174-189), summed un-weighted into the three corners
        float p[3][3], uv[3][2];
        for (int k = 0; k < 3; ++k) { for (int a = 0; a < 3; ++a) p[k][a] = V[14 * (size_t)id[k] + a]; uv[k][0] = V[14 * (size_t)id[k] + 6]; uv[k][1] = V[14 * (size_t)id[k] + 7]; }
        float e1[3], e2[3];
        for (int a = 0; a < 3; ++a) { e1[a] = p[1][a] - p[0][a]; e2[a] = p[2][a] - p[0][a]; }
        const float du1 = uv[1][0] - uv[0][0], dv1 = uv[1][1] - uv[0][1], du2 = uv[2][0] - uv[0][0], dv2 = uv[2][1] - uv[0][1];
        const float inv = 1.0f / (du1 * dv2 - du2 * dv1);
        float tg[3], bt[3];
        for (int a = 0; a < 3; ++a) { tg[a] = inv * (dv2 * e1[a] - dv1 * e2[a]); bt[a] = inv * (du2 * e1[a] - du1 * e2[a]); }
        for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) { V[14 * (size_t)id[k] + 8 + a] += tg[a]; V[14 * (size_t)id[k] + 11 + a] += bt[a]; }
    }
    if (dropped) out.warnings += std::to_string(dropped) + " triangle(s) reference vertices the file does not define: dropped\n";
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float hi[3] = {1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f};                               // numeric_limits<float>::min() [sic], Mesh.cpp:128
    for (size_t i = 0; i < V.size() / 14; ++i) {
        float* o = &V[14 * i];
        const float lt = 1.0f / std::sqrt(o[8] * o[8] + o[9] * o[9] + o[10] * o[10]);                   // glm::normalize = v * inversesqrt(dot(v, v))
        const float lb = 1.0f / std::sqrt(o[11] * o[11] + o[12] * o[12] + o[13] * o[13]);
        for (int a = 0; a < 3; ++a) { o[8 + a] *= lt; o[11 + a] *= lb; }
        for (int a = 0; a < 3; ++a) { lo[a] = o[a] < lo[a] ? o[a] : lo[a]; hi[a] = hi[a] < o[a] ? o[a] : hi[a]; }
    }
    for (int a = 0; a < 3; ++a) { out.bounds_min[a] = lo[a]; out.bounds_max[a] = hi[a]; }
    const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    out.radius = (ex > ey ? (ex > ez ? ex : ez) : (ey > ez ? ey : ez)) / 2.0f;
    for (size_t m = 0; m < n_mat; ++m) {
        out.indices.insert(out.indices.end(), per_material[m].begin(), per_material[m].end());
        out.tri_material.insert(out.tri_material.end(), per_material[m].size() / 3, (int32_t)m);
    }
    return true;
}

// ------------------------------------------------------------------------------------ mesh + materials + textures
// What `Mesh::Mesh(meshname)` leaves behind for Mesh::draw (Mesh.cpp:42-118, 326-371): the arrays above plus one decoded
// texture per distinct texture name (keyed by the name as written in the MTL, Mesh.cpp:58-60; '\\' -> '/' and the OBJ's
// directory prepended before opening, :66-71) and, per material, the six texture slots Mesh::draw binds (units 0 diffuse,
// 1 specular, 5 normal, 7 roughness, 8 metallic, 9 alpha).  The trailing default material takes
// <resource_dir>/default_texture.png (:91-107).  Every material has shininess 32 and diffuse (0,0,0): the reference's
// Material(material_t) constructor assigns its members to themselves (Mesh.h:33-43).
// Where the reference is undefined this loader is defined: a texture that cannot be opened or decoded, or that has two
// channels (GLHelper.cpp:171-205 logs the failure and carries on with an unallocated texture), becomes "no map" (-1)
// and a line in `log`.
struct IngestSlots { int tex[6] = {-1, -1, -1, -1, -1, -1}; float shininess = 32.0f; };   // diffuse, specular, normal, roughness, metallic, alpha
struct IngestScene {
    IngestMesh mesh;
    std::vector<Image> textures;
    std::vector<std::string> texture_names;
    std::vector<IngestSlots> slots;          // one per mesh.materials entry
    std::string log;
};

inline bool load_obj_scene(const std::string& obj_path, const std::string& resource_dir, IngestScene& out, bool decode_textures = true) {
    out = IngestScene();
    if (!load_obj(obj_path, out.mesh)) { out.log = out.mesh.warnings; return false; }
    out.log = out.mesh.warnings;
    const std::string base = obj_path.substr(0, obj_path.find_last_of('/') + 1);
    std::string res = resource_dir;
    if (!res.empty() && res.back() != '/') res += '/';
    std::map<std::string, int> by_name;
    auto bind = [&](const std::string& name) -> int {
        if (name.empty()) return -1;
        const auto it = by_name.find(name);
        if (it != by_name.end()) return it->second;
        std::string file;
        if (name[0] == '@') file = res + name.substr(1);
        else { file = name; for (char& ch : file) if (ch == '\\') ch = '/'; file = base + file; }
        int id = -1;
        if (decode_textures) {
            Image im = load_texture_file(file);
            if (!im.error.empty()) out.log += "TEXTURE::LOAD_FAILED::" + file + " (" + im.error + ") -> no map\n";
            else if (!(im.channels == 1 || im.channels == 3 || im.channels == 4)) out.log += "texture " + file + " has " + std::to_string(im.channels) + " channels: the reference allocates no storage for it -> no map\n";
            else { id = (int)out.textures.size(); out.textures.push_back(std::move(im)); out.texture_names.push_back(name); }
        } else { id = (int)out.textures.size(); out.textures.emplace_back(); out.texture_names.push_back(name); }
        by_name.emplace(name, id);
        return id;
    };
    for (const IngestMaterial& m : out.mesh.materials) {
        IngestSlots s;
        s.tex[0] = bind(m.diffuse); s.tex[1] = bind(m.specular); s.tex[2] = bind(m.normal);
        s.tex[3] = bind(m.roughness); s.tex[4] = bind(m.metallic); s.tex[5] = bind(m.alpha);
        out.slots.push_back(s);
    }
    return true;
}

}  // namespace vct
