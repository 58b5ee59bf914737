// vct_ingest_dump in.obj outdir — runs the library's own OBJ/MTL ingest (vct_ingest.hpp) and writes the flat arrays in the
// layout of tools/bake_mesh.cpp (vertices.f32, indices.u32, tri_material.i32, materials.txt), so that the two can be
// compared byte for byte (tests/test_ingest.py) and a scene can be prepared without the reference's loader.
// vct_ingest_dump --images file... — decodes each PNG / DDS file and prints "<width> <height> <channels> <fnv1a-64 of the
// level-0 bytes>" per file (the format of oracle/ref_rig/stb_dump.cpp); "0 0 0 0" for a file that does not decode.
#include <cstdio>
#include <cstdlib>

#include "vct_ingest.hpp"

static void write_file(const std::string& path, const void* data, size_t bytes) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { std::perror(path.c_str()); std::exit(1); }
    std::fwrite(data, 1, bytes, f);
    std::fclose(f);
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "--images")) {
        for (int i = 2; i < argc; ++i) {
            const vct::Image im = vct::load_texture_file(argv[i], false);
            if (!im.error.empty()) { std::fprintf(stderr, "%s: %s\n", argv[i], im.error.c_str()); std::printf("0 0 0 0\n"); continue; }
            uint64_t hash = 1469598103934665603ull;
            const size_t n = (size_t)im.width * im.height * im.channels;
            for (size_t k = 0; k < n; ++k) { hash ^= im.pixels[k]; hash *= 1099511628211ull; }
            std::printf("%d %d %d %016llx\n", im.width, im.height, im.channels, (unsigned long long)hash);
        }
        return 0;
    }
    if (argc < 3) { std::fprintf(stderr, "usage: vct_ingest_dump in.obj outdir | vct_ingest_dump --images file...\n"); return 2; }
    vct::IngestMesh m;
    if (!vct::load_obj(argv[1], m)) { std::fprintf(stderr, "%s\n", m.warnings.c_str()); return 1; }
    const std::string out = argv[2];
    write_file(out + "/vertices.f32", m.vertices.data(), m.vertices.size() * 4);
    write_file(out + "/indices.u32", m.indices.data(), m.indices.size() * 4);
    write_file(out + "/tri_material.i32", m.tri_material.data(), m.tri_material.size() * 4);
    FILE* f = std::fopen((out + "/materials.txt").c_str(), "w");
    if (!f) { std::perror("materials.txt"); return 1; }
    for (const auto& k : m.materials)
        std::fprintf(f, "%s|%s|%s|%s|%s|%s|%s\n", k.name.c_str(), k.diffuse.c_str(), k.specular.c_str(), k.normal.c_str(), k.roughness.c_str(),
                     k.metallic.c_str(), k.alpha.c_str());
    std::fclose(f);
    std::fprintf(stderr, "%s: %zu vertices, %zu triangles, %zu materials, radius %g%s%s\n", argv[1], m.vertices.size() / 14, m.indices.size() / 3,
                 m.materials.size(), m.radius, m.warnings.empty() ? "" : "\n", m.warnings.c_str());
    return 0;
}
