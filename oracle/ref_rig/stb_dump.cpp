// stb_dump file... — decodes each file with the reference's own vendored stb_image (ext/include/stb_image.h, compiled in
// place by oracle/Makefile, never copied) exactly as GLHelper::createTextureFromImage calls it
// (src/Graphics/GLHelper.cpp:172: stbi_load(name, &w, &h, &channels, STBI_default)) and prints, per file, one line
// "<width> <height> <channels> <fnv1a-64 of the pixel bytes>" — or, with --raw (one file), a 12-byte header
// {i32 width, height, channels} followed by the pixel bytes on stdout.  Test infrastructure (tests/test_ingest.py).
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

int main(int argc, char** argv) {
    const bool raw = argc > 1 && !std::strcmp(argv[1], "--raw");
    for (int i = raw ? 2 : 1; i < argc; ++i) {
        int w = 0, h = 0, ch = 0;
        unsigned char* px = stbi_load(argv[i], &w, &h, &ch, STBI_default);
        if (!px) { if (raw) return 1; std::printf("0 0 0 0\n"); continue; }
        if (raw) { int hdr[3] = {w, h, ch}; std::fwrite(hdr, 4, 3, stdout); std::fwrite(px, 1, (size_t)w * h * ch, stdout); stbi_image_free(px); return 0; }
        uint64_t hash = 1469598103934665603ull;
        for (size_t k = 0; k < (size_t)w * h * ch; ++k) { hash ^= px[k]; hash *= 1099511628211ull; }
        std::printf("%d %d %d %016llx\n", w, h, ch, (unsigned long long)hash);
        stbi_image_free(px);
    }
    return 0;
}
