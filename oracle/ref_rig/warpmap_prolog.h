// warpmap_prolog.h — scaffolding that lets the CPU half of the reference's warp-map generation
// (/root/reference/src/Application.cpp:311-370: the three partial-sum tables and the low/high weight table,
// plain C++ in the middle of Application::render) compile and run here.  The lines are piped from where they lie
// between this prolog and warpmap_epilog.inc (oracle/Makefile); only the binary lands in oracle/_ref/.
// What the block needs from its surroundings: warpDim (src/Application.h:178), the occupancy image it has just
// read back with glGetTextureImage (`warpTexture`, :303-309), glm::ivec3 and two Settings fields.
#include <cstdio>
#include <cstdlib>

typedef unsigned int GLuint;
namespace glm { struct ivec3 { int x, y, z; }; }
static const int warpDim = 32;
static struct { float warpTextureHighResolution, warpTextureLowResolution; } settings;
static GLuint warpTexture[warpDim][warpDim][warpDim];

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: warpmap_cpu occupancy.u32 high low\n"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(warpTexture, sizeof(GLuint), warpDim * warpDim * warpDim, f) != (size_t)warpDim * warpDim * warpDim) return 3;
    std::fclose(f);
    settings.warpTextureHighResolution = (float)std::atof(argv[2]);
    settings.warpTextureLowResolution = (float)std::atof(argv[3]);
    {
