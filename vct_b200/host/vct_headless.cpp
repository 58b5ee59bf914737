// vct_headless — headless driver over vct_host::Application (the reference's main.cpp render loop without a window,
// src/main.cpp:229-240): loads a VCTS scene file — or an OBJ through the library's own ingest, with the reference's two
// lights, like Application::init (src/Application.cpp:96-136) —, renders N frames through the C ABI, prints the GLBufferedTimer-style
// pass times and the VoxelizeInfo counters as one JSON line, and dumps the last frame as a binary PPM.
//
//   vct_headless scene.vcts|mesh.obj [--scale s] [--resources dir] [--dim 256] [--levels 6] [--size 1920x1080] [--shadow 4096] [--frames 3]
//                [--camera x y z yaw pitch | --eye x y z --front fx fy fz] [--volume min max] [--center x y z]
//                [--no-reflections] [--atomic-max] [--msaa] [--voxelize-multiplier M] [--tesselation] [--tesselation-warp] [--warp-texture] [--warp-voxels] [--temporal] [--fused] [--out frame.ppm]
//                [--gpus N] (one process driving N devices: z-slab sharded frames, image on device 0) [--track-camera]
// Exit status: 0 ok, 1 a pass reported an error (message on stderr), 2 usage.  There is no CPU fallback: without a
// CUDA device vct_create fails and the driver exits 1.
#include <chrono>
#include <cstdlib>

#include "vct_host.hpp"

using namespace vct_host;

static bool write_ppm(const char* path, const std::vector<uint8_t>& rgba, int w, int h) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    for (int y = h - 1; y >= 0; --y)                                    // glReadPixels rows are bottom-up
        for (int x = 0; x < w; ++x) std::fwrite(&rgba[((size_t)y * w + x) * 4], 1, 3, f);
    std::fclose(f);
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2 || !std::strcmp(argv[1], "--help") || !std::strcmp(argv[1], "-h")) {
        std::fprintf(argc < 2 ? stderr : stdout,
                     "usage: vct_headless scene.vcts|mesh.obj [--scale s] [--resources dir] [--dim D] [--levels L] [--size WxH] [--shadow S] [--frames N]\n"
                     "       [--camera x y z yaw pitch | --eye x y z --front fx fy fz] [--volume min max] [--center x y z]\n"
                     "       [--no-reflections] [--atomic-max] [--msaa] [--voxelize-multiplier M] [--tesselation] [--tesselation-warp] [--warp-texture] [--warp-voxels] [--temporal] [--fused] [--out frame.ppm]\n"
                     "       [--gpus N] [--track-camera]\n"
                     "       [--view voxels|normals|dominant-axis|occlusion|indirect|reflections|material-diffuse|material-roughness|material-metallic] [--miplevel x]\n");
        return argc < 2 ? 2 : 0;
    }
    Application app;
    app.width = 1920; app.height = 1080;
    app.camera.position = {5, 1, 0}; app.camera.yaw = 180.0f;           // reference start pose, Application.cpp:139-141
    int frames = 3, shadow = Application::SHADOWMAP_WIDTH, gpus = 0; bool fused = false; const char* out = nullptr;
    float scale = 1.0f; std::string resources;
    auto need = [&](int i, int n) { if (i + n >= argc) { std::fprintf(stderr, "missing value after %s\n", argv[i]); std::exit(2); } };
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--dim") { need(i, 1); app.vct.voxelDim = std::atoi(argv[++i]); }
        else if (a == "--scale") { need(i, 1); scale = (float)std::atof(argv[++i]); }
        else if (a == "--resources") { need(i, 1); resources = argv[++i]; }
        else if (a == "--levels") { need(i, 1); app.vct.voxelLevels = std::atoi(argv[++i]); }
        else if (a == "--size") { need(i, 1); if (std::sscanf(argv[++i], "%dx%d", &app.width, &app.height) != 2) return 2; }
        else if (a == "--shadow") { need(i, 1); shadow = std::atoi(argv[++i]); }
        else if (a == "--frames") { need(i, 1); frames = std::atoi(argv[++i]); }
        else if (a == "--camera") { need(i, 5); app.camera.position = {(float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3])}; app.camera.yaw = (float)std::atof(argv[i + 4]); app.camera.pitch = (float)std::atof(argv[i + 5]); i += 5; }
        else if (a == "--eye") { need(i, 3); app.camera.position = {(float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3])}; i += 3; }
        else if (a == "--front") { need(i, 3); app.camera.hasFront = true; app.camera.frontOverride = {(float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3])}; i += 3; }
        else if (a == "--volume") { need(i, 2); const float lo = (float)std::atof(argv[i + 1]), hi = (float)std::atof(argv[i + 2]); app.vct.min = {lo, lo, lo}; app.vct.max = {hi, hi, hi}; i += 2; }
        else if (a == "--center") { need(i, 3); app.vct.center = {(float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3])}; i += 3; }
        else if (a == "--no-reflections") app.settings.enableReflections = false;
        else if (a == "--atomic-max") app.settings.voxelizeAtomicMax = true;
        else if (a == "--voxelize-multiplier") { need(i, 1); app.settings.voxelizeMultiplier = (float)std::atof(argv[++i]); }   // Settings::voxelizeMultiplier
        else if (a == "--msaa") app.settings.conservativeRasterization = Settings::MSAA;   // Settings::conservativeRasterization = MSAA (the reference's default)
        else if (a == "--tesselation") app.settings.voxelizeTesselation = true;   // the reference's default voxeliser (with --atomic-max: its default frame)
        else if (a == "--tesselation-warp") app.settings.voxelizeTesselationWarp = true;   // Settings::voxelizeTesselationWarp (with --tesselation: the frustum-aligned grid)
        else if (a == "--warp-texture") app.settings.warpTexture = true;
        else if (a == "--warp-voxels") app.settings.warpVoxels = true;
        else if (a == "--temporal") app.settings.temporalFilterRadiance = true;
        else if (a == "--view") {                                           // one of the reference's debug toggles (Overlay.cpp), by name
            need(i, 1); const std::string v = argv[++i];
            Settings& st = app.settings;
            if (v == "debug-voxels") st.debugVoxels = true;                  // Settings::debugVoxels: cubes (pass-by-pass render only)
            else if (v == "voxels") st.drawVoxels = true; else if (v == "voxel-normals") st.drawVoxels = st.drawNormals = true;
            else if (v == "warp-texture") st.drawVoxels = st.debugWarpTexture = true; else if (v == "warp-texture-tc") st.drawVoxels = st.debugWarpTexture = st.toggle = true; else if (v == "normals") st.drawNormals = true; else if (v == "dominant-axis") st.drawDominantAxis = true;
            else if (v == "occlusion") st.debugOcclusion = true; else if (v == "indirect") st.debugIndirect = true; else if (v == "reflections") st.debugReflections = true;
            else if (v == "material-diffuse") st.debugMaterialDiffuse = true; else if (v == "material-roughness") st.debugMaterialRoughness = true;
            else if (v == "material-metallic") st.debugMaterialMetallic = true;
            else { std::fprintf(stderr, "unknown view %s\n", v.c_str()); return 2; }
        }
        else if (a == "--miplevel") { need(i, 1); app.settings.miplevel = (float)std::atof(argv[++i]); }
        else if (a == "--fused") fused = true;
        else if (a == "--gpus") { need(i, 1); gpus = std::atoi(argv[++i]); fused = true; }   // one process, N devices: the sharded frame is vct_frame's
        else if (a == "--track-camera") app.settings.voxelTrackCamera = true;
        else if (a == "--out") { need(i, 1); out = argv[++i]; }
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    Scene scene;
    const std::string input = argv[1];
    if (input.size() > 4 && input.compare(input.size() - 4, 4, ".obj") == 0) {
        if (resources.empty()) resources = input.substr(0, input.find_last_of('/') + 1);
        if (!scene.addObj(argv[1], resources.c_str(), scale)) return 1;
        scene.addReferenceLights();
    } else if (!scene.load(argv[1])) return 1;
    if (scene.lights.empty()) { std::fprintf(stderr, "[ERROR] scene has no light (Application.cpp:125-130 needs the shadow-casting main light)\n"); return 1; }
    if (!app.init(&scene, shadow, 0, gpus)) return 1;
    bool ok = true;
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) ok &= fused ? app.renderFused(1.0f / 60.0f) : app.render(1.0f / 60.0f);
    ok &= vct_sync(app.vct.ctx) == 0;
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (frames > 0 ? frames : 1);
    std::vector<uint8_t> rgba;
    ok &= app.readPixels(rgba);
    unsigned long long sum = 0, fnv = 1469598103934665603ull; for (uint8_t b : rgba) { sum += b; fnv = (fnv ^ b) * 1099511628211ull; }
    std::printf("{\"frames\": %d, \"wall_ms_per_frame\": %.4f, \"total_fragments\": %u, \"unique_voxels\": %u, \"max_fragments_per_voxel\": %u, "
                "\"image_byte_sum\": %llu, \"image_fnv1a\": \"%016llx\", \"gpus\": %d, \"fused\": %s, \"timers_ms\": {\"voxelize\": %.4f, \"shadowmap\": %.4f, \"radiance\": %.4f, \"mipmap\": %.4f, \"render\": %.4f, \"total\": %.4f}, \"ok\": %s}\n",
                frames, wall_ms, app.voxelizeInfo.total_fragments, app.voxelizeInfo.unique_voxels, app.voxelizeInfo.max_fragments_per_voxel,
                sum, fnv, gpus > 1 ? gpus : 1, fused ? "true" : "false", app.timers.voxelize, app.timers.shadowmap, app.timers.radiance, app.timers.mipmap, app.timers.render, app.timers.total, ok ? "true" : "false");
    if (out && !write_ppm(out, rgba, app.width, app.height)) { std::fprintf(stderr, "[ERROR] cannot write %s\n", out); ok = false; }
    return ok ? 0 : 1;
}
