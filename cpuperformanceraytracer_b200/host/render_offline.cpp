// render_offline.cpp -- command-line mirror of ApplicationState::RenderOffline (Application.cpp:400-458):
// load the env texture, allocate + zero the back buffer (Resize, :104-155), initialise the renderer,
// two warm-up frames that DO accumulate (:421-422), NUM_FRAMES_TO_RENDER frames, print total ms and
// ms/frame (:444-452), tone-map and write output_image.bmp (PostprocessAndWriteImageToFile, :381-398).
// It is written against the reference's own entry-point names (demofox_render.h).
//
//   render_offline [--variant v4|v2|simt|v3redo|v3redo0] [--width W --height H --tiles-x X --tiles-y Y]
//                  [--frames N] [--bounces B] [--env file.hdr | --cubemap px nx py ny pz nz]
//                  [--bilinear] [--fast] [--per-frame-calls] [--out out.bmp] [--dump-f32 file]
//                  [--exact-exp] [--sincos-unit-vectors] [--exact-aces] [--exact-gamma]   (global_preprocessor_flags.h:62-65 set to 0)
//                  [--gpus N [--shard spp|tiles] [--combine nccl|peer|fused]] [--device D]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "demofox_render.h"

int main(int argc, char** argv)
{
    std::string variant = "v4", out = "output_image.bmp", dump, env_path;
    char* cube[6] = {0, 0, 0, 0, 0, 0};
    int W = RENDER_BUFFER_PIXEL_WIDTH, H = RENDER_BUFFER_PIXEL_HEIGHT, ntx = NUM_TILES_X, nty = NUM_TILES_Y;
    int frames = NUM_FRAMES_TO_RENDER, bounces = -1;
    bool per_frame_calls = false;
    B200RenderOptions opt;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> char* { return (i + 1 < argc) ? argv[++i] : (char*)""; };
        if (a == "--variant") variant = next();
        else if (a == "--width") W = atoi(next());
        else if (a == "--height") H = atoi(next());
        else if (a == "--tiles-x") ntx = atoi(next());
        else if (a == "--tiles-y") nty = atoi(next());
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--bounces") bounces = atoi(next());
        else if (a == "--env") env_path = next();
        else if (a == "--cubemap") { for (int k = 0; k < 6; k++) cube[k] = next(); opt.use_env_cubemap = 1; }
        else if (a == "--bilinear") opt.use_random_jitter_texture_sampling = 0;
        else if (a == "--fast") opt.math_mode = 1;
        else if (a == "--exact-exp") opt.use_fast_approximate_exp = 0;
        else if (a == "--sincos-unit-vectors") opt.use_unit_vector_rejection_sampling = 0;
        else if (a == "--exact-aces") opt.use_fast_approximate_aces_tonemap = 0;
        else if (a == "--exact-gamma") opt.use_fast_approximate_gamma = 0;
        else if (a == "--per-frame-calls") per_frame_calls = true;
        else if (a == "--gpus") opt.num_gpus = atoi(next());
        else if (a == "--device") opt.device = atoi(next());
        else if (a == "--shard") opt.sharding = std::string(next()) == "tiles" ? 1 : 0;
        else if (a == "--combine") { const std::string v = next(); opt.combine = v == "peer" ? 1 : (v == "fused" ? 2 : 0); }
        else if (a == "--out") out = next();
        else if (a == "--dump-f32") dump = next();
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    // CheckValidSettings, Application.cpp:36-94
    if (ntx <= 0 || nty <= 0 || W % ntx || H % nty || (W / ntx) % 8) {
        std::fprintf(stderr, "invalid settings: need W %% tiles-x == 0, H %% tiles-y == 0, tile width %% 8 == 0\n");
        return 2;
    }
    if (bounces >= 0) opt.v2_num_bounces = opt.v4_num_bounces = bounces;
    texture Texture;
    if (cube[0]) Texture = LoadCubemapTexture(cube);
    else if (!env_path.empty()) Texture = LoadTexture((char*)env_path.c_str());
    const bool needs_env = (variant == "simt") || (variant == "v4") || (variant == "v3redo") || (variant == "v3redo0");
    if (needs_env && !Texture.Data) {
        if (variant == "v4") opt.use_env_map = 0;  // no texture given: constant ambient
        else { std::fprintf(stderr, "--variant %s needs --env file.hdr\n", variant.c_str()); return 2; }
    }
    if (variant == "v3redo0") {  // demofox_path_tracing_v3_redo.cpp with `#define SCENE 0`
        opt.v3_redo_scene = 0;
        variant = "v3redo";
    }
    B200SetRenderOptions(opt);

    const int TW = W / ntx, TH = H / nty;
    std::vector<f32> RenderTarget((size_t)W * H * 3, 0.f);  // Resize: zeroed f32 target + u32 back buffer
    std::vector<u32> Memory((size_t)W * H, 0u);
    auto Render = [&](int n) {
        if (variant == "v4") {
            if (per_frame_calls) for (int f = 0; f < n; f++) DemofoxRenderOptV4(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, Memory.data());
            else DemofoxRenderOptV4Frames(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, Memory.data(), n);
        } else if (variant == "v2") {
            if (per_frame_calls) for (int f = 0; f < n; f++) DemofoxRenderV2(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture);
            else DemofoxRenderV2Frames(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, n);
        } else if (variant == "v3redo") {
            if (per_frame_calls) for (int f = 0; f < n; f++) DemofoxRenderV3Redo(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture);
            else DemofoxRenderV3RedoFrames(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, n);
        } else {
            if (per_frame_calls) for (int f = 0; f < n; f++) DemofoxRenderSimtTextured(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture);
            else DemofoxRenderSimtTexturedFrames(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, n);
        }
    };
    if (variant != "v4" && variant != "v2" && variant != "simt" && variant != "v3redo") { std::fprintf(stderr, "unknown variant\n"); return 2; }

    Render(2);  // two warm-up frames (they accumulate), Application.cpp:421-422
    const auto t0 = std::chrono::steady_clock::now();
    Render(frames);
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    const B200RenderStats st = B200GetRenderStats(variant == "v2" ? 0 : variant == "simt" ? 1 : variant == "v3redo" ? 3 : 2);
    std::printf("Total render time: %.3f ms, average frame time: %.5f ms (%d frames, %dx%d, %.1f Mpaths/s wall; "
                "last kernel %.3f ms on device, %d GPU%s)\n", ms, ms / frames, frames, W, H, (double)W * H * frames / ms * 1e-3,
                st.last_render_ms, opt.num_gpus, opt.num_gpus > 1 ? "s" : "");
    if (opt.num_gpus > 1) std::printf("Cross-GPU combine step: %.3f ms on device\n", st.combine_ms);
    if (!dump.empty()) {
        FILE* f = std::fopen(dump.c_str(), "wb");
        if (f) { std::fwrite(RenderTarget.data(), sizeof(f32), RenderTarget.size(), f); std::fclose(f); }
    }
    // PostprocessAndWriteImageToFile: the tone map lives in the v4 translation unit of the reference
    CopyOutputToFile(RenderTarget.data(), W, H, ntx, nty, TW, TH, 3, Texture, Memory.data());
    WriteImage((char*)out.c_str(), W, H, 4, Memory.data());
    return 0;
}
