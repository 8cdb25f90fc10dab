// demofox_render.cpp -- see demofox_render.h.
#include "demofox_render.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "b200pt.h"
#include "image_io.h"

namespace {

B200RenderOptions g_options;
b200pt_context* g_ctx[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // one per reference translation unit (v3_redo: one per SCENE)
b200pt_group* g_group[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // num_gpus > 1: the same, sharded over GPUs
bool g_tile_data_changed = true;                         // tileDataChanged, v4.cpp:1348

[[noreturn]] void die(const char* what, b200pt_context* ctx, int rc)
{
    std::fprintf(stderr, "b200pt: %s failed: %s%s%s\n", what, b200pt_error_string(rc), ctx ? ": " : "",
                 ctx ? b200pt_last_error(ctx) : "");
    std::abort();  // the reference __debugbreak()s on invalid settings (Application.cpp:50-91)
}

b200pt_params params_for(int profile)
{
    b200pt_params p;
    int rc = b200pt_default_params(profile, &p);
    if (rc != B200PT_OK) die("b200pt_default_params", nullptr, rc);
    p.device = g_options.device;
    p.math_mode = g_options.math_mode;
    p.num_bounces = (profile == B200PT_PROFILE_OPT_V4 || profile == B200PT_PROFILE_V3_REDO || profile == B200PT_PROFILE_V3_REDO_SCENE0)
                        ? g_options.v4_num_bounces : g_options.v2_num_bounces;
    if (profile == B200PT_PROFILE_OPT_V4) {
        p.env_kind = !g_options.use_env_map ? B200PT_ENV_NONE : (g_options.use_env_cubemap ? B200PT_ENV_CUBEMAP : B200PT_ENV_EQUIRECT);
        p.env_sampler = g_options.use_random_jitter_texture_sampling ? B200PT_SAMPLER_RANDOM : B200PT_SAMPLER_BILINEAR;
        p.output_to_screen = g_options.output_to_screen;
        p.exact_exp = !g_options.use_fast_approximate_exp;
        p.sincos_unit_vectors = !g_options.use_unit_vector_rejection_sampling;
    }
    // CopyOutputToFile / OutputToScreen are the v4 translation unit's for every renderer (Application.cpp:381-398)
    p.exact_tonemap = (g_options.use_fast_approximate_aces_tonemap ? 0 : B200PT_TONEMAP_EXACT_ACES) |
                      (g_options.use_fast_approximate_gamma ? 0 : B200PT_TONEMAP_EXACT_GAMMA);
    return p;
}

b200pt_context* context_for(int profile)
{
    if (g_ctx[profile]) return g_ctx[profile];
    const b200pt_params p = params_for(profile);
    const int rc = b200pt_create(&p, &g_ctx[profile]);
    if (rc != B200PT_OK) die("b200pt_create (a B200 is required; there is no CPU fallback)", nullptr, rc);
    return g_ctx[profile];
}

b200pt_group* group_for(int profile)
{
    if (g_group[profile]) return g_group[profile];
    const b200pt_params p = params_for(profile);
    int32_t devices[16];
    const int n = g_options.num_gpus > 16 ? 16 : g_options.num_gpus;
    for (int i = 0; i < n; i++) devices[i] = g_options.device + i;
    const int rc = b200pt_group_create(&p, devices, n, g_options.sharding ? B200PT_SHARD_TILES : B200PT_SHARD_SPP,
                                       g_options.combine == 2 ? B200PT_COMBINE_FUSED : (g_options.combine ? B200PT_COMBINE_PEER : B200PT_COMBINE_NCCL),
                                       &g_group[profile]);
    if (rc != B200PT_OK) die("b200pt_group_create (num_gpus B200s are required; there is no CPU fallback)", nullptr, rc);
    return g_group[profile];
}

void render(int profile, f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture,
            void* ScreenBufferData, i32 NumFrames)
{
    b200pt_texture t;
    t.Data = Texture.Data;
    t.Width = Texture.Width;
    t.Height = Texture.Height;
    t.Components = Texture.Components;
    if (g_options.num_gpus > 1) {
        b200pt_group* grp = group_for(profile);
        const int rc = b200pt_group_render_host(grp, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, t, ScreenBufferData, NumFrames);
        if (rc != B200PT_OK) {
            std::fprintf(stderr, "b200pt: b200pt_group_render_host failed: %s: %s\n", b200pt_error_string(rc), b200pt_group_last_error(grp));
            std::abort();
        }
    } else {
        b200pt_context* ctx = context_for(profile);
        const int rc = b200pt_render_host(ctx, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, t, ScreenBufferData, NumFrames);
        if (rc != B200PT_OK) die("b200pt_render_host", ctx, rc);
    }
    g_tile_data_changed = false;
}

}  // namespace

void B200SetRenderOptions(const B200RenderOptions& options) { g_options = options; }

void InitializeGlobalRenderResources() { context_for(B200PT_PROFILE_OPT_V4); }

void ReinitializeRenderTileData() { g_tile_data_changed = true; }  // geometry is re-derived from every call's arguments

void DemofoxRenderOptV4(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture,
                        void* ScreenBufferData)
{
    render(B200PT_PROFILE_OPT_V4, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, ScreenBufferData, 1);
}

void DemofoxRenderOptV4Frames(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture,
                              void* ScreenBufferData, i32 NumFrames)
{
    render(B200PT_PROFILE_OPT_V4, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, ScreenBufferData, NumFrames);
}

void DemofoxRenderV2(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture)
{
    render(B200PT_PROFILE_V2, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, nullptr, 1);
}

void DemofoxRenderV2Frames(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture,
                           i32 NumFrames)
{
    render(B200PT_PROFILE_V2, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, nullptr, NumFrames);
}

void DemofoxRenderSimtTextured(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture)
{
    render(B200PT_PROFILE_SIMT_TEXTURED, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, nullptr, 1);
}

void DemofoxRenderSimtTexturedFrames(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels,
                                     texture Texture, i32 NumFrames)
{
    render(B200PT_PROFILE_SIMT_TEXTURED, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels, Texture, nullptr, NumFrames);
}

void DemofoxRenderV3Redo(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture)
{
    render(g_options.v3_redo_scene == 0 ? B200PT_PROFILE_V3_REDO_SCENE0 : B200PT_PROFILE_V3_REDO, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels,
           Texture, nullptr, 1);
}

void DemofoxRenderV3RedoFrames(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture Texture,
                               i32 NumFrames)
{
    render(g_options.v3_redo_scene == 0 ? B200PT_PROFILE_V3_REDO_SCENE0 : B200PT_PROFILE_V3_REDO, BufferOut, W, H, NTX, NTY, TW, TH, NumChannels,
           Texture, nullptr, NumFrames);
}

// CopyOutputToFile (v4.cpp:1729-1760): tone-maps the f32 buffer into ScreenBufferData
// (A=FF | B<<16 | G<<8 | R, row-major) and, like the reference, bumps the frame counter.
void CopyOutputToFile(f32* BufferOut, i32 W, i32 H, i32 NTX, i32 NTY, i32 TW, i32 TH, i32 NumChannels, texture, void* ScreenBufferData)
{
    b200pt_context* ctx = context_for(B200PT_PROFILE_OPT_V4);
    if (NumChannels != 3 || TW * NTX != W || TH * NTY != H) die("CopyOutputToFile (tiling)", ctx, B200PT_ERR_INVALID_ARGUMENT);
    int32_t frame = 0;
    b200pt_get_frame_counter(ctx, &frame);
    int rc = b200pt_resize(ctx, W, H, NTX, NTY);  // no-op reallocation when the size is unchanged; zeroes, so re-upload
    if (rc != B200PT_OK) die("b200pt_resize", ctx, rc);
    b200pt_set_frame_counter(ctx, frame);
    rc = b200pt_upload_target(ctx, BufferOut);
    if (rc != B200PT_OK) die("b200pt_upload_target", ctx, rc);
    rc = b200pt_resolve_ldr(ctx, static_cast<uint32_t*>(ScreenBufferData), B200PT_LDR_FILE_RGBA, 1);
    if (rc != B200PT_OK) die("b200pt_resolve_ldr", ctx, rc);
}

texture LoadTexture(char* filename)
{
    texture t;
    b200pt::HostImage img;
    std::string err;
    bool ok = false;
    try {
        ok = b200pt::LoadRadianceHDR(filename, &img, &err);
    } catch (const std::exception& e) {
        err = e.what();
    }
    if (!ok) {
        std::fprintf(stderr, "LoadTexture(%s): %s\n", filename, err.c_str());
        return t;  // Data == 0, like a failed stbi_loadf
    }
    t.Data = static_cast<f32*>(std::malloc(img.rgb.size() * sizeof(f32)));
    std::memcpy(t.Data, img.rgb.data(), img.rgb.size() * sizeof(f32));
    t.Width = img.width;
    t.Height = img.height;
    t.Components = 3;
    return t;
}

texture LoadCubemapTexture(char* filename[6])
{
    texture t;
    b200pt::HostImage img;
    std::string err, paths[6];
    for (int i = 0; i < 6; i++) paths[i] = filename[i];
    bool ok = false;
    try {
        ok = b200pt::LoadCubemapAtlas(paths, &img, &err);
    } catch (const std::exception& e) {
        err = e.what();
    }
    if (!ok) {
        std::fprintf(stderr, "LoadCubemapTexture: %s\n", err.c_str());
        return t;
    }
    t.Data = static_cast<f32*>(std::malloc(img.rgb.size() * sizeof(f32)));
    std::memcpy(t.Data, img.rgb.data(), img.rgb.size() * sizeof(f32));
    t.Width = img.width;
    t.Height = img.height;
    t.Components = 3;
    return t;
}

void WriteImage(char* filename, i32 width, i32 height, i32 components, void* data)
{
    std::string err;
    if (components != 4 || !b200pt::WriteBMP32(filename, width, height, static_cast<const uint32_t*>(data), &err))
        std::fprintf(stderr, "WriteImage(%s): %s\n", filename, components != 4 ? "only 4-component buffers are supported" : err.c_str());
}

B200RenderStats B200GetRenderStats(int variant)
{
    B200RenderStats s{};
    if (variant == 3 && g_options.v3_redo_scene == 0) variant = 4;
    if (variant < 0 || variant > 4 || (!g_ctx[variant] && !g_group[variant])) return s;
    b200pt_counters c;
    if (g_group[variant] ? b200pt_group_get_counters(g_group[variant], &c, &s.combine_ms) == B200PT_OK
                         : b200pt_get_counters(g_ctx[variant], &c) == B200PT_OK) {
        s.last_render_ms = c.last_render_ms;
        s.paths = c.paths;
        s.segments = c.segments;
        s.escapes = c.escapes;
        s.launches = c.launches;
    }
    return s;
}
