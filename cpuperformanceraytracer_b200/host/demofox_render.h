// demofox_render.h -- the reference's render entry points, by name and signature, on top of the
// C ABI (include/b200pt.h).  A caller written against the reference
// (ApplicationState::Render / RenderOffline / PostprocessAndWriteImageToFile,
// Application.cpp:381-477) compiles unchanged against this header and links libdemofox_b200.so.
//
//   reference declaration                                  file:line
//   DemofoxRenderOptV4 / CopyOutputToFile /
//   InitializeGlobalRenderResources /
//   ReinitializeRenderTileData                             demofox_path_tracing_optimization_v4.h:14-26
//   DemofoxRenderV2                                        demofox_path_tracing_v2.h:8-10
//   DemofoxRenderSimtTextured                              demofox_path_tracing_simt_textured.h:8-10
//   DemofoxRenderV3Redo                                    demofox_path_tracing_v3_redo.h
//   struct texture                                         texture.h:6-12
//   LoadTexture / LoadCubemapTexture / WriteImage          asset_loading.h
//
// Like the reference, the entry points keep their state in file-scope statics (frame counter,
// scene, tile table), are not re-entrant, block until the frame is done, and report nothing: a
// failure (no GPU, invalid tiling) prints the reason and aborts, where the reference would
// __debugbreak() (Application.cpp:50-91).
#pragma once
#include <cstdint>

#include "global_preprocessor_flags.h"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef float f32;
typedef double f64;

struct texture {
    f32* Data = 0;
    i32 Width = 0;
    i32 Height = 0;
    i32 Components = 3;
};

// runtime counterparts of the reference's compile-time switches; defaults come from the macros
struct B200RenderOptions {
    int device = 0;
    int math_mode = 0;            // 0 = parity (bit-exact vs the oracle), 1 = fast
    int v2_num_bounces = 4;       // c_numBounces, demofox_path_tracing_v2.cpp:22
    int v4_num_bounces = 8;       // c_numBounces, demofox_path_tracing_optimization_v4.cpp:23 (also v3_redo.cpp:19)
    int use_env_map = USE_ENV_MAP;
    int use_env_cubemap = USE_ENV_CUBEMAP;
    int use_random_jitter_texture_sampling = USE_RANDOM_JITTER_TEXTURE_SAMPLING;
    int output_to_screen = OUTPUT_TO_SCREEN;
    // 0 selects the exact variants.  The two tone-map switches act on CopyOutputToFile / OUTPUT_TO_SCREEN of every variant,
    // the other two on DemofoxRenderOptV4 only (the other variants' sources do not read them)
    int use_fast_approximate_gamma = USE_FAST_APPROXIMATE_GAMMA;
    int use_fast_approximate_aces_tonemap = USE_FAST_APPROXIMATE_ACES_TONEMAP;
    int use_fast_approximate_exp = USE_FAST_APPROXIMATE_EXP;
    int use_unit_vector_rejection_sampling = USE_UNIT_VECTOR_REJECTION_SAMPLING;
    int v3_redo_scene = 1;        // `#define SCENE` of demofox_path_tracing_v3_redo.cpp:379 (1 = the checked-in choice, 0 = the Cornell box)
    // Several GPUs behind the same entry points: the reference fans a render call out to its worker threads below
    // DemofoxRenderOptV4 (..._optimization_v4.cpp:1696-1721); with num_gpus > 1 the call is sharded over devices
    // device .. device + num_gpus - 1 of this process (b200pt_group_*, include/b200pt.h).
    int num_gpus = 1;
    int sharding = 0;             // 0 = frames of a ...Frames call (spp), 1 = tiles of every frame (bit-identical to 1 GPU)
    int combine = 0;              // spp: 0 = NCCL reduce, 1 = the library's own kernel over NVLink peer memory, 2 = fused into the render kernel
};
// must be called before the first render call of a variant (the contexts are created lazily)
void B200SetRenderOptions(const B200RenderOptions& options);

// ---- the reference's entry points ----------------------------------------------------------------
void DemofoxRenderOptV4(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                        i32 TileHeight, i32 NumChannels, texture Texture, void* ScreenBufferData);
void CopyOutputToFile(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                      i32 TileHeight, i32 NumChannels, texture Texture, void* ScreenBufferData);
void InitializeGlobalRenderResources();
void ReinitializeRenderTileData();
void DemofoxRenderV2(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                     i32 TileHeight, i32 NumChannels, texture Texture);
void DemofoxRenderSimtTextured(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                               i32 TileHeight, i32 NumChannels, texture Texture);

void DemofoxRenderV3Redo(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                         i32 TileHeight, i32 NumChannels, texture Texture);

// ---- batched forms: NumFrames consecutive calls of the entry point above in ONE kernel launch and
// ---- one host<->device round trip (what RenderOffline's frame loop amounts to, Application.cpp:426-438)
void DemofoxRenderOptV4Frames(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                              i32 TileHeight, i32 NumChannels, texture Texture, void* ScreenBufferData, i32 NumFrames);
void DemofoxRenderV2Frames(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                           i32 TileHeight, i32 NumChannels, texture Texture, i32 NumFrames);
void DemofoxRenderSimtTexturedFrames(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY,
                                     i32 TileWidth, i32 TileHeight, i32 NumChannels, texture Texture, i32 NumFrames);

void DemofoxRenderV3RedoFrames(f32* BufferOut, i32 BufferWidth, i32 BufferHeight, i32 NumTilesX, i32 NumTilesY, i32 TileWidth,
                               i32 TileHeight, i32 NumChannels, texture Texture, i32 NumFrames);

// ---- asset_loading.h ---------------------------------------------------------------------------------
texture LoadTexture(char* filename);
texture LoadCubemapTexture(char* filename[6]);
void WriteImage(char* filename, i32 width, i32 height, i32 components, void* data);

// device time of the last render call in ms (CUDA events), paths/segments since start-up
struct B200RenderStats {
    double last_render_ms;
    u64 paths, segments, escapes, launches;
    double combine_ms;  // num_gpus > 1: device time of the last cross-GPU combine step
};
B200RenderStats B200GetRenderStats(int variant /*0 = v2, 1 = simt_textured, 2 = opt_v4, 3 = v3_redo*/);
