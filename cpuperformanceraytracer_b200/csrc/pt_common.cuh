// pt_common.cuh -- shared types of the B200 path-tracing kernels (host + device).
// Reference paths are relative to /root/reference/CPUPerformanceRayTracer/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200pt {

struct v3 {
    float x, y, z;
};

// ---- scene tables, filled on the host (host/scene_setup.cpp) with the reference's own
// ---- float expressions and handed to the kernel as a __grid_constant__ parameter ------------

// Cornell box of demofox_path_tracing_v2.cpp:320-454 / ..._simt_textured.cpp:278-385.
// `n` is normalize(cross(c - a, c - b)), the per-ray expression of v2.cpp:166 hoisted to the
// host: it only depends on the vertices, so the hoist is bit-neutral.
struct LegacyQuad {
    v3 a, b, c, d, n;
};
struct LegacyMaterial {
    v3 albedo, emissive, specularColor;
    float percentSpecular, roughness;
};
constexpr int kCornellQuads = 6;
constexpr int kCornellSpheres = 3;
constexpr int kCornellObjects = kCornellQuads + kCornellSpheres;
// The Cornell box is hard-coded in the reference (literals of v2.cpp:328-331,345-348,362-365,379-382,
// 396-399,413-416 plus sceneTranslation :323).  One table feeds both the host-built CornellScene and
// the kernel's compile-time specialisation, in which the vertex coordinates are immediates so that
// the 72 "vertex - rayPos" subtractions of a trace collapse to the 15 distinct ones.
constexpr float kCornellTranslation[3] = {0.0f, 0.0f, 10.0f};
constexpr float kCornellQuadVerts[kCornellQuads][4][3] = {
    {{-12.6f, -12.6f, 25.0f}, {12.6f, -12.6f, 25.0f}, {12.6f, 12.6f, 25.0f}, {-12.6f, 12.6f, 25.0f}},        // back wall
    {{-12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 15.0f}, {-12.6f, -12.45f, 15.0f}},  // floor
    {{-12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 15.0f}, {-12.6f, 12.5f, 15.0f}},          // ceiling
    {{-12.5f, -12.6f, 25.0f}, {-12.5f, -12.6f, 15.0f}, {-12.5f, 12.6f, 15.0f}, {-12.5f, 12.6f, 25.0f}},      // left wall
    {{12.5f, -12.6f, 25.0f}, {12.5f, -12.6f, 15.0f}, {12.5f, 12.6f, 15.0f}, {12.5f, 12.6f, 25.0f}},          // right wall
    {{-5.0f, 12.4f, 22.5f}, {5.0f, 12.4f, 22.5f}, {5.0f, 12.4f, 17.5f}, {-5.0f, 12.4f, 17.5f}}};             // light
struct CornellScene {
    LegacyQuad quad[kCornellQuads];
    float4 sphere[kCornellSpheres];  // xyz + radius
    LegacyMaterial mat[kCornellObjects];
};

// Scene of ..._optimization_v4.cpp:1403-1496 with PrecomputeQuadData (:269-319).
struct V4Quad {
    v3 V0, NxV01, NxV20, NxV02, NxV30, n;
};
struct V4Material {
    v3 albedo, emissive, specularColor, refractionColor;
    float specularChance, specularRoughness, IOR, refractionChance, refractionRoughness;
};
constexpr int kV4Quads = 4;    // the built-in scene of InitializeScene
constexpr int kV4Spheres = 7;
constexpr int kV4Objects = kV4Quads + kV4Spheres;
constexpr int kV4MaxObjects = 12;  // MAX_OBJECTS / MAX_MATERIALS, v4.cpp:327-328: one material per object
// `struct Scene` (v4.cpp:364-378) + `struct Camera` (:380-386) as data: the built-in scene, or whatever
// b200pt_set_scene_v4 installed.  Object (= material) index: quads first, then spheres (v4.cpp:702-715).
struct V4Scene {
    int numQuads, numSpheres;
    V4Quad quad[kV4MaxObjects];
    float4 sphere[kV4MaxObjects];
    V4Material mat[kV4MaxObjects];
    v3 cameraPosition;
    float cameraDistance;
};

// The built-in v4 scene as compile-time knowledge (kernels specialised with STATIC):
//  * sphere i has centre (-18 + 6 i, -8, 10) and radius 2.8 (v4.cpp:1474-1495): m.y, m.z and the inner terms
//    of dot(m, d) / dot(m, m) are shared by the seven tests;
//  * the quads are axis-aligned, so the host-computed vectors of V4Quad have exact zeros.  Bit k of a mask =
//    component k is nonzero, order n, NxV01, NxV20, NxV02, NxV30.  Products with those zeros are +-0 for a
//    finite ray, so a dot product reduces to its nonzero terms (same nesting, same roundings; only the sign
//    of an exactly-zero result can differ, which no comparison below observes).
//    b200pt_create verifies the masks against the scene it built and falls back to the generic kernel.
constexpr float kV4SphereY = -8.0f, kV4SphereZ = 10.0f, kV4SphereRadius = 2.8f;
__host__ __device__ constexpr float v4_sphere_x(int i) { return -18.0f + 6.0f * (float)i; }
// Cornell spheres (v2.cpp:429-447): centres (-9 | 0 | 9, -9.5, 20) + sceneTranslation, radius 3
__host__ __device__ constexpr float cornell_sphere_x(int i) { return -9.0f + 9.0f * (float)i; }
constexpr float kCornellSphereY = -9.5f, kCornellSphereZ = 30.0f, kCornellSphereRadius = 3.0f;
constexpr int kV4QuadMasks[kV4Quads][5] = {{2, 4, 5, 5, 1}, {4, 2, 3, 3, 1}, {2, 4, 5, 5, 1}, {2, 4, 5, 5, 1}};
inline int v3_nonzero_mask(const v3& v) { return (v.x != 0.f ? 1 : 0) | (v.y != 0.f ? 2 : 0) | (v.z != 0.f ? 4 : 0); }
// true when `s` is the built-in scene the STATIC v4 kernels assume
inline bool v4_scene_matches_static_tables(const V4Scene& s)
{
    if (s.numQuads != kV4Quads || s.numSpheres != kV4Spheres) return false;
    for (int i = 0; i < kV4Quads; i++) {
        const V4Quad& q = s.quad[i];
        const int m[5] = {v3_nonzero_mask(q.n), v3_nonzero_mask(q.NxV01), v3_nonzero_mask(q.NxV20), v3_nonzero_mask(q.NxV02),
                          v3_nonzero_mask(q.NxV30)};
        for (int k = 0; k < 5; k++)
            if (m[k] != kV4QuadMasks[i][k]) return false;
    }
    for (int i = 0; i < kV4Spheres; i++)
        if (s.sphere[i].x != v4_sphere_x(i) || s.sphere[i].y != kV4SphereY || s.sphere[i].z != kV4SphereZ || s.sphere[i].w != kV4SphereRadius)
            return false;
    return true;
}

// the legacy kernels' immediate sphere data (Cornell: v2 / simt_textured) against a host-built scene
inline bool cornell_spheres_match_static_tables(const float4* sphere)
{
    for (int i = 0; i < 3; i++)
        if (sphere[i].x != cornell_sphere_x(i) || sphere[i].y != kCornellSphereY || sphere[i].z != kCornellSphereZ ||
            sphere[i].w != kCornellSphereRadius)
            return false;
    return true;
}
inline bool v3redo_spheres_match_static_tables(const float4* sphere)
{
    for (int i = 0; i < kV4Spheres; i++)
        if (sphere[i].x != v4_sphere_x(i) || sphere[i].y != kV4SphereY || sphere[i].z != kV4SphereZ || sphere[i].w != kV4SphereRadius)
            return false;
    return true;
}

// Scene of demofox_path_tracing_v3_redo.cpp, SCENE 1 (:485-600): the v4 geometry tested with the
// legacy (ScalarTriple) quad test, exact-arithmetic Fresnel materials, a striped backdrop whose
// albedo is computed at the hit (:511-515), GetZeroedMaterial IOR = 1 for the quads (:155-168).
constexpr int kV3Quads = 4;
constexpr int kV3Spheres = 7;
constexpr int kV3Objects = kV3Quads + kV3Spheres;
constexpr int kV3BackdropQuad = 1;
struct V3RedoScene {
    LegacyQuad quad[kV3Quads];
    float4 sphere[kV3Spheres];
    V4Material mat[kV3Objects];
    v3 cameraPosition;
};
// vertex literals of v3_redo.cpp:486-489,505-508,520-523,535-538; the backdrop is not translated
constexpr float kV3QuadVerts[kV3Quads][4][3] = {
    {{-25.0f, -12.5f, 5.0f}, {25.0f, -12.5f, 5.0f}, {25.0f, -12.5f, -5.0f}, {-25.0f, -12.5f, -5.0f}},
    {{-25.0f, -1.5f, 5.0f}, {25.0f, -1.5f, 5.0f}, {25.0f, -10.5f, 5.0f}, {-25.0f, -10.5f, 5.0f}},
    {{-7.5f, 12.5f, 5.0f}, {7.5f, 12.5f, 5.0f}, {7.5f, 12.5f, -5.0f}, {-7.5f, 12.5f, -5.0f}},
    {{-5.0f, 12.4f, 2.5f}, {5.0f, 12.4f, 2.5f}, {5.0f, 12.4f, -2.5f}, {-5.0f, 12.4f, -2.5f}}};
constexpr float kV3Translation[kV3Quads][3] = {{0.0f, 0.0f, 10.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 10.0f}, {0.0f, 0.0f, 10.0f}};

// The same renderer compiled with `#define SCENE 0` (v3_redo.cpp:379, :392-479, :530-580): the Cornell box of v2 (no
// translation: the camera sits at (0, 0, 40) and looks down -z), a light far outside the box, three spheres with
// Fresnel-specular materials.
constexpr int kV3S0Quads = 6;
constexpr int kV3S0Spheres = 3;
constexpr int kV3S0Objects = kV3S0Quads + kV3S0Spheres;
constexpr float kV3S0QuadVerts[kV3S0Quads][4][3] = {
    {{-12.6f, -12.6f, 25.0f}, {12.6f, -12.6f, 25.0f}, {12.6f, 12.6f, 25.0f}, {-12.6f, 12.6f, 25.0f}},        // back wall
    {{-12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 25.0f}, {12.6f, -12.45f, 15.0f}, {-12.6f, -12.45f, 15.0f}},  // floor
    {{-12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 25.0f}, {12.6f, 12.5f, 15.0f}, {-12.6f, 12.5f, 15.0f}},          // ceiling
    {{-12.5f, -12.6f, 25.0f}, {-12.5f, -12.6f, 15.0f}, {-12.5f, 12.6f, 15.0f}, {-12.5f, 12.6f, 25.0f}},      // left wall
    {{12.5f, -12.6f, 25.0f}, {12.5f, -12.6f, 15.0f}, {12.5f, 12.6f, 15.0f}, {12.5f, 12.6f, 25.0f}},          // right wall
    {{-5.0f, 12.4f, -22.5f}, {5.0f, 12.4f, -22.5f}, {5.0f, 12.4f, -17.5f}, {-5.0f, 12.4f, -17.5f}}};         // light
constexpr float kV3S0SphereY = -9.5f, kV3S0SphereZ = 20.0f, kV3S0SphereRadius = 3.0f;  // x = cornell_sphere_x(i)
struct V3RedoScene0 {
    LegacyQuad quad[kV3S0Quads];
    float4 sphere[kV3S0Spheres];
    V4Material mat[kV3S0Objects];
    v3 cameraPosition;
};
inline bool v3redo0_spheres_match_static_tables(const float4* sphere)
{
    for (int i = 0; i < kV3S0Spheres; i++)
        if (sphere[i].x != cornell_sphere_x(i) || sphere[i].y != kV3S0SphereY || sphere[i].z != kV3S0SphereZ || sphere[i].w != kV3S0SphereRadius)
            return false;
    return true;
}

constexpr int kMaxCullRects = 12;
constexpr int kMaxScatterRanks = 16;

struct DeviceCounters {
    unsigned long long segments;
    unsigned long long escapes;
    unsigned long long culled;  // segments of camera-culled pixels: counted in `segments`, but no scene trace ran
};

// One launch = `nframes` consecutive render calls of the reference over the whole image.
struct RenderParams {
    float* target;        // tile-major SoA8 accumulation buffer (RenderTile, v4.cpp:1189-1252)
    uint32_t* rng_out;    // optional: final RNG state per pixel, row-major (debug/parity)
    uint32_t* screen;     // optional: row-major u32 image, tone-mapped in the kernel tail (OUTPUT_TO_SCREEN)
    int screen_mode;      // bit 0: 0 = file packing, 1 = screen packing; bit 1: exact ACES curve
    int v4_flags;         // generic OPT_V4 kernels only: bit 0 exact exp, bit 1 sin/cos unit vectors (b200pt_params)
    int* work_counter;    // atomic work-item counter: replaces work_queue.cpp's ring + CAS pop
    // ACCUM_SUM only, fused render + reduce-scatter of a multi-GPU group: a finished pixel's sum is stored straight into the
    // staging slot of the rank that OWNS its part of the image (NVLink peer memory) instead of the local target.
    // scatter_stage[o] = owner o's stage + this rank's slot; scatter_gpo = SoA8 groups per owner (0 = off)
    float* scatter_stage[kMaxScatterRanks];
    int scatter_gpo;
    const int* item_order; // optional: the k-th pulled work item is item_order[k] (expensive items first: see b200pt_capi.cu)
    DeviceCounters* counters;
    cudaTextureObject_t env;  // RGBA32F linear texture, texel t = reference float index 3t
    int env_w, env_h;
    int width, height;
    int tile_w, tile_h, num_tiles_x;
    int groups_per_tile_row;  // tile_w / 8
    int groups_per_tile;      // tile_w / 8 * tile_h
    int num_groups;           // SoA8 groups rendered by this launch (whole image, or a band of tile rows)
    int group_offset;         // first group of the band (tile-shard: a rank's tile rows are contiguous)
    int num_items;            // work items of this launch (ceil(num_groups / 4): one warp = 4 groups = 32 pixels; fewer with a tile stride)
    int tile_mod, tile_rem;   // tile_mod > 1: only tiles with FlatTileIndex % tile_mod == tile_rem (the items come from item_order)
    int order_domain_items;   // items the pull-order builder classifies: num_items, or all items of the image with a tile stride
    int block_items;          // tile_h % 4 == 0: an item is an 8x4 pixel block (4 groups stacked in y) instead of a 32x1 strip
    int first_frame;          // 1-based iFrame of the first render call in this launch
    int nframes;
    int num_bounces;
    float cameraDistance;     // 1 / tan(FOV/2), computed on the host like v2.cpp:546
    float rcp_width, rcp_height;  // RN(1/W), RN(1/H), host-computed
    float aspect;             // RN(W/H): aspectRatio of mainImage (v2.cpp:556), host-computed
    int res_div_exact;        // W and H have <= 16 significant bits: x / W == fma(fma(-q, W, x), rcp, q), q = x * rcp
    // conservative screen-space bounds of the scene's primitives in fragCoord space
    // (x0, y0, x1, y1; y = flipped row): a pixel whose jitter footprint overlaps none of them cannot
    // hit anything, so its paths skip the scene trace (host/scene_setup.cpp: compute_cull_rects)
    int num_cull_rects;       // < 0: culling disabled
    float4 cull_rect[kMaxCullRects];
};

enum : int { kProfileV2 = 0, kProfileSimtTextured = 1, kProfileV4 = 2, kProfileV3Redo = 3, kProfileV3RedoS0 = 4 };
__host__ __device__ constexpr bool is_v3redo(int profile) { return profile == kProfileV3Redo || profile == kProfileV3RedoS0; }
enum : int { kEnvNone = 0, kEnvEquirect = 1, kEnvCubemap = 2 };
enum : int { kSamplerPoint = 0, kSamplerBilinear = 1, kSamplerRandom = 2 };
enum : int { kAccumAverage = 0, kAccumSum = 1 };

// threads per CTA of pt_render_kernel, per kernel family (tuning knobs; multiples of 32)
#ifndef B200PT_THREADS_CORNELL
#define B200PT_THREADS_CORNELL 256
#endif
#ifndef B200PT_THREADS_V4
#define B200PT_THREADS_V4 256
#endif
#ifndef B200PT_THREADS_V3REDO
#define B200PT_THREADS_V3REDO 256
#endif
__host__ __device__ constexpr int block_threads_for_profile(int profile)
{
    return profile == kProfileV4 ? B200PT_THREADS_V4 : (is_v3redo(profile) ? B200PT_THREADS_V3REDO : B200PT_THREADS_CORNELL);
}

struct LaunchConfig {
    int profile, env_kind, env_sampler, accum_mode;
    int static_scene;  // Cornell profiles: vertex coordinates as immediates (same bits, fewer instructions)
    int v4_flags;      // OPT_V4 + static_scene: the non-default shading switches compiled into the kernel (0 = the default kernels)
    int grid, block;
};

// implemented in pt_kernels_parity.cu / pt_kernels_fast.cu
struct SceneSet {
    CornellScene cornell;
    V4Scene v4;
    V3RedoScene v3redo;
    V3RedoScene0 v3redo0;
};
cudaError_t launch_render_parity(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t launch_render_fast(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t occupancy_parity(const LaunchConfig& lc, int* blocks_per_sm);
cudaError_t occupancy_fast(const LaunchConfig& lc, int* blocks_per_sm);
// pt_kernels_{parity,fast}_v4sw.cu: lc.v4_flags != 0
cudaError_t launch_render_parity_v4sw(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t launch_render_fast_v4sw(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t occupancy_parity_v4sw(const LaunchConfig& lc, int* blocks_per_sm);
cudaError_t occupancy_fast_v4sw(const LaunchConfig& lc, int* blocks_per_sm);
// the CTA-sorted variant (pt_wavefront.cuh), in pt_kernels_parity_sorted.cu / pt_kernels_fast_sorted.cu
cudaError_t launch_render_sorted_parity(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t launch_render_sorted_fast(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream);
cudaError_t occupancy_sorted_parity(const LaunchConfig& lc, int* blocks_per_sm);
cudaError_t occupancy_sorted_fast(const LaunchConfig& lc, int* blocks_per_sm);

// pt_post.cu
cudaError_t launch_resolve_ldr(const float* target, uint32_t* out, int width, int height, int tile_w, int tile_h,
                               int num_tiles_x, int mode, cudaStream_t stream);
cudaError_t launch_scale(float* target, size_t n, float scale, cudaStream_t stream);
cudaError_t launch_ffma_peak(float* scratch, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_build_item_order(const RenderParams& rp, int* order, cudaStream_t stream);
cudaError_t launch_tile_gather(const float* src, float* dst, int num_tiles, int mod, int rem, size_t floats_per_tile, int sm_count,
                               cudaStream_t stream);
cudaError_t launch_eval_portable(int op, const float* a, const float* b, float* out, size_t n, cudaStream_t stream);
cudaError_t launch_check_portable_tiers(int op, unsigned long long first, unsigned long long count, unsigned long long* counts,
                                        cudaStream_t stream);
cudaError_t launch_pack_env(const float* rgb, float4* rgba, size_t texels, cudaStream_t stream);

}  // namespace b200pt
