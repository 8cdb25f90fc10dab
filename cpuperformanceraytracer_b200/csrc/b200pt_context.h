// b200pt_context.h -- private state behind the opaque handles of include/b200pt.h, shared by
// b200pt_capi.cu (one context = one GPU) and b200pt_group.cu (several contexts of one process).
#pragma once
#include "../../include/b200pt.h"

#include <string>
#include <vector>

#include "../host/scene_setup.h"
#include "pt_common.cuh"

using namespace b200pt;

// present ring: two frames in flight + the one the caller still reads (its pointer stays valid until the next
// acquire), so three slots
constexpr int kRingSlots = 3;

struct b200pt_context {
    b200pt_params params{};
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // own_stream or a caller-provided one
    // two event pairs, used alternately: the launch path never waits for an earlier launch (it only polls)
    cudaEvent_t ev0[2] = {nullptr, nullptr}, ev1[2] = {nullptr, nullptr};
    bool timing_pending[2] = {false, false};
    unsigned timing_slot = 0;

    SceneSet scenes{};
    bool custom_scene = false;            // b200pt_set_scene_v4 installed a scene
    std::vector<float> scene_quads, scene_spheres;  // host copies for the culling rectangles
    float scene_cam[4] = {0.f, 0.f, 40.f, 1.f};
    float cameraDistance = 1.f;

    // target
    int width = 0, height = 0, ntx = 0, nty = 0, tile_w = 0, tile_h = 0;
    float* d_target_own = nullptr;
    float* d_target = nullptr;  // own or bound
    uint32_t* d_screen = nullptr;       // slot 0 of the present ring; also used by resolve_ldr / render_host
    uint32_t* d_ring[kRingSlots] = {nullptr, nullptr, nullptr};  // device frames of the present ring ([0] == d_screen)
    uint32_t* h_ring[kRingSlots] = {nullptr, nullptr, nullptr};  // pinned host frames of the present ring
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t render_done[kRingSlots] = {nullptr, nullptr, nullptr}, copy_done[kRingSlots] = {nullptr, nullptr, nullptr};
    int ring_frame[kRingSlots] = {0, 0, 0};
    unsigned long long submitted = 0, acquired = 0;
    uint32_t* d_rng = nullptr;
    int* d_work_counter = nullptr;
    // work items in the order they are pulled: the ones that trace the scene first, sky-only items last, so that
    // the end of a launch is filled with cheap items (rebuilt when the geometry / culling of a launch changes)
    int* d_item_order = nullptr;
    size_t item_order_capacity = 0;
    unsigned long long item_order_key = 0;        // geometry the table on the device was built for (0 = none in use)
    DeviceCounters* d_counters = nullptr;
    float* h_pinned = nullptr;  // staging for render_host (pinned, W*H*3 floats)
    uint32_t* h_pinned_screen = nullptr;
    size_t pinned_floats = 0;

    // env
    float* d_env_rgb = nullptr;
    float4* d_env_rgba = nullptr;
    cudaTextureObject_t env_tex = 0;
    int env_w = 0, env_h = 0;
    const float* last_env_ptr = nullptr;

    int iframe = 0;
    int first_tile = 0, num_tiles = 0;  // flat tile range rendered by this context (0, 0 = all tiles)
    // set by b200pt_group (B200PT_COMBINE_FUSED) around a launch: see RenderParams::scatter_stage
    float* scatter_stage[kMaxScatterRanks] = {};
    int scatter_gpo = 0;
    int tile_mod = 0, tile_rem = 0;     // or: every tile_mod-th tile (FlatTileIndex % tile_mod == tile_rem); 0 = off
    int blocks_per_sm = 0;
    bool sorted = false;           // B200PT_SCHED_SORTED: pt_render_sorted_kernel instead of pt_render_kernel
    int blocks_per_sm_sorted = 0;
    uint64_t paths = 0, launches = 0;
    double last_render_ms = 0.0;
    std::string last_error;
};

// Every entry point works on the context's device but leaves the caller's current device untouched
// (a host application -- or torch in the multi-GPU driver -- owns that setting).
struct DeviceGuard {
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        status = (prev == device) ? cudaSuccess : cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

inline int fail(b200pt_context* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->last_error = msg;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? B200PT_ERR_OUT_OF_MEMORY : B200PT_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
        }                                                                                           \
    } while (0)

