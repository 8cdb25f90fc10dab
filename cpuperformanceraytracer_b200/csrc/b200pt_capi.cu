// b200pt_capi.cu -- implementation of the C ABI declared in include/b200pt.h.
// Owns the device state the reference keeps in file-scope statics (frame counter, scene, tile
// table: demofox_path_tracing_optimization_v4.cpp:34,378,386,1343) plus the HBM copies of the
// caller's buffers.  No CPU rendering path exists here: without a usable GPU every call fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "b200pt_context.h"

namespace {

bool uses_env(const b200pt_params& p)
{
    if (p.profile == B200PT_PROFILE_SIMT_TEXTURED || p.profile == B200PT_PROFILE_V3_REDO || p.profile == B200PT_PROFILE_V3_REDO_SCENE0) return true;
    return p.profile == B200PT_PROFILE_OPT_V4 && p.env_kind != B200PT_ENV_NONE;
}

// Which scheduler a profile runs with unless the caller says otherwise: the faster one as measured on B200
// (DESIGN.md section 5).
int default_scheduler(int profile)
{
    (void)profile;
    return B200PT_SCHED_LANE;
}

LaunchConfig launch_config(const b200pt_context* c)
{
    LaunchConfig lc{};
    lc.profile = c->params.profile;
    lc.env_kind = (c->params.profile == B200PT_PROFILE_SIMT_TEXTURED || is_v3redo(c->params.profile)) ? kEnvEquirect
                 : c->params.profile == B200PT_PROFILE_V2           ? kEnvNone
                                                                    : c->params.env_kind;
    lc.env_sampler = is_v3redo(c->params.profile) ? kSamplerBilinear
                     : (c->params.profile == B200PT_PROFILE_OPT_V4 && c->params.env_kind != B200PT_ENV_NONE) ? c->params.env_sampler
                                                                                                              : kSamplerPoint;
    lc.accum_mode = c->params.accum_mode;
    lc.static_scene = (c->params.generic_scene_tables || c->custom_scene) ? 0 : 1;
    if (lc.profile == kProfileV4 && lc.static_scene && !v4_scene_matches_static_tables(c->scenes.v4)) lc.static_scene = 0;
    // the non-default shading switches: compiled into their own scene-specialised kernels (pt_kernels_*_v4sw.cu); the generic
    // kernels read them from RenderParams::v4_flags
    if (lc.profile == kProfileV4 && lc.static_scene) lc.v4_flags = (c->params.exact_exp ? 1 : 0) | (c->params.sincos_unit_vectors ? 2 : 0);
    if (lc.static_scene && lc.profile == kProfileV3Redo && !v3redo_spheres_match_static_tables(c->scenes.v3redo.sphere)) lc.static_scene = 0;
    if (lc.static_scene && lc.profile == kProfileV3RedoS0 && !v3redo0_spheres_match_static_tables(c->scenes.v3redo0.sphere)) lc.static_scene = 0;
    if (lc.static_scene && (lc.profile == kProfileV2 || lc.profile == kProfileSimtTextured) &&
        !cornell_spheres_match_static_tables(c->scenes.cornell.sphere))
        lc.static_scene = 0;
    lc.block = block_threads_for_profile(lc.profile);  // the sorted kernel always runs 256 threads (kWfThreads)
    lc.grid = 1;
    return lc;
}

void free_target(b200pt_context* c)
{
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);  // a present copy may still read a ring slot
    if (c->d_target_own) cudaFree(c->d_target_own);
    if (c->d_screen) cudaFree(c->d_screen);
    for (int i = 0; i < kRingSlots; i++) {
        if (i > 0 && c->d_ring[i]) cudaFree(c->d_ring[i]);  // d_ring[0] is d_screen
        c->d_ring[i] = nullptr;
        if (c->h_ring[i]) cudaFreeHost(c->h_ring[i]);
        c->h_ring[i] = nullptr;
    }
    c->submitted = c->acquired = 0;
    if (c->d_rng) cudaFree(c->d_rng);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_pinned_screen) cudaFreeHost(c->h_pinned_screen);
    c->d_target_own = c->d_target = nullptr;
    c->d_screen = nullptr;
    c->d_rng = nullptr;
    c->h_pinned = nullptr;
    c->h_pinned_screen = nullptr;
    c->pinned_floats = 0;
}

void free_env(b200pt_context* c)
{
    if (c->env_tex) cudaDestroyTextureObject(c->env_tex);
    if (c->d_env_rgb) cudaFree(c->d_env_rgb);
    if (c->d_env_rgba) cudaFree(c->d_env_rgba);
    c->env_tex = 0;
    c->d_env_rgb = nullptr;
    c->d_env_rgba = nullptr;
    c->env_w = c->env_h = 0;
    c->last_env_ptr = nullptr;
}

// Device time of render launches, CUDA events on the launching stream.  wait = false (the launch path): only
// launches that have already finished are harvested -- the host is never blocked behind an earlier kernel, so
// callers can queue ahead.  wait = true (synchronize / download / get_counters): everything pending is
// collected; the newest launch wins.
int collect_timing(b200pt_context* c, bool wait)
{
    for (unsigned k = 0; k < 2; k++) {
        const unsigned slot = (c->timing_slot + k) & 1u;  // older pair first
        if (!c->timing_pending[slot]) continue;
        if (!wait) {
            const cudaError_t q = cudaEventQuery(c->ev1[slot]);
            if (q == cudaErrorNotReady) continue;
            CUDA_TRY(c, q);
        } else {
            CUDA_TRY(c, cudaEventSynchronize(c->ev1[slot]));
        }
        float ms = 0.f;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0[slot], c->ev1[slot]));
        c->last_render_ms = ms;
        c->timing_pending[slot] = false;
    }
    return B200PT_OK;
}

// The pull order of a launch's work items (RenderParams::item_order).  Items are pulled from one atomic counter; an
// item that looks into the scene costs ~15x a sky-only one (camera-culled pixels never trace), so pulling the
// expensive ones first leaves the cheap ones to even out the end of the launch -- what the reference gets from its
// fine-grained per-tile queue (work_queue.cpp:7-66).  The table is built on the device (build_item_order_kernel,
// pt_post.cu: a few microseconds, stream-ordered, no host synchronisation) whenever the launch geometry changes.
int ensure_item_order(b200pt_context* c, const RenderParams& rp)
{
    const bool strided = rp.tile_mod > 1;  // the table then also says WHICH items the launch renders: mandatory
    if (!strided && (rp.num_cull_rects <= 0 || c->params.disable_item_order || rp.num_items < 2 * c->sm_count * 8)) {
        c->item_order_key = 0;
        return B200PT_OK;  // nothing to gain: launch without an order table
    }
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
    mix((unsigned)rp.width); mix((unsigned)rp.height); mix((unsigned)rp.tile_w); mix((unsigned)rp.tile_h); mix((unsigned)rp.num_tiles_x);
    mix((unsigned)rp.group_offset); mix((unsigned)rp.num_groups); mix((unsigned)rp.block_items); mix((unsigned)(rp.num_cull_rects + 1));
    mix((unsigned)rp.tile_mod); mix((unsigned)rp.tile_rem); mix((unsigned)rp.num_items);
    for (int k = 0; k < rp.num_cull_rects; k++) {
        unsigned u[4];
        std::memcpy(u, &rp.cull_rect[k], sizeof(u));
        for (unsigned w : u) mix(w);
    }
    if (key == 0) key = 1;
    if (key == c->item_order_key && c->d_item_order) return B200PT_OK;
    if (rp.nframes < 8 && !strided) {  // a new geometry for a few frames only (bands of a pipelined present): not worth two extra launches
        c->item_order_key = 0;
        return B200PT_OK;
    }
    const size_t need = (size_t)rp.num_items + ((size_t)rp.order_domain_items + 1023) / 1024;  // the table + per-block scratch
    if (c->item_order_capacity < need) {
        // launches that read the old table are ordered before the free on this stream
        if (c->d_item_order) CUDA_TRY(c, cudaFreeAsync(c->d_item_order, c->stream));
        c->d_item_order = nullptr;
        c->item_order_capacity = 0;
        CUDA_TRY(c, cudaMallocAsync(&c->d_item_order, need * sizeof(int), c->stream));
        c->item_order_capacity = need;
    }
    CUDA_TRY(c, launch_build_item_order(rp, c->d_item_order, c->stream));
    c->launches += 2;
    c->item_order_key = key;
    return B200PT_OK;
}

}  // namespace

extern "C" {

int b200pt_api_version(void) { return B200PT_API_VERSION; }

const char* b200pt_error_string(int code)
{
    switch (code) {
    case B200PT_OK: return "ok";
    case B200PT_ERR_INVALID_ARGUMENT: return "invalid argument";
    case B200PT_ERR_CUDA: return "CUDA error / no usable device";
    case B200PT_ERR_NOT_READY: return "not ready (resize / set_env first)";
    case B200PT_ERR_OUT_OF_MEMORY: return "out of device memory";
    default: return "unknown error";
    }
}

const char* b200pt_last_error(b200pt_context* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int b200pt_default_params(int profile, b200pt_params* p)
{
    if (!p || profile < B200PT_PROFILE_V2 || profile > B200PT_PROFILE_V3_REDO_SCENE0) return B200PT_ERR_INVALID_ARGUMENT;
    std::memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(b200pt_params);
    p->device = 0;
    p->profile = profile;
    p->math_mode = B200PT_MATH_PARITY;
    p->num_bounces = -1;
    p->accum_mode = B200PT_ACCUM_RUNNING_AVERAGE;
    p->output_to_screen = 0;
    if (profile == B200PT_PROFILE_OPT_V4) {
        p->env_kind = B200PT_ENV_EQUIRECT;       // USE_ENV_MAP 1, USE_ENV_CUBEMAP 0
        p->env_sampler = B200PT_SAMPLER_RANDOM;  // USE_RANDOM_JITTER_TEXTURE_SAMPLING 1
    } else if (profile == B200PT_PROFILE_SIMT_TEXTURED) {
        p->env_kind = B200PT_ENV_EQUIRECT;
        p->env_sampler = B200PT_SAMPLER_POINT;
    } else if (profile == B200PT_PROFILE_V3_REDO || profile == B200PT_PROFILE_V3_REDO_SCENE0) {
        p->env_kind = B200PT_ENV_EQUIRECT;  // EquirectangularTextureSampleBilinear, v3_redo.cpp:638
        p->env_sampler = B200PT_SAMPLER_BILINEAR;
    }
    return B200PT_OK;
}

int b200pt_create(const b200pt_params* params, b200pt_context** out_ctx)
{
    if (!params || !out_ctx) return B200PT_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    if (params->struct_size != (int32_t)sizeof(b200pt_params)) return B200PT_ERR_INVALID_ARGUMENT;
    if (params->profile < B200PT_PROFILE_V2 || params->profile > B200PT_PROFILE_V3_REDO_SCENE0) return B200PT_ERR_INVALID_ARGUMENT;
    if (params->math_mode != B200PT_MATH_PARITY && params->math_mode != B200PT_MATH_FAST) return B200PT_ERR_INVALID_ARGUMENT;
    if (params->accum_mode != B200PT_ACCUM_RUNNING_AVERAGE && params->accum_mode != B200PT_ACCUM_SUM) return B200PT_ERR_INVALID_ARGUMENT;
    if (params->exact_tonemap & ~(B200PT_TONEMAP_EXACT_ACES | B200PT_TONEMAP_EXACT_GAMMA)) return B200PT_ERR_INVALID_ARGUMENT;
    // the reference reads USE_FAST_APPROXIMATE_EXP / USE_UNIT_VECTOR_REJECTION_SAMPLING in the v4 source only
    if (params->profile != B200PT_PROFILE_OPT_V4 && (params->exact_exp || params->sincos_unit_vectors)) return B200PT_ERR_INVALID_ARGUMENT;
    if (params->profile == B200PT_PROFILE_OPT_V4) {
        if (params->env_kind < B200PT_ENV_NONE || params->env_kind > B200PT_ENV_CUBEMAP) return B200PT_ERR_INVALID_ARGUMENT;
        if (params->env_kind != B200PT_ENV_NONE && params->env_sampler != B200PT_SAMPLER_BILINEAR &&
            params->env_sampler != B200PT_SAMPLER_RANDOM)
            return B200PT_ERR_INVALID_ARGUMENT;
    }
    b200pt_context* c = new (std::nothrow) b200pt_context();
    if (!c) return B200PT_ERR_OUT_OF_MEMORY;
    c->params = *params;
    if (c->params.num_bounces < 0) c->params.num_bounces = (params->profile == B200PT_PROFILE_OPT_V4 || is_v3redo(params->profile)) ? 8 : 4;
    c->device = params->device;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || c->device < 0 || c->device >= ndev) {
        delete c;
        return B200PT_ERR_CUDA;  // no CPU fallback
    }
    DeviceGuard guard(c->device);
    cudaDeviceProp prop{};
    if (guard.status != cudaSuccess || cudaGetDeviceProperties(&prop, c->device) != cudaSuccess ||
        prop.major < 10) {
        delete c;
        return B200PT_ERR_CUDA;  // kernels are sm_100a only
    }
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->render_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->render_done[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->render_done[2], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_done[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_done[2], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->ev0[0]) != cudaSuccess || cudaEventCreate(&c->ev1[0]) != cudaSuccess ||
        cudaEventCreate(&c->ev0[1]) != cudaSuccess || cudaEventCreate(&c->ev1[1]) != cudaSuccess ||
        cudaMalloc(&c->d_work_counter, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&c->d_counters, sizeof(DeviceCounters)) != cudaSuccess ||
        cudaMemset(c->d_counters, 0, sizeof(DeviceCounters)) != cudaSuccess) {
        b200pt_destroy(c);
        return B200PT_ERR_CUDA;
    }
    c->stream = c->own_stream;

    // InitializeCamera / InitializeScene (v4.cpp:1403-1502) and the Cornell vertex tables
    c->cameraDistance = camera_distance();
    build_cornell_scene(&c->scenes.cornell, params->profile == B200PT_PROFILE_SIMT_TEXTURED);
    build_v4_scene(&c->scenes.v4);
    build_v3redo_scene(&c->scenes.v3redo);
    build_v3redo_scene0(&c->scenes.v3redo0);

    LaunchConfig lc = launch_config(c);
    int bps = 0;
    if (lc.v4_flags) e = (c->params.math_mode == B200PT_MATH_PARITY) ? occupancy_parity_v4sw(lc, &bps) : occupancy_fast_v4sw(lc, &bps);
    else e = (c->params.math_mode == B200PT_MATH_PARITY) ? occupancy_parity(lc, &bps) : occupancy_fast(lc, &bps);
    if (e != cudaSuccess || bps <= 0) {
        b200pt_destroy(c);
        return B200PT_ERR_CUDA;
    }
    c->blocks_per_sm = bps;
    // the CTA-sorted scheduler: asked for by the caller, by the environment (A/B measurements), or the profile's default
    int sched = c->params.scheduler;
    if (const char* env = std::getenv("B200PT_SCHEDULER")) {
        if (!std::strcmp(env, "lane")) sched = B200PT_SCHED_LANE;
        else if (!std::strcmp(env, "sorted")) sched = B200PT_SCHED_SORTED;
    }
    if (sched == B200PT_SCHED_DEFAULT) sched = default_scheduler(c->params.profile);
    if (sched != B200PT_SCHED_LANE && sched != B200PT_SCHED_SORTED) {
        b200pt_destroy(c);
        return B200PT_ERR_INVALID_ARGUMENT;
    }
    c->sorted = sched == B200PT_SCHED_SORTED;
    if (c->sorted) {
        bps = 0;
        if (lc.v4_flags) lc.static_scene = lc.v4_flags = 0;  // the sorted kernels take the switches at run time (generic kernel)
        e = (c->params.math_mode == B200PT_MATH_PARITY) ? occupancy_sorted_parity(lc, &bps) : occupancy_sorted_fast(lc, &bps);
        if (e != cudaSuccess || bps <= 0) {
            b200pt_destroy(c);
            return B200PT_ERR_CUDA;
        }
        c->blocks_per_sm_sorted = bps;
    }
    *out_ctx = c;
    return B200PT_OK;
}

int b200pt_destroy(b200pt_context* c)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    free_target(c);
    free_env(c);
    if (c->d_work_counter) cudaFree(c->d_work_counter);
    if (c->d_item_order) cudaFreeAsync(c->d_item_order, c->own_stream);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int i = 0; i < kRingSlots; i++) {
        if (c->render_done[i]) cudaEventDestroy(c->render_done[i]);
        if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (c->ev0[i]) cudaEventDestroy(c->ev0[i]);
        if (c->ev1[i]) cudaEventDestroy(c->ev1[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return B200PT_OK;
}

int b200pt_set_env(b200pt_context* c, b200pt_texture tex)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!tex.Data || tex.Width <= 0 || tex.Height <= 0 || tex.Components != 3)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "env texture must be RGB f32 with positive size");
    // the reference indexes texels through binary32 arithmetic (texture.cpp:56-65,84): exact up to 2^24 floats
    if ((long long)tex.Width * tex.Height * 3 >= (1LL << 24))
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "env texture too large for the reference's float texel indexing");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const size_t texels = (size_t)tex.Width * tex.Height;
    if (texels != (size_t)c->env_w * c->env_h) {
        free_env(c);
        CUDA_TRY(c, cudaMalloc(&c->d_env_rgb, texels * 3 * sizeof(float)));
        CUDA_TRY(c, cudaMalloc(&c->d_env_rgba, texels * sizeof(float4)));
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = c->d_env_rgba;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = texels * sizeof(float4);
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        td.filterMode = cudaFilterModePoint;
        td.addressMode[0] = cudaAddressModeBorder;
        td.normalizedCoords = 0;
        CUDA_TRY(c, cudaCreateTextureObject(&c->env_tex, &rd, &td, nullptr));
    }
    c->env_w = tex.Width;
    c->env_h = tex.Height;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_env_rgb, tex.Data, texels * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, launch_pack_env(c->d_env_rgb, c->d_env_rgba, texels, c->stream));
    c->launches++;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->last_env_ptr = tex.Data;
    return B200PT_OK;
}

int b200pt_set_scene_v4(b200pt_context* c, const b200pt_quad* quads, int32_t num_quads, const b200pt_sphere* spheres,
                        int32_t num_spheres, const b200pt_material* materials, const b200pt_camera* camera)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (c->params.profile != B200PT_PROFILE_OPT_V4) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "runtime scenes exist for the OPT_V4 profile only");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (num_quads == 0 && num_spheres == 0) {  // back to InitializeScene's scene
        build_v4_scene(&c->scenes.v4);
        c->custom_scene = false;
        return B200PT_OK;
    }
    if (!materials || !camera || !(camera->Distance > 0.f))
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "materials and a camera with Distance > 0 are required");
    static_assert(sizeof(b200pt_quad) == 12 * sizeof(float) && sizeof(b200pt_sphere) == 4 * sizeof(float) &&
                  sizeof(b200pt_material) == 17 * sizeof(float), "plain float records");
    V4Scene s;
    if (!build_v4_scene_from(&s, reinterpret_cast<const float*>(quads), num_quads, reinterpret_cast<const float*>(spheres),
                             num_spheres, reinterpret_cast<const float*>(materials), camera->Position, camera->Distance))
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "need 1..12 objects (MAX_OBJECTS / MAX_MATERIALS)");
    c->scenes.v4 = s;
    c->custom_scene = true;
    c->scene_quads.assign(reinterpret_cast<const float*>(quads), reinterpret_cast<const float*>(quads) + 12 * (size_t)num_quads);
    c->scene_spheres.assign(reinterpret_cast<const float*>(spheres), reinterpret_cast<const float*>(spheres) + 4 * (size_t)num_spheres);
    c->scene_cam[0] = camera->Position[0];
    c->scene_cam[1] = camera->Position[1];
    c->scene_cam[2] = camera->Position[2];
    c->scene_cam[3] = camera->Distance;
    return B200PT_OK;
}

int b200pt_set_scene_cornell(b200pt_context* c, const b200pt_quad* quads, const b200pt_sphere* spheres, const b200pt_material_legacy* materials)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (c->params.profile != B200PT_PROFILE_V2 && c->params.profile != B200PT_PROFILE_SIMT_TEXTURED)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "b200pt_set_scene_cornell is for the V2 / SIMT_TEXTURED profiles");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!quads) {  // back to the literals of v2.cpp:320-454
        build_cornell_scene(&c->scenes.cornell, c->params.profile == B200PT_PROFILE_SIMT_TEXTURED);
        c->custom_scene = false;
        return B200PT_OK;
    }
    if (!spheres || !materials) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "6 quads, 3 spheres and 9 materials are required");
    static_assert(sizeof(b200pt_material_legacy) == 11 * sizeof(float), "plain float record");
    CornellScene s;
    if (!build_cornell_scene_from(&s, reinterpret_cast<const float*>(quads), reinterpret_cast<const float*>(spheres),
                                  reinterpret_cast<const float*>(materials)))
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "coordinates must lie within +-1e6, radii in [1e-3, 1e6]");
    c->scenes.cornell = s;
    c->custom_scene = true;
    c->scene_quads.assign(reinterpret_cast<const float*>(quads), reinterpret_cast<const float*>(quads) + 12 * kCornellQuads);
    c->scene_spheres.assign(reinterpret_cast<const float*>(spheres), reinterpret_cast<const float*>(spheres) + 4 * kCornellSpheres);
    return B200PT_OK;
}

int b200pt_compute_cull_rects_scene_cornell(const b200pt_quad* quads, const b200pt_sphere* spheres, int32_t width, int32_t height,
                                            float* rects, int32_t* count)
{
    if (!quads || !spheres || !rects || !count || width <= 0 || height <= 0) return B200PT_ERR_INVALID_ARGUMENT;
    float4 r[kMaxCullRects];
    const int n = compute_cull_rects_cornell(reinterpret_cast<const float*>(quads), reinterpret_cast<const float*>(spheres), width, height, r);
    *count = n;
    for (int i = 0; i < n; i++) {
        rects[4 * i + 0] = r[i].x; rects[4 * i + 1] = r[i].y; rects[4 * i + 2] = r[i].z; rects[4 * i + 3] = r[i].w;
    }
    return B200PT_OK;
}

int b200pt_resize(b200pt_context* c, int32_t width, int32_t height, int32_t ntx, int32_t nty)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    // CheckValidSettings, Application.cpp:36-94
    if (width <= 0 || height <= 0 || ntx <= 0 || nty <= 0 || width % ntx || height % nty || (width / ntx) % 8)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "invalid tiling: need W % ntx == 0, H % nty == 0, tile width % 8 == 0");
    if ((long long)width * height * 3 >= (1LL << 31)) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "image too large");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const size_t nfloats = (size_t)width * height * 3;
    const bool bound = c->d_target && c->d_target != c->d_target_own;
    const bool realloc = (size_t)c->width * c->height != (size_t)width * height || !c->d_target_own;
    // a caller-bound device target (b200pt_bind_device_target) was sized for the old image: refuse instead of
    // silently rendering somewhere else than the buffer the caller keeps reducing
    if (bound && realloc)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "a device target of another size is bound: b200pt_bind_device_target(ctx, NULL) first");
    if (realloc) {
        free_target(c);
        c->width = c->height = 0;  // a failed allocation leaves a consistent "not ready" context
        cudaError_t e = cudaMalloc(&c->d_target_own, nfloats * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_screen, (size_t)width * height * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_rng, (size_t)width * height * sizeof(uint32_t));
        if (e != cudaSuccess) {
            free_target(c);
            CUDA_TRY(c, e);
        }
        c->d_target = c->d_target_own;
    } else if (!bound) {
        c->d_target = c->d_target_own;
    }
    c->width = width;
    c->height = height;
    c->ntx = ntx;
    c->nty = nty;
    c->tile_w = width / ntx;
    c->tile_h = height / nty;
    c->first_tile = c->num_tiles = 0;
    c->tile_mod = c->tile_rem = 0;
    // frames still in the present ring belong to the old image
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    c->submitted = c->acquired = 0;
    return b200pt_reset(c);
}

int b200pt_reset(b200pt_context* c)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    const size_t nfloats = (size_t)c->width * c->height * 3;
    CUDA_TRY(c, cudaMemsetAsync(c->d_target, 0, nfloats * sizeof(float), c->stream));  // Application.cpp:151
    CUDA_TRY(c, cudaMemsetAsync(c->d_screen, 0, (size_t)c->width * c->height * 4, c->stream));
    c->iframe = 0;  // counters stay cumulative since create
    return B200PT_OK;
}

int b200pt_set_frame_counter(b200pt_context* c, int32_t iframe)
{
    if (!c || iframe < 0) return B200PT_ERR_INVALID_ARGUMENT;
    c->iframe = iframe;
    return B200PT_OK;
}

int b200pt_get_frame_counter(b200pt_context* c, int32_t* iframe)
{
    if (!c || !iframe) return B200PT_ERR_INVALID_ARGUMENT;
    *iframe = c->iframe;
    return B200PT_OK;
}

static int render_frames_impl(b200pt_context* c, int32_t nframes, uint32_t* screen);

int b200pt_render_frames(b200pt_context* c, int32_t nframes)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    // OUTPUT_TO_SCREEN: tone-map into the screen buffer as part of the render (v4.cpp:1562-1564)
    return render_frames_impl(c, nframes, c->params.output_to_screen ? c->d_screen : nullptr);
}

static int render_frames_impl(b200pt_context* c, int32_t nframes, uint32_t* screen)
{
    if (!c || nframes < 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (uses_env(c->params) && !c->env_tex) return fail(c, B200PT_ERR_NOT_READY, "this profile samples an env map: set_env first");
    if (nframes == 0) return B200PT_OK;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    if (collect_timing(c, false) != B200PT_OK) return B200PT_ERR_CUDA;  // poll only: never blocks the launch path

    RenderParams rp{};
    rp.target = c->d_target;
    rp.rng_out = c->d_rng;
    // the pow() gamma is binary64 work that stays out of the render kernels: they skip their fused tone map and the resolve
    // kernel follows on the same stream (every other tone map is fused; B200PT_LDR_EXACT_* = B200PT_TONEMAP_EXACT_* << 1)
    const bool resolve_after = screen && (c->params.exact_tonemap & B200PT_TONEMAP_EXACT_GAMMA);
    rp.screen = resolve_after ? nullptr : screen;
    rp.screen_mode = B200PT_LDR_SCREEN_BGRA | ((c->params.exact_tonemap & B200PT_TONEMAP_EXACT_ACES) << 1);
    rp.v4_flags = (c->params.exact_exp ? 1 : 0) | (c->params.sincos_unit_vectors ? 2 : 0);
    rp.work_counter = c->d_work_counter;
    rp.counters = c->d_counters;
    rp.env = c->env_tex;
    rp.env_w = c->env_w;
    rp.env_h = c->env_h;
    rp.width = c->width;
    rp.height = c->height;
    rp.tile_w = c->tile_w;
    rp.tile_h = c->tile_h;
    rp.num_tiles_x = c->ntx;
    rp.groups_per_tile_row = c->tile_w / 8;
    rp.groups_per_tile = rp.groups_per_tile_row * c->tile_h;
    // tiles follow one another in FlatTileIndex order in the buffer (RenderTile, v4.cpp:1189-1194)
    const int ntiles = c->num_tiles > 0 ? c->num_tiles : c->ntx * c->nty;
    rp.group_offset = (c->num_tiles > 0 ? c->first_tile : 0) * rp.groups_per_tile;
    rp.num_groups = ntiles * rp.groups_per_tile;
    rp.num_items = (rp.num_groups + 3) / 4;
    rp.block_items = (c->tile_h % 4 == 0) ? 1 : 0;
    rp.order_domain_items = rp.num_items;
    long long launch_groups = rp.num_groups;
    if (c->tile_mod > 1) {
        // every tile_mod-th tile of the whole image: the kernel walks the image's item space through the order table
        if (rp.groups_per_tile % 4) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "a tile stride needs tiles of a multiple of 32 pixels");
        const int all_tiles = c->ntx * c->nty;
        const int mine = c->tile_rem < all_tiles ? (all_tiles - c->tile_rem + c->tile_mod - 1) / c->tile_mod : 0;
        rp.tile_mod = c->tile_mod;
        rp.tile_rem = c->tile_rem;
        rp.order_domain_items = rp.num_items;
        rp.num_items = mine * (rp.groups_per_tile / 4);
        launch_groups = (long long)mine * rp.groups_per_tile;
        if (mine == 0) {  // nothing to render on this context: the frame counter still advances
            c->iframe += nframes;
            return B200PT_OK;
        }
    }
    rp.first_frame = c->iframe + 1;  // iFrame += 1 before the render, v4.cpp:1703
    rp.nframes = nframes;
    rp.num_bounces = c->params.num_bounces;
    rp.cameraDistance = c->cameraDistance;
    rp.rcp_width = 1.0f / (float)c->width;
    rp.rcp_height = 1.0f / (float)c->height;
    rp.aspect = (float)c->width / (float)c->height;
    {
        // significant bits of an integer = bit length minus trailing zeros
        auto sig_bits = [](unsigned v) { int len = 0, tz = 0; for (unsigned t = v; t; t >>= 1) len++; while (v && !(v & 1u)) { v >>= 1; tz++; } return len - tz; };
        rp.res_div_exact = (sig_bits((unsigned)c->width) <= 16 && sig_bits((unsigned)c->height) <= 16) ? 1 : 0;
    }
    if (c->params.disable_camera_culling) rp.num_cull_rects = -1;
    else if (c->custom_scene && c->params.profile != B200PT_PROFILE_OPT_V4)
        rp.num_cull_rects = compute_cull_rects_cornell(c->scene_quads.data(), c->scene_spheres.data(), c->width, c->height, rp.cull_rect);
    else if (c->custom_scene)
        rp.num_cull_rects = compute_cull_rects_v4(c->scene_quads.data(), c->scenes.v4.numQuads, c->scene_spheres.data(),
                                                  c->scenes.v4.numSpheres, c->scene_cam, c->scene_cam[3], c->width, c->height, rp.cull_rect);
    else rp.num_cull_rects = compute_cull_rects(c->params.profile, c->width, c->height, rp.cull_rect);

    if (c->scatter_gpo > 0 && c->params.accum_mode == B200PT_ACCUM_SUM && c->num_tiles == 0 && c->tile_mod <= 1) {
        rp.scatter_gpo = c->scatter_gpo;
        for (int i = 0; i < kMaxScatterRanks; i++) rp.scatter_stage[i] = c->scatter_stage[i];
    }
    {
        const int rc = ensure_item_order(c, rp);
        if (rc != B200PT_OK) return rc;
        rp.item_order = c->item_order_key ? c->d_item_order : nullptr;
    }

    LaunchConfig lc = launch_config(c);
    // persistent grid: every SM holds blocks_per_sm resident CTAs; warps pull 32-pixel items
    // the sorted kernel packs the bounce count into 8 bits and the pixel coordinates into 16 each
    const bool sorted = c->sorted && rp.num_bounces <= 250 && rp.width <= 65535 && rp.height <= 65535;
    if (sorted && lc.v4_flags) lc.static_scene = lc.v4_flags = 0;  // the sorted kernels take the switches at run time
    const int warps_per_block = (sorted ? 256 : lc.block) / 32;
    const int max_useful_blocks = (rp.num_items + warps_per_block - 1) / warps_per_block;
    lc.grid = c->sm_count * (sorted ? c->blocks_per_sm_sorted : c->blocks_per_sm);
    if (lc.grid > max_useful_blocks) lc.grid = max_useful_blocks;
    if (lc.grid < 1) lc.grid = 1;

    CUDA_TRY(c, cudaMemsetAsync(c->d_work_counter, 0, sizeof(int), c->stream));
    const unsigned ts = c->timing_slot;
    c->timing_slot ^= 1u;
    CUDA_TRY(c, cudaEventRecord(c->ev0[ts], c->stream));
    cudaError_t e;
    if (sorted)
        e = (c->params.math_mode == B200PT_MATH_PARITY) ? launch_render_sorted_parity(lc, rp, c->scenes, c->stream)
                                                        : launch_render_sorted_fast(lc, rp, c->scenes, c->stream);
    else if (lc.v4_flags)
        e = (c->params.math_mode == B200PT_MATH_PARITY) ? launch_render_parity_v4sw(lc, rp, c->scenes, c->stream)
                                                        : launch_render_fast_v4sw(lc, rp, c->scenes, c->stream);
    else
        e = (c->params.math_mode == B200PT_MATH_PARITY) ? launch_render_parity(lc, rp, c->scenes, c->stream)
                                                        : launch_render_fast(lc, rp, c->scenes, c->stream);
    CUDA_TRY(c, e);
    if (resolve_after) {
        CUDA_TRY(c, launch_resolve_ldr(c->d_target, screen, c->width, c->height, c->tile_w, c->tile_h, c->ntx,
                                       B200PT_LDR_SCREEN_BGRA | (c->params.exact_tonemap << 1), c->stream));
        c->launches++;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev1[ts], c->stream));
    c->timing_pending[ts] = true;
    c->launches++;
    c->iframe += nframes;
    c->paths += (uint64_t)launch_groups * 8u * (uint64_t)nframes;

    return B200PT_OK;
}

int b200pt_synchronize(b200pt_context* c)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return collect_timing(c, true);
}

int b200pt_upload_target(b200pt_context* c, const float* host_src)
{
    if (!c || !host_src) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaMemcpyAsync(c->d_target, host_src, (size_t)c->width * c->height * 3 * sizeof(float),
                                cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

int b200pt_download_target(b200pt_context* c, float* host_dst)
{
    if (!c || !host_dst) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaMemcpyAsync(host_dst, c->d_target, (size_t)c->width * c->height * 3 * sizeof(float),
                                cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return collect_timing(c, true);
}

int b200pt_render_host(b200pt_context* c, float* BufferOut, int32_t W, int32_t H, int32_t NumTilesX, int32_t NumTilesY,
                       int32_t TileWidth, int32_t TileHeight, int32_t NumChannels, b200pt_texture Texture,
                       void* ScreenBufferData, int32_t nframes)
{
    if (!c || !BufferOut || nframes < 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (NumChannels != 3) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "NumChannels must be 3");
    if (NumTilesX <= 0 || NumTilesY <= 0 || TileWidth * NumTilesX != W || TileHeight * NumTilesY != H)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "tiles must cover the buffer exactly");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    if (c->width != W || c->height != H || c->ntx != NumTilesX || c->nty != NumTilesY) {
        const int keep_frame = c->iframe;
        int rc = b200pt_resize(c, W, H, NumTilesX, NumTilesY);
        if (rc != B200PT_OK) return rc;
        c->iframe = keep_frame;  // the reference's static iFrame survives a resize
    }
    if (uses_env(c->params)) {
        if (Texture.Data && (Texture.Data != c->last_env_ptr || Texture.Width != c->env_w || Texture.Height != c->env_h)) {
            int rc = b200pt_set_env(c, Texture);
            if (rc != B200PT_OK) return rc;
        }
    }
    const size_t nfloats = (size_t)W * H * 3;
    // A caller buffer that is already page-locked (cudaHostAlloc / cudaHostRegister) is copied directly; pageable
    // memory (the reference's _aligned_malloc) goes through the context's pinned staging buffers.
    auto is_pinned = [](const void* ptr) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    };
    const bool direct = is_pinned(BufferOut);
    const bool want_screen = ScreenBufferData && c->params.output_to_screen;
    const bool direct_screen = want_screen && is_pinned(ScreenBufferData);
    if ((!direct || (want_screen && !direct_screen)) && c->pinned_floats != nfloats) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        if (c->h_pinned_screen) cudaFreeHost(c->h_pinned_screen);
        c->h_pinned = nullptr;
        c->h_pinned_screen = nullptr;
        CUDA_TRY(c, cudaMallocHost(&c->h_pinned, nfloats * sizeof(float)));
        CUDA_TRY(c, cudaMallocHost(&c->h_pinned_screen, (size_t)W * H * sizeof(uint32_t)));
        c->pinned_floats = nfloats;
    }
    // host accumulation state -> (pinned staging ->) HBM; render; HBM -> (pinned ->) host
    float* src = BufferOut;
    if (!direct) {
        std::memcpy(c->h_pinned, BufferOut, nfloats * sizeof(float));
        src = c->h_pinned;
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->d_target, src, nfloats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    int rc = b200pt_render_frames(c, nframes);
    if (rc != B200PT_OK) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(src, c->d_target, nfloats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (want_screen)
        CUDA_TRY(c, cudaMemcpyAsync(direct_screen ? ScreenBufferData : (void*)c->h_pinned_screen, c->d_screen,
                                    (size_t)W * H * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!direct) std::memcpy(BufferOut, c->h_pinned, nfloats * sizeof(float));
    if (want_screen && !direct_screen) std::memcpy(ScreenBufferData, c->h_pinned_screen, (size_t)W * H * sizeof(uint32_t));
    return collect_timing(c, true);
}

int b200pt_resolve_ldr(b200pt_context* c, uint32_t* host_dst, int32_t mode, int32_t bump_frame_counter)
{
    if (!c || !host_dst || (mode & ~(B200PT_LDR_SCREEN_BGRA | B200PT_LDR_EXACT_ACES | B200PT_LDR_EXACT_GAMMA))) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    mode |= c->params.exact_tonemap << 1;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));  // the present ring may still be reading slot 0
    CUDA_TRY(c, launch_resolve_ldr(c->d_target, c->d_screen, c->width, c->height, c->tile_w, c->tile_h, c->ntx, mode, c->stream));
    c->launches++;
    CUDA_TRY(c, cudaMemcpyAsync(host_dst, c->d_screen, (size_t)c->width * c->height * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (bump_frame_counter) c->iframe += 1;  // CopyOutputToFile: iFrame += 1.0f, v4.cpp:1741
    return B200PT_OK;
}

// ---- progressive present path: the windowed loop of ApplicationState::RunApp
// (Application.cpp:306-375) renders NUM_SAMPLES_PER_FRAME, tone-maps per tile (OutputToScreen) and
// presents.  Here: render + fused tone map into one of two device frames, asynchronous copy to one of
// two pinned host frames on a second stream, so the copy of frame k overlaps the render of frame k+1.
int b200pt_present_submit(b200pt_context* c, int32_t nframes)
{
    if (!c || nframes <= 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (c->submitted - c->acquired >= 2) return fail(c, B200PT_ERR_NOT_READY, "present ring full: acquire a frame first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    const size_t bytes = (size_t)c->width * c->height * sizeof(uint32_t);
    c->d_ring[0] = c->d_screen;
    for (int i = 0; i < kRingSlots; i++) {
        if (!c->d_ring[i]) CUDA_TRY(c, cudaMalloc(&c->d_ring[i], bytes));
        if (!c->h_ring[i]) CUDA_TRY(c, cudaMallocHost(&c->h_ring[i], bytes));
    }
    // Slot k % 3.  Frame k's slot is written again by submit k + 3, which needs acquired >= k + 2, i.e. the caller
    // has already asked for frame k + 1: the pointer handed out for frame k is dead by then (see b200pt.h).
    const int slot = (int)(c->submitted % kRingSlots);
    uint32_t* dscreen = c->d_ring[slot];
    const int rc = render_frames_impl(c, nframes, dscreen);
    if (rc != B200PT_OK) return rc;
    CUDA_TRY(c, cudaEventRecord(c->render_done[slot], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->render_done[slot], 0));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_ring[slot], dscreen, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_TRY(c, cudaEventRecord(c->copy_done[slot], c->copy_stream));
    c->ring_frame[slot] = c->iframe;
    c->submitted++;
    return B200PT_OK;
}

int b200pt_present_acquire(b200pt_context* c, const uint32_t** frame, int32_t* iframe)
{
    if (!c || !frame) return B200PT_ERR_INVALID_ARGUMENT;
    if (c->acquired >= c->submitted) return fail(c, B200PT_ERR_NOT_READY, "no frame in flight");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    const int slot = (int)(c->acquired % kRingSlots);
    CUDA_TRY(c, cudaEventSynchronize(c->copy_done[slot]));
    *frame = c->h_ring[slot];
    if (iframe) *iframe = c->ring_frame[slot];
    c->acquired++;
    return B200PT_OK;
}

int b200pt_present_blocking(b200pt_context* c, int32_t nframes, uint32_t* host_frame, int32_t bands)
{
    if (!c || nframes <= 0 || !host_frame || bands < -1) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (c->num_tiles > 0 || c->tile_mod > 1) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "a tile range is set: the blocking present renders the whole image");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    if (bands == -1) {
        // zero-copy: the kernel's tone-map epilogue stores every finished pixel straight into the caller's page-locked,
        // device-mapped frame over PCIe -- no copy operation at all, the transfer overlaps the render
        void* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, host_frame, 0) != cudaSuccess) {
            cudaGetLastError();
            return fail(c, B200PT_ERR_INVALID_ARGUMENT, "bands = -1 needs a page-locked, device-mapped host frame");
        }
        const int rc = render_frames_impl(c, nframes, static_cast<uint32_t*>(dptr));
        if (rc != B200PT_OK) return rc;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        return B200PT_OK;
    }
    if (bands == 0) bands = 2;  // measured on B200 (1080p, 1 frame): 0.345 / 0.300 / 0.301 / 0.321 ms for 1 / 2 / 3 / 4 bands
    if (bands > c->nty) bands = c->nty;
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));  // the present ring may still be reading d_screen (slot 0)
    const int first_iframe = c->iframe;
    const size_t row_bytes = (size_t)c->width * sizeof(uint32_t);
    int rc = B200PT_OK;
    for (int b = 0; b < bands && rc == B200PT_OK; b++) {
        const int row0 = (int)((long long)c->nty * b / bands), row1 = (int)((long long)c->nty * (b + 1) / bands);
        if (row1 == row0) continue;
        c->first_tile = row0 * c->ntx;
        c->num_tiles = (row1 - row0) * c->ntx;
        c->iframe = first_iframe;  // every band renders the same frames
        rc = render_frames_impl(c, nframes, c->d_screen);
        if (rc != B200PT_OK) break;
        const int slot = b % kRingSlots;
        const size_t off = (size_t)row0 * c->tile_h * c->width;  // first pixel of the band in the row-major frame
        const size_t bytes = (size_t)(row1 - row0) * c->tile_h * row_bytes;
        cudaError_t e = cudaEventRecord(c->render_done[slot], c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, c->render_done[slot], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host_frame + off, c->d_screen + off, bytes, cudaMemcpyDeviceToHost, c->copy_stream);
        if (e != cudaSuccess) {
            c->first_tile = c->num_tiles = 0;
            CUDA_TRY(c, e);
        }
    }
    c->first_tile = c->num_tiles = 0;
    // the bands' launches counted every path once; the frame counter advances once
    c->iframe = first_iframe + nframes;
    if (rc != B200PT_OK) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    return B200PT_OK;
}

int b200pt_bind_device_target(b200pt_context* c, void* device_ptr)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target_own) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    c->d_target = device_ptr ? static_cast<float*>(device_ptr) : c->d_target_own;
    return B200PT_OK;
}

int b200pt_get_device_target(b200pt_context* c, void** device_ptr, size_t* bytes)
{
    if (!c || !device_ptr) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    *device_ptr = c->d_target;
    if (bytes) *bytes = (size_t)c->width * c->height * 3 * sizeof(float);
    return B200PT_OK;
}

int b200pt_set_stream(b200pt_context* c, void* cuda_stream)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (collect_timing(c, true) != B200PT_OK) return B200PT_ERR_CUDA;
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return B200PT_OK;
}

int b200pt_set_tile_row_range(b200pt_context* c, int32_t first_tile_row, int32_t num_tile_rows)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (first_tile_row < 0 || num_tile_rows < 0 || first_tile_row + num_tile_rows > c->nty)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "tile row range outside the image");
    c->first_tile = first_tile_row * c->ntx;
    c->num_tiles = num_tile_rows * c->ntx;
    c->tile_mod = c->tile_rem = 0;
    return B200PT_OK;
}

int b200pt_set_tile_range(b200pt_context* c, int32_t first_flat_tile, int32_t num_tiles)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (first_flat_tile < 0 || num_tiles < 0 || first_flat_tile + num_tiles > c->ntx * c->nty)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "tile range outside the image");
    c->first_tile = first_flat_tile;
    c->num_tiles = num_tiles;
    c->tile_mod = c->tile_rem = 0;
    return B200PT_OK;
}

int b200pt_set_tile_stride(b200pt_context* c, int32_t remainder, int32_t modulus)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    if (modulus < 0 || remainder < 0 || (modulus > 0 && remainder >= modulus))
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "need 0 <= remainder < modulus (0, 0 = all tiles)");
    if (modulus > 1 && ((c->tile_w / 8) * c->tile_h) % 4)
        return fail(c, B200PT_ERR_INVALID_ARGUMENT, "a tile stride needs tiles of a multiple of 32 pixels");
    c->tile_mod = modulus > 1 ? modulus : 0;
    c->tile_rem = modulus > 1 ? remainder : 0;
    c->first_tile = c->num_tiles = 0;
    return B200PT_OK;
}

int b200pt_finalize_sum(b200pt_context* c, int32_t total_frames)
{
    if (!c || total_frames < 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    const float scale = 1.0f / ((float)total_frames + 1.f);
    CUDA_TRY(c, launch_scale(c->d_target, (size_t)c->width * c->height * 3, scale, c->stream));
    c->launches++;
    return B200PT_OK;
}

int b200pt_scale_target(b200pt_context* c, float factor)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, launch_scale(c->d_target, (size_t)c->width * c->height * 3, factor, c->stream));
    c->launches++;
    return B200PT_OK;
}

int b200pt_scale_target_span(b200pt_context* c, size_t float_offset, size_t float_count, float factor, void* cuda_stream)
{
    if (!c) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_target) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    const size_t nfl = (size_t)c->width * c->height * 3;
    if (float_offset > nfl || float_count > nfl - float_offset) return fail(c, B200PT_ERR_INVALID_ARGUMENT, "span outside the target");
    if (float_count == 0) return B200PT_OK;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, launch_scale(c->d_target + float_offset, float_count, factor, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream));
    c->launches++;
    return B200PT_OK;
}

int b200pt_download_rng_state(b200pt_context* c, uint32_t* host_dst)
{
    if (!c || !host_dst) return B200PT_ERR_INVALID_ARGUMENT;
    if (!c->d_rng) return fail(c, B200PT_ERR_NOT_READY, "resize first");
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaMemcpyAsync(host_dst, c->d_rng, (size_t)c->width * c->height * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return B200PT_OK;
}

int b200pt_eval_portable(b200pt_context* c, int fn, const float* a, const float* b, float* out, size_t n)
{
    const bool two = fn == B200PT_FN_ATAN2 || fn == B200PT_FN_POW;
    if (!c || !a || !out || ((fn < B200PT_FN_SIN || fn > B200PT_FN_EXP) && fn != B200PT_FN_POW) || (two && !b))
        return B200PT_ERR_INVALID_ARGUMENT;
    if (n == 0) return B200PT_OK;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc(&da, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&dout, n * sizeof(float));
    if (e == cudaSuccess && two) e = cudaMalloc(&db, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(da, a, n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && db) e = cudaMemcpyAsync(db, b, n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_eval_portable(fn, da, db, dout, n, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    CUDA_TRY(c, e);
    return B200PT_OK;
}

int b200pt_check_portable_tiers(b200pt_context* c, int fn, uint64_t first, uint64_t count, uint64_t* mismatches,
                                uint64_t* literal_path)
{
    if (!c || !mismatches || fn < B200PT_FN_ATAN2 || fn > B200PT_FN_EQUIRECT_TEXEL || fn == B200PT_FN_EXP) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    unsigned long long* d = nullptr;
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMalloc(&d, sizeof(h));
    if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, sizeof(h), c->stream);
    if (e == cudaSuccess) e = launch_check_portable_tiers(fn, first, count, d, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CUDA_TRY(c, e);
    *mismatches = h[0];
    if (literal_path) *literal_path = h[1];
    return B200PT_OK;
}

int b200pt_measure_fp32_peak(b200pt_context* c, double* tflops)
{
    if (!c || !tflops) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    const int blocks = c->sm_count * 4, iters = 4096;
    float* scratch = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&scratch, (size_t)blocks * 256 * sizeof(float));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4 && e == cudaSuccess; rep++) {  // first trip warms the clocks up
        e = cudaEventRecord(e0, c->stream);
        if (e == cudaSuccess) e = launch_ffma_peak(scratch, blocks, iters, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(e1, c->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && ms > 0.f) {
            const double flops = (double)blocks * 256.0 * (double)iters * 256.0 * 2.0;
            const double t = flops / (ms * 1e-3) * 1e-12;
            if (t > best) best = t;
        }
    }
    c->launches += 4;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(scratch);
    CUDA_TRY(c, e);
    *tflops = best;
    return B200PT_OK;
}

int b200pt_static_tables_match(int profile)
{
    switch (profile) {
    case B200PT_PROFILE_V2:
    case B200PT_PROFILE_SIMT_TEXTURED: {
        CornellScene s;
        build_cornell_scene(&s, profile == B200PT_PROFILE_SIMT_TEXTURED);
        return cornell_spheres_match_static_tables(s.sphere) ? 1 : 0;
    }
    case B200PT_PROFILE_OPT_V4: {
        V4Scene s;
        build_v4_scene(&s);
        return v4_scene_matches_static_tables(s) ? 1 : 0;
    }
    case B200PT_PROFILE_V3_REDO: {
        V3RedoScene s;
        build_v3redo_scene(&s);
        return v3redo_spheres_match_static_tables(s.sphere) ? 1 : 0;
    }
    case B200PT_PROFILE_V3_REDO_SCENE0: {
        V3RedoScene0 s;
        build_v3redo_scene0(&s);
        return v3redo0_spheres_match_static_tables(s.sphere) ? 1 : 0;
    }
    default: return -1;
    }
}

int b200pt_compute_cull_rects(int profile, int32_t width, int32_t height, float* rects, int32_t* count)
{
    if (!rects || !count || width <= 0 || height <= 0 || profile < B200PT_PROFILE_V2 || profile > B200PT_PROFILE_V3_REDO_SCENE0)
        return B200PT_ERR_INVALID_ARGUMENT;
    float4 r[kMaxCullRects];
    const int n = compute_cull_rects(profile, width, height, r);
    *count = n;
    for (int i = 0; i < n; i++) {
        rects[4 * i + 0] = r[i].x; rects[4 * i + 1] = r[i].y; rects[4 * i + 2] = r[i].z; rects[4 * i + 3] = r[i].w;
    }
    return B200PT_OK;
}

int b200pt_compute_cull_rects_scene_v4(const b200pt_quad* quads, int32_t num_quads, const b200pt_sphere* spheres,
                                       int32_t num_spheres, const b200pt_camera* camera, int32_t width, int32_t height,
                                       float* rects, int32_t* count)
{
    if (!rects || !count || !camera || width <= 0 || height <= 0 || num_quads < 0 || num_spheres < 0 ||
        num_quads + num_spheres > kV4MaxObjects || (num_quads && !quads) || (num_spheres && !spheres))
        return B200PT_ERR_INVALID_ARGUMENT;
    float4 r[kMaxCullRects];
    const int n = compute_cull_rects_v4(reinterpret_cast<const float*>(quads), num_quads, reinterpret_cast<const float*>(spheres),
                                        num_spheres, camera->Position, camera->Distance, width, height, r);
    *count = n;
    for (int i = 0; i < n; i++) {
        rects[4 * i + 0] = r[i].x; rects[4 * i + 1] = r[i].y; rects[4 * i + 2] = r[i].z; rects[4 * i + 3] = r[i].w;
    }
    return B200PT_OK;
}

int b200pt_get_counters(b200pt_context* c, b200pt_counters* out)
{
    if (!c || !out) return B200PT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    CUDA_TRY(c, guard.status);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (collect_timing(c, true) != B200PT_OK) return B200PT_ERR_CUDA;
    DeviceCounters dc{};
    CUDA_TRY(c, cudaMemcpy(&dc, c->d_counters, sizeof(dc), cudaMemcpyDeviceToHost));
    out->paths = c->paths;
    out->segments = dc.segments;
    out->escapes = dc.escapes;
    out->launches = c->launches;
    out->last_render_ms = c->last_render_ms;
    out->culled_segments = dc.culled;
    return B200PT_OK;
}

}  // extern "C"
