/*
 * b200pt.h -- C ABI of the B200 path-tracing engine (libb200pt.so).
 *
 * This is the drop-in boundary for ONE hot path of torgeiba/CPUPerformanceRayTracer: the
 * per-pixel demofox path loop + env lookup + running-average accumulation that the reference
 * runs behind its render entry points.  Plain pointers and sizes only; no C++ or torch types.
 * Every entry point cites the reference interface it replaces (paths relative to
 * /root/reference/CPUPerformanceRayTracer/).  The C++ overloads with the reference's exact
 * names/signatures live in cpuperformanceraytracer_b200/host/demofox_render.h and forward here.
 *
 * There is no CPU fallback: every call that needs the GPU fails with B200PT_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PT_API_VERSION 2

typedef struct b200pt_context b200pt_context;

/* return codes (the reference's entry points are void and __debugbreak() on bad tiling,
 * Application.cpp:36-94; here every call reports) */
enum {
    B200PT_OK = 0,
    B200PT_ERR_INVALID_ARGUMENT = 1, /* null pointer, bad enum, tiling that CheckValidSettings rejects */
    B200PT_ERR_CUDA = 2,             /* CUDA runtime error or no usable device */
    B200PT_ERR_NOT_READY = 3,        /* render before resize / env-using profile without env */
    B200PT_ERR_OUT_OF_MEMORY = 4
};

/* which reference renderer's semantics the kernel reproduces */
enum {
    B200PT_PROFILE_V2 = 0,            /* DemofoxRenderV2, demofox_path_tracing_v2.cpp:654 (Cornell box,
                                         diffuse/emissive/specular, jitter, constant ambient) */
    B200PT_PROFILE_SIMT_TEXTURED = 1, /* DemofoxRenderSimtTextured, ..._simt_textured.cpp:560 (Cornell,
                                         diffuse/emissive, equirect point-sampled env) */
    B200PT_PROFILE_OPT_V4 = 2,        /* DemofoxRenderOptV4, ..._optimization_v4.cpp:1696 (7 spheres,
                                         Fresnel/refraction/absorption, equirect or cubemap env) */
    B200PT_PROFILE_V3_REDO = 3,       /* DemofoxRenderV3Redo, demofox_path_tracing_v3_redo.cpp:886, SCENE 1: the v4
                                         scene and shading without the approximations (exact divisions, exp(),
                                         sin/cos unit vectors, striped backdrop), bilinear equirect env, 8 bounces */
    B200PT_PROFILE_V3_REDO_SCENE0 = 4 /* the same renderer compiled with `#define SCENE 0` (v3_redo.cpp:379, :392-479,
                                         :530-580): the Cornell box seen from (0, 0, 40), three Fresnel-specular
                                         spheres, the light outside the box */
};

/* arithmetic policy */
enum {
    B200PT_MATH_PARITY = 0, /* IEEE div/sqrt, fused ops only where the reference writes fmadd/fmsub/
                               fnmadd, portable sin/cos/atan2/asin: bit-exact against the oracle */
    B200PT_MATH_FAST = 1    /* FMA contraction + MUFU approximations; same RNG streams and control
                               flow, ULP-level differences (tolerances in tests/test_gpu_parity.py) */
};

/* global_preprocessor_flags.h:56-57 USE_ENV_MAP / USE_ENV_CUBEMAP */
enum { B200PT_ENV_NONE = 0, B200PT_ENV_EQUIRECT = 1, B200PT_ENV_CUBEMAP = 2 };
/* texture.cpp:101 (point), :164/:275 (bilinear), :186/:341 (random jitter,
 * global_preprocessor_flags.h:66 USE_RANDOM_JITTER_TEXTURE_SAMPLING) */
enum { B200PT_SAMPLER_POINT = 0, B200PT_SAMPLER_BILINEAR = 1, B200PT_SAMPLER_RANDOM = 2 };

/* how frames are folded into the f32 target */
enum {
    B200PT_ACCUM_RUNNING_AVERAGE = 0, /* reference semantics: avg += (c - avg) / (iFrame + 1),
                                         ..._optimization_v4.cpp:1200,1239 / ..._v2.cpp:623 */
    B200PT_ACCUM_SUM = 1              /* target += c; used by spp-sharded multi-GPU renders, followed by
                                         a sum-reduce and b200pt_finalize_sum() */
};

/* Mapping of paths to GPU threads; the reference's counterpart is its tile queue + 8-wide masked lanes
 * (work_queue.cpp:7-66, RenderTile ..._optimization_v4.cpp:1189-1252).  Bit-identical results either way. */
enum {
    B200PT_SCHED_DEFAULT = 0,  /* the measured-faster one for the profile (environment variable B200PT_SCHEDULER=lane|sorted overrides) */
    B200PT_SCHED_LANE = 1,     /* one pixel per lane for the whole launch; a finished path restarts in place */
    B200PT_SCHED_SORTED = 2    /* every loop trip the CTA sorts its 256 paths by what they need next (shade / miss + restart),
                                  so warps are role-pure; per-pixel state lives in shared memory */
};

/* ScreenBufferData packing, ..._optimization_v4.cpp:1285-1290 (screen) / :1321-1325 (file) */
enum { B200PT_LDR_FILE_RGBA = 0, B200PT_LDR_SCREEN_BGRA = 1,
       /* OR into either packing (the values are b200pt_params.exact_tonemap << 1): */
       B200PT_LDR_EXACT_ACES = 2,  /* the exact ACES curve (USE_FAST_APPROXIMATE_ACES_TONEMAP 0, ..._optimization_v4.cpp:172-175)
                                      instead of the fast one (:168-171) */
       B200PT_LDR_EXACT_GAMMA = 4  /* 1.055 pow(c, 1/2.4) - 0.055 (USE_FAST_APPROXIMATE_GAMMA 0, :185) instead of fast_pow_gamma
                                      (:144-155); pow_ps is MSVC SVML in the reference, here the binary64 exp(y log x) the oracle
                                      defines (oracle/portable_math.h pm_powf: the correctly rounded power on every tested input) */ };
enum { B200PT_TONEMAP_EXACT_ACES = 1, B200PT_TONEMAP_EXACT_GAMMA = 2 }; /* bits of b200pt_params.exact_tonemap */

/* mirrors struct texture, texture.h:6-12 (row-major RGB f32, row 0 = bottom after stbi's flip) */
typedef struct b200pt_texture {
    const float* Data;
    int32_t Width;
    int32_t Height;
    int32_t Components; /* must be 3 */
} b200pt_texture;

typedef struct b200pt_params {
    int32_t struct_size;  /* sizeof(b200pt_params), for ABI evolution */
    int32_t device;       /* CUDA device ordinal */
    int32_t profile;      /* B200PT_PROFILE_* */
    int32_t math_mode;    /* B200PT_MATH_* */
    int32_t num_bounces;  /* c_numBounces (v2.cpp:22 = 4, v4.cpp:23 = 8); <0 = profile default */
    int32_t env_kind;     /* OPT_V4 only: B200PT_ENV_* (SIMT_TEXTURED always equirect/point) */
    int32_t env_sampler;  /* OPT_V4 only: B200PT_SAMPLER_BILINEAR or _RANDOM */
    int32_t accum_mode;   /* B200PT_ACCUM_* */
    int32_t output_to_screen; /* OUTPUT_TO_SCREEN (global_preprocessor_flags.h:58): tone-map into the
                                 screen buffer after every render call */
    int32_t disable_camera_culling; /* 1: trace the scene even for pixels whose jitter footprint provably
                                       misses every primitive (A/B measurements; results are identical) */
    int32_t generic_scene_tables;   /* 1: read the Cornell vertices from the scene table instead of the
                                       compile-time specialisation (A/B measurements; results are identical) */
    int32_t scheduler;              /* B200PT_SCHED_*: how paths are mapped to threads (results are identical) */
    int32_t disable_item_order;     /* 1: pull the work items in buffer order instead of scene-first / sky-last
                                       (A/B measurements; results are identical) */
    /* The reference's compile-time switches of global_preprocessor_flags.h:62-65, NON-default side; 0 keeps the checked-in
     * defaults (all "fast").  The first two are OPT_V4 only (the other profiles' sources do not read them): scene-specialised
     * kernels exist for every combination; with a run-time scene or B200PT_SCHED_SORTED the generic kernel reads them at run time. */
    int32_t exact_exp;              /* 1: USE_FAST_APPROXIMATE_EXP 0 -- Beer-Lambert absorption through exp_ps instead of the
                                       (1 + x/16.68)^16 approximation (..._optimization_v4.cpp:783-787) */
    int32_t sincos_unit_vectors;    /* 1: USE_UNIT_VECTOR_REJECTION_SAMPLING 0 -- RandomUnitVector (2 draws, sin/cos) and exact
                                       normalisations instead of the normalised cube sample (3 draws), :838-861 */
    int32_t exact_tonemap;          /* B200PT_TONEMAP_EXACT_ACES | _EXACT_GAMMA: USE_FAST_APPROXIMATE_ACES_TONEMAP 0 (:63) and / or
                                       USE_FAST_APPROXIMATE_GAMMA 0 (:62) -- every tone map of this context (resolve_ldr, present,
                                       OUTPUT_TO_SCREEN) uses the exact ACES curve (:172-175) / the pow() gamma (:185).  Any profile. */
} b200pt_params;

typedef struct b200pt_counters {
    uint64_t paths;     /* (pixel, frame) samples traced since create */
    uint64_t segments;  /* scene traces executed by live paths */
    uint64_t escapes;   /* paths that ended on a miss (one env lookup each) */
    uint64_t launches;  /* CUDA kernels launched by this library on this context */
    double last_render_ms; /* device time of the most recent finished b200pt_render_frames launch, CUDA events */
    uint64_t culled_segments; /* part of `segments`: one per path of a pixel whose camera ray provably misses the
                                 scene (camera culling) -- the kernel ran no scene trace for them */
} b200pt_counters;

/* Fills *p with the reference's checked-in defaults for `profile`
 * (global_preprocessor_flags.h:56-66, Appendix B of SURVEY.md). */
int b200pt_default_params(int profile, b200pt_params* p);

/* InitializeGlobalRenderResources (..._optimization_v4.cpp:1640-1661): camera, scene tables;
 * the thread pool it spawns is replaced by the persistent kernel's atomic work counter. */
int b200pt_create(const b200pt_params* params, b200pt_context** out_ctx);
int b200pt_destroy(b200pt_context* ctx);

/* Runtime scene description for B200PT_PROFILE_OPT_V4: the data the reference
 * hard-codes in InitializeScene / InitializeCamera (..._optimization_v4.cpp:1403-1502), installed
 * through the equivalents of AddQuadObjectToScene / AddSphereObjectToScene / AddMaterialToScene
 * (:1368-1401).  Object i owns material i; quads come first, then spheres; at most 12 objects
 * (MAX_OBJECTS / MAX_MATERIALS, :327-328).  Quad data is precomputed exactly like PrecomputeQuadData
 * (:269-319).  Like AddMaterialToScene, albedo[1] and albedo[2] are ignored (albedo[0] is stored in all
 * three channels, :1370-1372).  Pass num_quads = num_spheres = 0 to return to the built-in scene. */
typedef struct b200pt_quad { float V0[3], V1[3], V2[3], V3[3]; } b200pt_quad;          /* QuadSceneObject, :248-255 */
typedef struct b200pt_sphere { float PositionAndRadius[4]; } b200pt_sphere;               /* SphereSceneObject, :321-324 */
typedef struct b200pt_material {                                                         /* SceneMaterial, :351-362 */
    float albedo[3], emissive[3];
    float specularChance, specularRoughness;
    float specularColor[3];
    float IOR, refractionChance, refractionRoughness;
    float refractionColor[3];
} b200pt_material;
typedef struct b200pt_camera { float Position[3]; float Distance; } b200pt_camera;       /* Camera, :380-386 */
/* SMaterialInfo of the Cornell-family renderers (demofox_path_tracing_v2.cpp:39-50; simt_textured uses albedo / emissive) */
typedef struct b200pt_material_legacy {
    float albedo[3], emissive[3], specularColor[3];
    float percentSpecular, roughness;
} b200pt_material_legacy;
int b200pt_set_scene_v4(b200pt_context* ctx, const b200pt_quad* quads, int32_t num_quads, const b200pt_sphere* spheres,
                        int32_t num_spheres, const b200pt_material* materials, const b200pt_camera* camera);

/* Uploads the environment texture once (the reference re-passes `texture Texture` by value on
 * every render call, ..._optimization_v4.cpp:1696-1699).  Cubemaps are the W x 6H atlas that
 * LoadCubemapTexture builds (asset_loading.cpp:18-44). */
int b200pt_set_env(b200pt_context* ctx, b200pt_texture tex);

/* win32_offscreen_buffer::Resize + ReinitializeRenderTileData (Application.cpp:104-155,
 * ..._optimization_v4.cpp:1723-1726): (re)allocates the zeroed W*H*3 f32 target and the W*H u32
 * screen buffer in HBM and fixes the tile geometry.  Same validity rules as CheckValidSettings
 * (Application.cpp:36-94): W % ntx == 0, H % nty == 0, tile width % 8 == 0. */
int b200pt_resize(b200pt_context* ctx, int32_t width, int32_t height, int32_t num_tiles_x, int32_t num_tiles_y);

/* zero the target and the frame counter (the state a fresh Resize leaves behind) */
int b200pt_reset(b200pt_context* ctx);

/* static f32 iFrame (..._optimization_v4.cpp:34): number of render calls made so far */
int b200pt_set_frame_counter(b200pt_context* ctx, int32_t iframe);
int b200pt_get_frame_counter(b200pt_context* ctx, int32_t* iframe);

/* == nframes consecutive calls of DemofoxRenderOptV4 / DemofoxRenderV2 / DemofoxRenderSimtTextured
 * on the device-resident target: iFrame += 1, render every tile, fold into the target.
 * Asynchronous on the context's stream. */
int b200pt_render_frames(b200pt_context* ctx, int32_t nframes);
int b200pt_synchronize(b200pt_context* ctx);

/* copies of the f32 accumulation buffer in the reference's tile-major SoA8 layout
 * (RenderTile, ..._optimization_v4.cpp:1189-1252); W*H*3 floats */
int b200pt_upload_target(b200pt_context* ctx, const float* host_src);
int b200pt_download_target(b200pt_context* ctx, float* host_dst);

/* The reference-facing call: same arguments as DemofoxRenderOptV4
 * (demofox_path_tracing_optimization_v4.h:14-17) plus a frame count.  BufferOut is the caller's
 * HOST accumulation buffer (persists across calls); it is copied to the device, nframes are
 * rendered, and it is copied back, all inside the call.  Texture is uploaded when its pointer or
 * shape changed since the last call.  ScreenBufferData (may be NULL) receives the tone-mapped
 * u32 image when output_to_screen is set.  Blocks until done, like the reference. */
int b200pt_render_host(b200pt_context* ctx, float* BufferOut, int32_t BufferWidth, int32_t BufferHeight,
                       int32_t NumTilesX, int32_t NumTilesY, int32_t TileWidth, int32_t TileHeight,
                       int32_t NumChannels, b200pt_texture Texture, void* ScreenBufferData, int32_t nframes);

/* CopyOutputToFile / OutputToScreen (..._optimization_v4.cpp:1260-1331, :1729-1760):
 * ACES + sRGB + 8-bit pack of the device target into a row-major host u32[W*H].
 * Like the reference's CopyOutputToFile it also bumps the frame counter when
 * bump_frame_counter != 0 (..._optimization_v4.cpp:1741). */
int b200pt_resolve_ldr(b200pt_context* ctx, uint32_t* host_dst, int32_t mode, int32_t bump_frame_counter);

/* Progressive present path -- the windowed loop of ApplicationState::RunApp (Application.cpp:306-375):
 * render nframes (NUM_SAMPLES_PER_FRAME), tone-map (OutputToScreen packing, fused into the render
 * kernel) and copy the u32 frame to pinned host memory asynchronously.  Two frames can be in flight:
 * submit k+1 overlaps the copy of k (a third submit without an acquire fails with B200PT_ERR_NOT_READY).
 * acquire blocks until the oldest submitted frame is in host memory.  The ring has three slots, so the
 * returned pointer stays valid -- whatever is submitted meanwhile -- until the NEXT b200pt_present_acquire
 * (or resize / destroy). */
int b200pt_present_submit(b200pt_context* ctx, int32_t nframes);
int b200pt_present_acquire(b200pt_context* ctx, const uint32_t** frame, int32_t* iframe);
/* The same loop iteration as ONE blocking call (Application.cpp:330-360: Render, then the frame is in BackBuffer.Memory):
 * nframes render calls with the tone map fused into the kernel, the u32 frame (OutputToScreen packing) in
 * host_frame when the call returns.  The image is rendered in `bands` groups of tile rows (0 = the library's choice);
 * the rows of a finished band are contiguous in the row-major frame, so their copy to the host runs on the copy
 * engine while the next band renders: latency ~ render + one band's copy instead of render + tone map + whole copy.
 * host_frame should be page-locked (cudaHostAlloc / cudaHostRegister); pageable memory works but is copied
 * synchronously by the driver.  bands = -1: no copy operation at all -- the kernel stores every finished pixel
 * straight into the (page-locked, device-mapped) host frame over PCIe. */
int b200pt_present_blocking(b200pt_context* ctx, int32_t nframes, uint32_t* host_frame, int32_t bands);

/* multi-GPU plumbing: render into / reduce over a caller-owned device buffer (e.g. the storage of
 * a tensor handed to an NCCL all-reduce).  Pass NULL to go back to the internal buffer. */
int b200pt_bind_device_target(b200pt_context* ctx, void* device_ptr);
int b200pt_get_device_target(b200pt_context* ctx, void** device_ptr, size_t* bytes);
/* launch on a caller-owned cudaStream_t (NULL = the context's own stream) */
int b200pt_set_stream(b200pt_context* ctx, void* cuda_stream);
/* tile-shard: restrict render calls to tile rows [first, first + count) -- one contiguous span of the
 * band-major buffer (float offset first * TileHeight * W * 3).  (0, 0) = all rows (default after resize).
 * The render is bit-identical to the same rows of a full render. */
int b200pt_set_tile_row_range(b200pt_context* ctx, int32_t first_tile_row, int32_t num_tile_rows);
/* The same at tile granularity -- what the reference enqueues per work-queue entry
 * (AddWorkQueueEntry_Custom(&WorkData[FlatTileIndex]), ..._optimization_v4.cpp:1714-1718): restrict
 * render calls to the tiles FlatTileIndex in [first, first + count), FlatTileIndex = TileX + NumTilesX *
 * TileY.  Consecutive flat indices are contiguous in the buffer.  (0, 0) = all tiles. */
int b200pt_set_tile_range(b200pt_context* ctx, int32_t first_flat_tile, int32_t num_tiles);
/* ... or to every modulus-th tile: FlatTileIndex % modulus == remainder (tiles of N ranks interleave over the image, which
 * balances cheap sky tiles and expensive scene tiles without a cost model).  Needs tiles of a multiple of 32 pixels.
 * (0, 0) or modulus 1 = all tiles.  Replaces a tile range and vice versa. */
int b200pt_set_tile_stride(b200pt_context* ctx, int32_t remainder, int32_t modulus);
/* ACCUM_SUM epilogue: target *= 1/(total_frames + 1), the value the reference's running average
 * reaches after render calls 1..total_frames on a zeroed buffer.  total_frames is
 * the LAST frame index of the job, not the number of frames of one shard. */
int b200pt_finalize_sum(b200pt_context* ctx, int32_t total_frames);
/* target *= factor on the context's stream (ACCUM_SUM prologue of a continued job: a buffer holding the
 * running average after F frames, times (F + 1), is the sum of those F samples) */
int b200pt_scale_target(b200pt_context* ctx, float factor);
/* the same on floats [float_offset, float_offset + float_count) of the target, on `cuda_stream` (NULL = the
 * context's stream): the epilogue of ONE band of a band-pipelined spp-sharded render */
int b200pt_scale_target_span(b200pt_context* ctx, size_t float_offset, size_t float_count, float factor, void* cuda_stream);

/* ---- several GPUs of one box behind the same entry points ------------------------------------------
 * The reference fans one render call out to its worker threads below the entry point
 * (DemofoxRenderOptV4 -> AddWorkQueueEntry_Custom per tile -> CompleteAllWork_Custom,
 * ..._optimization_v4.cpp:1696-1721; MakeWorkQueue, work_queue.cpp:81-108).  A group does the same with
 * GPUs: ONE host process, one context per device, the frames (or tile rows) of a render call sharded
 * over them, the f32 accumulation buffers combined on device 0 of the group.
 *   B200PT_SHARD_SPP    every (pixel, iFrame) sample re-seeds its RNG (..._optimization_v4.cpp:1096-1101), so
 *                       rank r renders a contiguous block of the call's frame range into a SUM buffer; the N
 *                       buffers are summed into rank 0's and scaled by 1/(iFrame + 1).  Same samples as the
 *                       sequential render, different summation order: not bit-identical, the difference grows
 *                       like sqrt(frames) x 2^-24 relative (6e-7 at 16 frames, 6e-6 at 1024 frames, measured).
 *                       A continued job (iFrame = F > 0) first turns rank 0's average back into a sum
 *                       (x (F + 1)), so N more frames give exactly the reference's average after F + N calls.
 *   B200PT_SHARD_TILES  rank r renders tile rows [a_r, b_r) of every frame with the reference's running
 *                       average and the contiguous spans are copied into rank 0's buffer: bit-identical to
 *                       the single-GPU render.
 * How the SUM buffers meet (B200PT_SHARD_SPP):
 *   B200PT_COMBINE_NCCL ncclReduce(sum, root 0) over NVLink + one scale kernel on rank 0
 *   B200PT_COMBINE_PEER one kernel per GPU over NVLink peer memory: rank r sums slice r of all N buffers in
 *                       rank order (deterministic), scales, and stores the slice straight into rank 0's buffer
 *   B200PT_COMBINE_FUSED render and reduce-scatter in ONE kernel: the render kernel stores every finished pixel's sum straight
 *                       into a staging slot on the GPU that owns that part of the image (remote stores over NVLink, spread
 *                       over the whole launch), so when the renders end only a LOCAL sum of N slots is left on each owner
 *                       (rank order: deterministic) plus the store of its slice into rank 0's buffer.  Costs one extra
 *                       image-sized buffer per GPU; at most 16 ranks.
 * libnccl.so.2 is loaded on first use (dlopen); without it B200PT_COMBINE_NCCL fails with B200PT_ERR_NOT_READY. */
typedef struct b200pt_group b200pt_group;
enum { B200PT_SHARD_SPP = 0, B200PT_SHARD_TILES = 1 };
enum { B200PT_COMBINE_NCCL = 0, B200PT_COMBINE_PEER = 1, B200PT_COMBINE_FUSED = 2 };
/* params->device is ignored (devices[] rules); params->accum_mode is chosen by the sharding */
int b200pt_group_create(const b200pt_params* params, const int32_t* devices, int32_t num_devices, int32_t sharding,
                        int32_t combine, b200pt_group** out_group);
int b200pt_group_destroy(b200pt_group* group);
int b200pt_group_size(b200pt_group* group);
/* the per-device context (rank 0 holds the image): for set_scene_v4 on every rank, counters, streams */
b200pt_context* b200pt_group_context(b200pt_group* group, int32_t rank);
int b200pt_group_set_env(b200pt_group* group, b200pt_texture tex);
int b200pt_group_resize(b200pt_group* group, int32_t width, int32_t height, int32_t num_tiles_x, int32_t num_tiles_y);
int b200pt_group_reset(b200pt_group* group);
/* B200PT_SHARD_SPP: the image is rendered in `bands` groups of tile rows and the combine of band b (one contiguous
 * span of the buffer) runs on second streams while band b + 1 renders, so the exchange of a large image hides behind
 * the render.  0 (default) = 1 band: with hundreds of frames per call the exchange is < 1 % of the call even for an
 * 8192 x 8192 image; banding pays for short calls on large images. */
int b200pt_group_set_bands(b200pt_group* group, int32_t bands);
int b200pt_group_set_frame_counter(b200pt_group* group, int32_t iframe);
int b200pt_group_get_frame_counter(b200pt_group* group, int32_t* iframe);
/* == nframes consecutive render calls of the reference, sharded; asynchronous; the image lives on rank 0 */
int b200pt_group_render_frames(b200pt_group* group, int32_t nframes);
int b200pt_group_synchronize(b200pt_group* group);
int b200pt_group_upload_target(b200pt_group* group, const float* host_src);
int b200pt_group_download_target(b200pt_group* group, float* host_dst);
/* b200pt_render_host on a group: host accumulation buffer in, nframes sharded render calls, buffer out */
int b200pt_group_render_host(b200pt_group* group, float* BufferOut, int32_t BufferWidth, int32_t BufferHeight,
                             int32_t NumTilesX, int32_t NumTilesY, int32_t TileWidth, int32_t TileHeight,
                             int32_t NumChannels, b200pt_texture Texture, void* ScreenBufferData, int32_t nframes);
/* CopyOutputToFile on rank 0's image */
int b200pt_group_resolve_ldr(b200pt_group* group, uint32_t* host_dst, int32_t mode, int32_t bump_frame_counter);
/* counters summed over the ranks; last_render_ms = the longest rank's last launch; *combine_ms (may be NULL) =
 * device time between the end of rank 0's last render launch and the end of the combine (reduce + scale, or span
 * copies): the part of the exchange that is NOT hidden behind rendering (it includes waiting for slower ranks) */
int b200pt_group_get_counters(b200pt_group* group, b200pt_counters* out, double* combine_ms);
const char* b200pt_group_last_error(b200pt_group* group);

/* The Cornell-family scene (B200PT_PROFILE_V2 / _SIMT_TEXTURED) as data.  The reference writes it as literals inside its
 * trace function (demofox_path_tracing_v2.cpp:320-454, ..._simt_textured.cpp:278-385): exactly 6 quads (vertices A B C D,
 * scene translation already applied), 3 spheres, 9 materials (quads first).  The primitive counts are the renderer's own
 * (its trace is unrolled over them); positions, sizes and materials are free.  The generic kernel (vertex data read from
 * the scene table instead of immediates) renders it, bit-exact against the oracle on random scenes.  Coordinates must lie
 * within +-1e6, radii in [1e-3, 1e6].  quads == NULL: back to the reference's box.  Not thread-safe against in-flight
 * renders on the same context (it synchronises the stream first). */
int b200pt_set_scene_cornell(b200pt_context* ctx, const b200pt_quad* quads /* 6 */, const b200pt_sphere* spheres /* 3 */,
                             const b200pt_material_legacy* materials /* 9 */);
/* host-only: the culling rectangles such a scene gets (count < 0: culling impossible) */
int b200pt_compute_cull_rects_scene_cornell(const b200pt_quad* quads, const b200pt_sphere* spheres, int32_t width, int32_t height,
                                            float* rects, int32_t* count);

/* measurement hook: the FP32-pipe peak of this GPU as an FFMA micro-benchmark achieves it (TFLOP/s, FFMA = 2 flop; best of 4
 * launches of ~1e12 flop) -- the measured denominator next to the nominal 148 SMs x 128 lanes x 2 x clock */
int b200pt_measure_fp32_peak(b200pt_context* ctx, double* tflops);

/* debug/parity hook: u32 RNG state of every pixel after the last rendered frame's path ended
 * (row-major W*H, row 0 = top); checks wang_hash stream parity bit for bit */
int b200pt_download_rng_state(b200pt_context* ctx, uint32_t* host_dst);

/* debug/parity hooks for the parity-mode transcendentals (the stand-ins for the reference's SVML calls,
 * mathlib.h:449-499).  eval: out[i] = f(a[i] [, b[i]]) computed on the device exactly as the parity kernels
 * do; host buffers; b is only read by ATAN2 (atan2(a, b)).  check_tiers: the device compares the short
 * first-tier evaluation of ATAN2 / ASIN against the literal algorithm on `count` generated inputs starting
 * at number `first` (ASIN: input i = the binary32 bit pattern i, so first 0 / count 2^32 is exhaustive;
 * ATAN2: hashed pairs) and returns how many results differ (must be 0) and how many inputs took the
 * literal path.  SQRT / RCP: the kernels' unchecked square root / reciprocal sequences (used where the
 * operand is known to be a normal number of moderate size) against the IEEE operations over their whole
 * valid range, same enumeration as ASIN; literal_path counts the bit patterns outside the range, which
 * are skipped.  DIV: the unchecked division sequence against IEEE division on hashed operand pairs.
 * EQUIRECT_TEXEL: the texel index of the random-jitter equirect lookup obtained by bracketing binary32
 * approximations of the two angles, against the index from the exact angles, on hashed directions / jitters /
 * map sizes; a mismatch is also counted when an approximate angle strays more than a third of the bracket
 * half-width from the exact one; literal_path = lookups whose bracket was not decisive. */
enum { B200PT_FN_SIN = 0, B200PT_FN_COS = 1, B200PT_FN_ATAN2 = 2, B200PT_FN_ASIN = 3, B200PT_FN_EXP = 4,
       B200PT_FN_SQRT = 5, B200PT_FN_RCP = 6, B200PT_FN_DIV = 7, B200PT_FN_EQUIRECT_TEXEL = 8 /* 5-8: check_tiers only */,
       B200PT_FN_POW = 9 /* eval only: pow(a, b) of the exact-gamma tone map (B200PT_LDR_EXACT_GAMMA) */ };
int b200pt_eval_portable(b200pt_context* ctx, int fn, const float* a, const float* b, float* out, size_t n);
int b200pt_check_portable_tiers(b200pt_context* ctx, int fn, uint64_t first, uint64_t count, uint64_t* mismatches,
                                uint64_t* literal_path);

/* Host-only check (no GPU needed): 1 when the built-in scene of `profile`, as the host builds it, agrees with the
 * compile-time tables the scene-specialised kernels assume (sphere centres / radii as immediates, zero components of
 * the v4 quad tables); 0 = the library would fall back to the generic kernels; -1 = unknown profile. */
int b200pt_static_tables_match(int profile);

/* Host-only helper (no GPU needed): the conservative fragCoord-space rectangles (x0, y0, x1, y1;
 * y = flipped row index) outside of which a camera ray of `profile` cannot hit the scene; the kernel
 * skips the scene trace for such pixels.  rects must hold 4 * 12 floats; *count < 0 = no culling. */
int b200pt_compute_cull_rects(int profile, int32_t width, int32_t height, float* rects, int32_t* count);
/* same for a run-time OPT_V4 scene (see b200pt_set_scene_v4) */
int b200pt_compute_cull_rects_scene_v4(const b200pt_quad* quads, int32_t num_quads, const b200pt_sphere* spheres,
                                       int32_t num_spheres, const b200pt_camera* camera, int32_t width, int32_t height,
                                       float* rects, int32_t* count);

int b200pt_get_counters(b200pt_context* ctx, b200pt_counters* out);
const char* b200pt_last_error(b200pt_context* ctx);
const char* b200pt_error_string(int code);
int b200pt_api_version(void);

#ifdef __cplusplus
}
#endif
#endif
