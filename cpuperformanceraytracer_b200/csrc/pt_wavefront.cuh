// pt_wavefront.cuh -- the CTA-sorted variant of the path-tracing megakernel.
//
// pt_render_kernel (pt_device.cuh) ties a pixel to a lane: in one loop trip the lanes of a warp whose path just
// missed (env lookup, fold the sample, start the pixel's next frame) and the lanes whose path hit something (shade,
// set up the bounce) are disjoint sets executed back to back, so on average half the lanes of an issued instruction
// are off (ncu: 16.9/32 for the v4 profile, 21.3/32 for the Cornell profile, round 1).
//
// Here a path is NOT tied to a lane.  One loop trip of a CTA is
//     A  per-role work      shade a hit | env lookup + fold + start the pixel's next path      (role-pure warps)
//     B  scene trace        every live path, whatever its role was                             (full warps)
//     C  sort               every thread writes its path record (5 x 16 bytes) to shared memory at the position its role
//                           gives it (ballots inside the warp, one shared-memory atomicAdd per role and warp, the role
//                           totals after a __syncthreads), __syncthreads, and reads the record at its own thread index
// so after C the threads [0, n0) hold the paths to shade, [n0, n1) the ones that also end at the bounce limit,
// [n1, n2) the misses, [n2, n3) the misses of camera-culled pixels (no trace either), and the rest are idle: at most
// one warp per role boundary is mixed.  A warp whose 32 threads are idle pulls the next 32-pixel work item.
//
// Measured on B200 (profiles/r02_a_*): 27.1 (v2) / 24.3 (v4) active lanes per instruction against 21.3 /
// 17.2 for pt_render_kernel, and 17 % / 4 % SLOWER: the sort costs 20-25 % of the issued instructions and its two barriers
// per trip take 10 points of issue-slot utilisation.  It is selectable (B200PT_SCHED_SORTED), not the default.
//
// What stays per pixel lives in shared memory, indexed by a slot number that travels with the path: the running
// average, the frame counter, the pixel coordinates.  One path per pixel is in flight at any time, so a pixel's
// samples are still folded in frame order with the reference's own expression: results are bit-identical to
// pt_render_kernel (and the oracle) -- only WHICH thread evaluates a segment changes.
//
// Reference paths are relative to /root/reference/CPUPerformanceRayTracer/.
#pragma once

#include "pt_device.cuh"

namespace b200pt {

constexpr int kWfThreads = 256;
constexpr int kWfWarps = kWfThreads / 32;
enum : int { kRoleShade = 0, kRoleShadeLast = 1, kRoleMiss = 2, kRoleMissCulled = 3, kRoleIdle = 4, kWfRoles = 5 };
constexpr int kWfRecordQuads = 5;   // a path record = 5 x 16 bytes: pos dir thr ret (12 words) rng meta dist normal (3) + 2 spare
constexpr int kCulledFramesPerTrip = 3;  // a camera-culled pixel never traces: its paths are folded several per trip

template <int PROFILE> struct WfShared {
    static constexpr int kFields = (PROFILE == kProfileV4 || is_v3redo(PROFILE)) ? kV4MatFields : kLegacyMatFields;
    float smat[kFields * kMatStride];
    typename SharedOf<PROFILE, kWfThreads>::type trace;  // Cornell-family trace: variant table + per-thread candidate stacks
    uint4 rec[kWfRecordQuads][kWfThreads];     // the sort's transit records (128-bit accesses, consecutive threads)
    float avg[3][kWfThreads];                  // per pixel slot: running average (or sum)
    int frame[kWfThreads];                     //                 next frame to fold
    uint32_t pixel[kWfThreads];                //                 x | yflip << 16
    int addr[kWfThreads];                      //                 float index of the pixel's R value in the target
    int cls[2][8];                             // paths per role of the current trip (double-buffered by trip parity)
};

// the sorted kernel keeps the round-2 launch bounds it was measured with (4 CTAs per SM for the v4 / v3_redo families)
template <int PROFILE> struct MinBlocksSorted {
    static constexpr int value = (PROFILE == kProfileV4 || is_v3redo(PROFILE)) ? 4 : B200PT_MIN_BLOCKS_CORNELL;
};

template <int PROFILE, int ENVK, int ENVS, int ACCUM, bool STATIC, class M>
__global__ void __launch_bounds__(kWfThreads, MinBlocksSorted<PROFILE>::value)
pt_render_sorted_kernel(const __grid_constant__ RenderParams p, const __grid_constant__ typename SceneOf<PROFILE>::type scene)
{
    constexpr int kFields = WfShared<PROFILE>::kFields;
    constexpr int kObjects = (PROFILE == kProfileV4) ? kV4MaxObjects : (PROFILE == kProfileV3Redo ? kV3Objects : (PROFILE == kProfileV3RedoS0 ? kV3S0Objects : kCornellObjects));
    extern __shared__ __align__(16) unsigned char wf_raw[];
    WfShared<PROFILE>& S = *reinterpret_cast<WfShared<PROFILE>*>(wf_raw);
    for (int i = threadIdx.x; i < kFields * kMatStride; i += blockDim.x) {
        const int field = i / kMatStride, obj = i % kMatStride;
        S.smat[i] = obj < kObjects ? reinterpret_cast<const float*>(&scene.mat[obj])[field] : 0.f;
    }
    if constexpr (PROFILE != kProfileV4) build_legacy_variants(S.trace, scene);
    else if constexpr (!STATIC) build_v4_shared(S.trace, scene);
    if (threadIdx.x < 16) (&S.cls[0][0])[threadIdx.x] = 0;
    __syncthreads();

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int frame_end = p.first_frame + p.nframes;
    unsigned nseg = 0, nesc = 0, ncull = 0;

    PathState s;
    s.pos = s.dir = s.thr = s.ret = mk(0.f, 0.f, 0.f);
    s.rng = 0;
    s.bounce = 0;
    Hit h;
    h.dist = c_superFar;
    h.normal = mk(0.f, 0.f, 0.f);
    h.matIndex = 0;
    h.fromInside = false;
    int slot = tid;        // the pixel slot this thread's path (or idle thread) owns
    int role = kRoleIdle;  // what phase A has to do with the record this thread holds
    bool culled = false;   // the path's pixel cannot hit anything (camera culling): no scene trace
    bool more_items = true;
    unsigned trip = 0;

    for (;;) {
        // ---- A: shade the traced segment; a finished path folds its sample and starts the pixel's next one -----
        // (the paths of a camera-culled pixel are sure misses: several of them are folded per trip)
        bool live = role != kRoleIdle;
        int reps = role == kRoleMissCulled ? kCulledFramesPerTrip : 1;
        while (live && reps-- > 0) {
            const bool done = shade_segment<PROFILE, ENVK, ENVS, STATIC, M>(s, h, p, S.smat, nesc);
            if (done) {
                v3 color;
                if constexpr (PROFILE == kProfileV4) color = fma3s(1.f, s.ret, mk(0.f, 0.f, 0.f));  // v4.cpp:1128
                else color = mk(0.f, 0.f, 0.f) + s.ret * 1.f;                                       // v2.cpp:565
                v3 avg = mk(S.avg[0][slot], S.avg[1][slot], S.avg[2][slot]);
                int frame = S.frame[slot];
                if constexpr (ACCUM == kAccumSum) {
                    avg = avg + color;
                } else {
                    const float blend = M::rcp_mid((float)frame + 1.f);  // 1.0f / f32(iFrame + 1.f)
                    if constexpr (PROFILE == kProfileV4) avg = fma3s(blend, color - avg, avg);      // v4.cpp:1239
                    else avg = lerp3(avg, color, blend);                                            // v2.cpp:623
                }
                frame++;
                const uint32_t xy = S.pixel[slot];
                const int x = (int)(xy & 0xffffu), yflip = (int)(xy >> 16);
                if (frame < frame_end) {
                    S.avg[0][slot] = avg.x;
                    S.avg[1][slot] = avg.y;
                    S.avg[2][slot] = avg.z;
                    S.frame[slot] = frame;
                    init_path<PROFILE, STATIC, M>(s, p, scene, x, yflip, frame);
                    if (reps > 0) {  // culled: this path's only segment is a miss as well
                        nseg++;
                        ncull++;
                    }
                } else {  // the pixel is finished: its slot is free again
                    float* px = p.target + S.addr[slot];
                    if (ACCUM == kAccumSum && p.scatter_gpo > 0) {  // fused reduce-scatter: see pt_render_kernel
                        const int a = S.addr[slot], g = a / 24, owner = g / p.scatter_gpo;
                        px = p.scatter_stage[owner] + (size_t)(g - owner * p.scatter_gpo) * 24 + (a - g * 24);
                    }
                    px[0] = avg.x;
                    px[8] = avg.y;
                    px[16] = avg.z;
                    const size_t idx = (size_t)(p.height - 1 - yflip) * p.width + x;
                    if (p.rng_out) p.rng_out[idx] = s.rng;
                    if (p.screen) p.screen[idx] = tonemap::pack<false>(avg.x, avg.y, avg.z, p.screen_mode);  // OUTPUT_TO_SCREEN
                    live = false;
                }
            }
        }
        // ---- a warp of 32 idle threads pulls the next work item (32 pixels, all frames of the launch) -----------
        if (more_items && __ballot_sync(0xffffffffu, live) == 0u) {
            int item = 0;
            if (lane == 0) item = atomicAdd(p.work_counter, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= p.num_items) {
                more_items = false;
            } else {
                if (p.item_order) item = __ldg(p.item_order + item);
                int gl = item * 4 + (lane >> 3);
                if (p.block_items) {  // an 8x4 pixel block: see pt_render_kernel
                    const int per_tile = p.groups_per_tile >> 2;
                    const int t = item / per_tile, it = item - t * per_tile;
                    const int band = it / p.groups_per_tile_row, gx = it - band * p.groups_per_tile_row;
                    gl = t * p.groups_per_tile + (band * 4 + (lane >> 3)) * p.groups_per_tile_row + gx;
                }
                if (gl < p.num_groups) {
                    const int g = p.group_offset + gl;
                    const int t = g / p.groups_per_tile, r = g - t * p.groups_per_tile;
                    const int ty = t / p.num_tiles_x, tx = t - ty * p.num_tiles_x;
                    const int ly = r / p.groups_per_tile_row, gx = r - ly * p.groups_per_tile_row;
                    const int x = tx * p.tile_w + gx * 8 + (lane & 7);
                    const int y = ty * p.tile_h + ly;
                    const int yflip = p.height - 1 - y;
                    const int a = g * 24 + (lane & 7);
                    const float* px = p.target + a;
                    const bool from_zero = ACCUM == kAccumSum && p.scatter_gpo > 0;
                    S.avg[0][slot] = from_zero ? 0.f : px[0];
                    S.avg[1][slot] = from_zero ? 0.f : px[8];
                    S.avg[2][slot] = from_zero ? 0.f : px[16];
                    S.frame[slot] = p.first_frame;
                    S.pixel[slot] = (uint32_t)x | ((uint32_t)yflip << 16);
                    S.addr[slot] = a;
                    bool sure_miss = p.num_cull_rects >= 0;
                    for (int k = 0; k < p.num_cull_rects; k++) {
                        const float4 rc = p.cull_rect[k];
                        if ((float)x + 0.5f >= rc.x && (float)x - 0.5f <= rc.z && (float)yflip + 0.5f >= rc.y && (float)yflip - 0.5f <= rc.w)
                            sure_miss = false;
                    }
                    culled = sure_miss;
                    init_path<PROFILE, STATIC, M>(s, p, scene, x, yflip, p.first_frame);
                    live = true;
                }
            }
        }
        // ---- B: one scene trace for every live path ---------------------------------------------------------------
        int key = kRoleIdle;
        if (live) {
            nseg++;
            h.dist = c_superFar;
            h.normal = mk(0.f, 0.f, 0.f);
            h.matIndex = 0;
            h.fromInside = false;
            if (!culled) trace_scene<PROFILE, STATIC, M>(s.pos, s.dir, h, scene, S.trace);
            else ncull++;
            const bool miss = (h.dist == c_superFar);
            key = miss ? (culled ? kRoleMissCulled : kRoleMiss) : (s.bounce >= p.num_bounces ? kRoleShadeLast : kRoleShade);
        }
        // ---- C: sort the CTA's paths by role ------------------------------------------------------------------------
        // inside the warp: ballots; across warps: the first lane of every role present reserves the warp's block of
        // that role with one shared-memory atomic (the order of the warps inside a role does not matter)
        const unsigned b0 = __ballot_sync(0xffffffffu, key == 0), b1 = __ballot_sync(0xffffffffu, key == 1);
        const unsigned b2 = __ballot_sync(0xffffffffu, key == 2), b3 = __ballot_sync(0xffffffffu, key == 3);
        unsigned mine = ~(b0 | b1 | b2 | b3);
        mine = key == 0 ? b0 : mine;
        mine = key == 1 ? b1 : mine;
        mine = key == 2 ? b2 : mine;
        mine = key == 3 ? b3 : mine;
        const int leader = __ffs(mine) - 1;
        int* cls = S.cls[trip & 1];
        int base = 0;
        if (lane == leader) base = atomicAdd(&cls[key], __popc(mine));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (tid == 0) *reinterpret_cast<int4*>(&S.cls[(trip & 1) ^ 1][0]) = make_int4(0, 0, 0, 0), S.cls[(trip & 1) ^ 1][4] = 0;
        __syncthreads();
        const int4 tot = *reinterpret_cast<const int4*>(&cls[0]);
        if (cls[kRoleIdle] == kWfThreads) break;  // nothing in flight and nothing left to pull (CTA-uniform)
        int dest = base + __popc(mine & lt_mask);
        dest += key > 0 ? tot.x : 0;
        dest += key > 1 ? tot.y : 0;
        dest += key > 2 ? tot.z : 0;
        dest += key > 3 ? tot.w : 0;
        {
            const uint32_t meta = (uint32_t)s.bounce | ((uint32_t)slot << 8) | ((uint32_t)h.matIndex << 16) |
                                  ((uint32_t)h.fromInside << 21) | ((uint32_t)culled << 22) | ((uint32_t)key << 24);
            S.rec[0][dest] = make_uint4(__float_as_uint(s.pos.x), __float_as_uint(s.pos.y), __float_as_uint(s.pos.z), __float_as_uint(s.dir.x));
            S.rec[1][dest] = make_uint4(__float_as_uint(s.dir.y), __float_as_uint(s.dir.z), __float_as_uint(s.thr.x), __float_as_uint(s.thr.y));
            S.rec[2][dest] = make_uint4(__float_as_uint(s.thr.z), __float_as_uint(s.ret.x), __float_as_uint(s.ret.y), __float_as_uint(s.ret.z));
            S.rec[3][dest] = make_uint4(s.rng, meta, __float_as_uint(h.dist), __float_as_uint(h.normal.x));
            if (key <= kRoleShadeLast) S.rec[4][dest] = make_uint4(__float_as_uint(h.normal.y), __float_as_uint(h.normal.z), 0u, 0u);
        }
        __syncthreads();
        {
            const uint4 r0 = S.rec[0][tid], r1 = S.rec[1][tid], r2 = S.rec[2][tid], r3 = S.rec[3][tid];
            s.pos = mk(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z));
            s.dir = mk(__uint_as_float(r0.w), __uint_as_float(r1.x), __uint_as_float(r1.y));
            s.thr = mk(__uint_as_float(r1.z), __uint_as_float(r1.w), __uint_as_float(r2.x));
            s.ret = mk(__uint_as_float(r2.y), __uint_as_float(r2.z), __uint_as_float(r2.w));
            s.rng = r3.x;
            const uint32_t meta = r3.y;
            h.dist = __uint_as_float(r3.z);
            s.bounce = (int)(meta & 0xffu);
            slot = (int)((meta >> 8) & 0xffu);
            h.matIndex = (int)((meta >> 16) & 0x1fu);
            h.fromInside = (meta >> 21) & 1u;
            culled = (meta >> 22) & 1u;
            role = (int)(meta >> 24);
            h.normal = mk(__uint_as_float(r3.w), 0.f, 0.f);
            if (role <= kRoleShadeLast) {  // only a hit carries a normal
                const uint4 r4 = S.rec[4][tid];
                h.normal.y = __uint_as_float(r4.x);
                h.normal.z = __uint_as_float(r4.y);
            }
        }
        trip++;
    }

    // one atomic per counter and warp (64-bit: 32 lanes x many items can exceed 2^32 segments)
    unsigned long long seg64 = nseg, esc64 = nesc, cull64 = ncull;
    for (int o = 16; o > 0; o >>= 1) {
        seg64 += __shfl_xor_sync(0xffffffffu, seg64, o);
        esc64 += __shfl_xor_sync(0xffffffffu, esc64, o);
        cull64 += __shfl_xor_sync(0xffffffffu, cull64, o);
    }
    if (lane == 0 && p.counters) {
        atomicAdd(&p.counters->segments, seg64);
        atomicAdd(&p.counters->escapes, esc64);
        if (cull64) atomicAdd(&p.counters->culled, cull64);
    }
}

// limits of the packed record: bounce count in 8 bits, coordinates in 16
inline bool sorted_kernel_supports(const RenderParams& rp)
{
    return rp.num_bounces >= 0 && rp.num_bounces <= 250 && rp.width <= 65535 && rp.height <= 65535;
}

// same configuration space as dispatch_config (pt_device.cuh), for the sorted kernel
template <class M, class F>
inline cudaError_t dispatch_config_sorted(const LaunchConfig& lc, F&& f)
{
#define B200PT_CASE(P, EK, ES, ST)                                                             \
    if (lc.accum_mode == kAccumSum) return f(pt_render_sorted_kernel<P, EK, ES, kAccumSum, ST, M>, sizeof(WfShared<P>)); \
    return f(pt_render_sorted_kernel<P, EK, ES, kAccumAverage, ST, M>, sizeof(WfShared<P>));
    if (lc.profile == kProfileV2) {
        if (lc.static_scene) { B200PT_CASE(kProfileV2, kEnvNone, kSamplerPoint, true) }
        B200PT_CASE(kProfileV2, kEnvNone, kSamplerPoint, false)
    }
    if (lc.profile == kProfileSimtTextured) {
        if (lc.static_scene) { B200PT_CASE(kProfileSimtTextured, kEnvEquirect, kSamplerPoint, true) }
        B200PT_CASE(kProfileSimtTextured, kEnvEquirect, kSamplerPoint, false)
    }
    if (lc.profile == kProfileV3Redo) {
        if (lc.static_scene) { B200PT_CASE(kProfileV3Redo, kEnvEquirect, kSamplerBilinear, true) }
        B200PT_CASE(kProfileV3Redo, kEnvEquirect, kSamplerBilinear, false)
    }
    if (lc.profile == kProfileV3RedoS0) {
        if (lc.static_scene) { B200PT_CASE(kProfileV3RedoS0, kEnvEquirect, kSamplerBilinear, true) }
        B200PT_CASE(kProfileV3RedoS0, kEnvEquirect, kSamplerBilinear, false)
    }
    if (lc.profile == kProfileV4) {
#define B200PT_V4CASE(EK, ES)                                   \
    if (lc.static_scene) { B200PT_CASE(kProfileV4, EK, ES, true) } \
    B200PT_CASE(kProfileV4, EK, ES, false)
        if (lc.env_kind == kEnvNone) { B200PT_V4CASE(kEnvNone, kSamplerPoint) }
        if (lc.env_kind == kEnvEquirect && lc.env_sampler == kSamplerRandom) { B200PT_V4CASE(kEnvEquirect, kSamplerRandom) }
        if (lc.env_kind == kEnvEquirect && lc.env_sampler == kSamplerBilinear) { B200PT_V4CASE(kEnvEquirect, kSamplerBilinear) }
        if (lc.env_kind == kEnvCubemap && lc.env_sampler == kSamplerRandom) { B200PT_V4CASE(kEnvCubemap, kSamplerRandom) }
        if (lc.env_kind == kEnvCubemap && lc.env_sampler == kSamplerBilinear) { B200PT_V4CASE(kEnvCubemap, kSamplerBilinear) }
#undef B200PT_V4CASE
    }
#undef B200PT_CASE
    return cudaErrorInvalidValue;
}

template <class M>
inline cudaError_t launch_sorted(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return dispatch_config_sorted<M>(lc, [&](auto kernel, size_t smem) -> cudaError_t {
        using KernelT = decltype(kernel);
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if constexpr (std::is_same<KernelT, void (*)(RenderParams, V4Scene)>::value) {
            kernel<<<lc.grid, kWfThreads, smem, stream>>>(rp, scenes.v4);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene)>::value) {
            kernel<<<lc.grid, kWfThreads, smem, stream>>>(rp, scenes.v3redo);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene0)>::value) {
            kernel<<<lc.grid, kWfThreads, smem, stream>>>(rp, scenes.v3redo0);
        } else {
            kernel<<<lc.grid, kWfThreads, smem, stream>>>(rp, scenes.cornell);
        }
        return cudaGetLastError();
    });
}

template <class M>
inline cudaError_t occupancy_sorted(const LaunchConfig& lc, int* blocks_per_sm)
{
    return dispatch_config_sorted<M>(lc, [&](auto kernel, size_t smem) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, kWfThreads, smem);
    });
}

}  // namespace b200pt
