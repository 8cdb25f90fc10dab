// b200pt_group.cu -- several GPUs of one box behind the render entry points (include/b200pt.h, "group").
//
// The reference fans one render call out to worker threads below the entry point
// (DemofoxRenderOptV4 -> AddWorkQueueEntry_Custom per tile -> CompleteAllWork_Custom,
// demofox_path_tracing_optimization_v4.cpp:1696-1721; MakeWorkQueue, work_queue.cpp:81-108) and its offline
// driver is a plain C++ loop over render calls (Application.cpp:400-458).  A group is the same fan-out over
// GPUs, in ONE host process and without torch: one b200pt_context per device, every launch asynchronous on that
// device's stream, cross-device ordering by CUDA events, the partial images combined either by NCCL (loaded with
// dlopen on first use) or by this library's own kernel over NVLink peer memory.
//
// Why frames shard: every (pixel, iFrame) sample re-seeds its RNG from (x, y, iFrame)
// (..._optimization_v4.cpp:1096-1101), so the frames of a render call are independent streams and only the
// accumulation couples them.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is resolved at run time (no link dependency)

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "b200pt_context.h"

namespace {

// ---- NCCL through dlopen -----------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;

    bool load()
    {
        if (handle) return true;
        // B200PT_NCCL_LIB overrides; the soname is what both the system package and torch's bundled copy carry
        // (inside a process that already imported torch, dlopen returns that already-loaded copy)
        const char* names[3] = {std::getenv("B200PT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return false;
        }
#define B200PT_NCCL_SYM(field, name)                                   \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));   \
    if (!field) {                                                      \
        error = std::string("libnccl has no symbol ") + name;          \
        dlclose(handle);                                               \
        handle = nullptr;                                              \
        return false;                                                  \
    }
        B200PT_NCCL_SYM(GetVersion, "ncclGetVersion")
        B200PT_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        B200PT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        B200PT_NCCL_SYM(GroupStart, "ncclGroupStart")
        B200PT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        B200PT_NCCL_SYM(Reduce, "ncclReduce")
        B200PT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef B200PT_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;

constexpr int kMaxGroup = 16;

// ---- K5: the group's own combine step over NVLink peer memory -----------------------------------------------
// Rank r owns float4 slice [begin, begin + count) of the image: it reads that slice of all N SUM buffers (its own
// from local HBM, N-1 over NVLink), adds them in rank order (a fixed summation order: the result does not depend
// on timing, unlike a ring/tree all-reduce whose order is the library's business), applies the 1/(iFrame+1) scale
// and stores the finished slice straight into rank 0's buffer.  HBM/NVLink bound: 16 N bytes read + 16 written
// per float4; every load of a thread's batch is issued before the first add.
struct PeerCombineParams {
    const float4* src[kMaxGroup];
    float4* dst;
    int nsrc;
    size_t begin, count;  // in float4
    float scale;
};

template <int N>
__global__ void __launch_bounds__(256) peer_combine_kernel(const __grid_constant__ PeerCombineParams p)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < p.count; k += stride) {
        const size_t i = p.begin + k;
        float4 v[N];
#pragma unroll
        for (int q = 0; q < N; q++) v[q] = __ldg(p.src[q] + i);
        float4 a = v[0];
#pragma unroll
        for (int q = 1; q < N; q++) {
            a.x += v[q].x;
            a.y += v[q].y;
            a.z += v[q].z;
            a.w += v[q].w;
        }
        a.x *= p.scale;
        a.y *= p.scale;
        a.z *= p.scale;
        a.w *= p.scale;
        p.dst[i] = a;
    }
}

// ---- K6: the owner's half of the FUSED render + reduce-scatter (B200PT_COMBINE_FUSED) ------------------------------------
// The render kernels of all ranks have already stored their partial sums of THIS rank's part of the image into this rank's
// staging buffer, slot q for rank q (RenderParams::scatter_stage: remote stores over NVLink issued pixel by pixel while the
// render runs, so the exchange is spread over the whole launch instead of following it).  What is left is local: add the
// nsrc slots in rank order, add rank 0's previous image times (F + 1) (the sum it stands for), scale by 1/(iFrame + 1), and
// store the finished slice into rank 0's buffer.  16 nsrc B of local HBM reads + one 16 B remote read and write per float4.
struct StagedCombineParams {
    const float4* stage;  // this rank's staging buffer: nslots slots of `slot4` float4
    float4* image;        // rank 0's target + this rank's slice
    size_t slot4, count4;
    int nsrc;
    float prev_factor, scale;
};

__global__ void __launch_bounds__(256) staged_combine_kernel(const __grid_constant__ StagedCombineParams p)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.count4; i += stride) {
        const float4 prev = p.image[i];
        float4 a = __ldg(p.stage + i);
        for (int q = 1; q < p.nsrc; q++) {
            const float4 v = __ldg(p.stage + (size_t)q * p.slot4 + i);
            a.x += v.x;
            a.y += v.y;
            a.z += v.z;
            a.w += v.w;
        }
        a.x = (a.x + prev.x * p.prev_factor) * p.scale;
        a.y = (a.y + prev.y * p.prev_factor) * p.scale;
        a.z = (a.z + prev.z * p.prev_factor) * p.scale;
        a.w = (a.w + prev.w * p.prev_factor) * p.scale;
        p.image[i] = a;
    }
}

cudaError_t launch_staged_combine(const StagedCombineParams& p, int sm_count, cudaStream_t stream)
{
    if (p.count4 == 0) return cudaSuccess;
    const size_t want = (p.count4 + 255) / 256, cap = (size_t)sm_count * 8;
    staged_combine_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_peer_combine(const PeerCombineParams& p, int sm_count, cudaStream_t stream)
{
    if (p.count == 0) return cudaSuccess;
    const int block = 256;
    size_t want = (p.count + block - 1) / block;
    const size_t cap = (size_t)sm_count * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    switch (p.nsrc) {
#define B200PT_PC(N) case N: peer_combine_kernel<N><<<grid, block, 0, stream>>>(p); break;
        B200PT_PC(1) B200PT_PC(2) B200PT_PC(3) B200PT_PC(4) B200PT_PC(5) B200PT_PC(6) B200PT_PC(7) B200PT_PC(8)
        B200PT_PC(9) B200PT_PC(10) B200PT_PC(11) B200PT_PC(12) B200PT_PC(13) B200PT_PC(14) B200PT_PC(15) B200PT_PC(16)
#undef B200PT_PC
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

struct b200pt_group {
    int n = 0, sharding = B200PT_SHARD_SPP, combine = B200PT_COMBINE_NCCL;
    b200pt_context* ctx[kMaxGroup] = {};
    int device[kMaxGroup] = {};
    ncclComm_t comm[kMaxGroup] = {};
    bool have_comm = false;
    bool distinct_devices = true;
    cudaEvent_t render_done[kMaxGroup] = {}, combine_done[kMaxGroup] = {};
    cudaStream_t comm_stream[kMaxGroup] = {};  // per rank: the combine of band b runs here while band b + 1 renders
    int bands = 0;                             // spp sharding: 0 = by buffer size (b200pt_group_set_bands)
    cudaEvent_t cb0 = nullptr, cb1 = nullptr;  // on rank 0's device: the combine step
    bool combine_timing_pending = false;
    double combine_ms = 0.0;
    int width = 0, height = 0, ntx = 0, nty = 0;
    int tile_first[kMaxGroup] = {}, tile_count[kMaxGroup] = {};  // B200PT_SHARD_TILES: flat tile ranges ...
    float* stage[kMaxGroup] = {};   // B200PT_COMBINE_FUSED: per rank, n slots of one slice each
    size_t stage_floats = 0;        // size of every stage
    int stage_gpo = 0;              // SoA8 groups per owner
    bool tile_stride = false;  // ... or rank r renders the tiles with FlatTileIndex % n == r (interleaved: balanced by construction)
    bool peer_all = true;      // every rank can address rank 0's memory
    int iframe = 0;
    uint64_t launches = 0;  // combine kernels (the contexts count their own)
    float* h_pinned = nullptr;  // staging of b200pt_group_render_host for pageable caller buffers
    size_t pinned_floats = 0;
    std::string last_error;
};

namespace {

int gfail(b200pt_group* g, int code, const std::string& msg)
{
    if (g) g->last_error = msg;
    return code;
}

// a failing per-context call: keep its message
int cfail(b200pt_group* g, int rank, int code, const char* what)
{
    g->last_error = std::string(what) + " on rank " + std::to_string(rank) + ": " + b200pt_last_error(g->ctx[rank]);
    return code;
}

#define GROUP_CUDA(g, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return gfail(g, e_ == cudaErrorMemoryAllocation ? B200PT_ERR_OUT_OF_MEMORY : B200PT_ERR_CUDA, \
                         std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
    } while (0)

#define GROUP_CTX(g, r, expr)                                 \
    do {                                                      \
        const int rc_ = (expr);                               \
        if (rc_ != B200PT_OK) return cfail(g, r, rc_, #expr); \
    } while (0)

size_t image_floats(const b200pt_group* g) { return (size_t)g->width * g->height * 3; }

// contiguous, near-equal blocks of `total` units (frames, tiles, float4s) for rank r of n
void block_of(size_t total, int n, int r, size_t* first, size_t* count)
{
    const size_t base = total / (size_t)n, rem = total % (size_t)n;
    *count = base + ((size_t)r < rem ? 1 : 0);
    *first = (size_t)r * base + ((size_t)r < rem ? (size_t)r : rem);
}

int collect_combine_timing(b200pt_group* g)
{
    if (g->combine_timing_pending) {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        float ms = 0.f;
        GROUP_CUDA(g, cudaEventSynchronize(g->cb1));
        GROUP_CUDA(g, cudaEventElapsedTime(&ms, g->cb0, g->cb1));
        g->combine_ms = ms;
        g->combine_timing_pending = false;
    }
    return B200PT_OK;
}

bool host_pointer_is_pinned(const void* ptr)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// rank r's stream waits for event `ev` (recorded on another device's stream)
int stream_wait(b200pt_group* g, int r, cudaEvent_t ev)
{
    DeviceGuard guard(g->device[r]);
    GROUP_CUDA(g, guard.status);
    GROUP_CUDA(g, cudaStreamWaitEvent(g->ctx[r]->stream, ev, 0));
    return B200PT_OK;
}

int record(b200pt_group* g, int r, cudaEvent_t ev)
{
    DeviceGuard guard(g->device[r]);
    GROUP_CUDA(g, guard.status);
    GROUP_CUDA(g, cudaEventRecord(ev, g->ctx[r]->stream));
    return B200PT_OK;
}

// ---- B200PT_SHARD_SPP ---------------------------------------------------------------------------------------
// The image is rendered in `bands` groups of tile rows.  A band is one contiguous span of the tile-major buffer
// (RenderTile, ..._optimization_v4.cpp:1189-1194), so as soon as every rank has rendered band b its spans are
// combined on the ranks' second streams while band b + 1 renders: for a large image (config 5: 805 MB of f32) the
// exchange disappears behind the render instead of following it.
int auto_bands(const b200pt_group* g)
{
    // Default: one band.  Measured on 2 B200 (config 5: 8192 x 8192, 805 MB buffer, 256 spp per GPU): the exchange after
    // the render is 2.0 ms of 512 ms, and cutting the render into 8 launches costs more (+1.3 %) than hiding it saves.
    // Banding pays when the render per call is short against the exchange (few frames per call on a large image).
    if (g->bands <= 0) return 1;
    return g->bands < g->nty ? g->bands : g->nty;
}

// B200PT_COMBINE_FUSED: the render kernels scatter their partial sums to the owners' stages while they run (K1 epilogue),
// the owners finish locally (K6).  Rank 0's buffer keeps the image (running average) between calls; the other ranks'
// targets are not used at all.
int render_spp_fused(b200pt_group* g, int nframes)
{
    const int F = g->iframe, n = g->n;
    const size_t slice = (size_t)g->stage_gpo * 24, nfl = image_floats(g);
    int active = 0;
    for (int r = 0; r < n; r++) {
        size_t first, count;
        block_of((size_t)nframes, n, r, &first, &count);
        b200pt_context* c = g->ctx[r];
        c->scatter_gpo = g->stage_gpo;
        for (int o = 0; o < n; o++) c->scatter_stage[o] = g->stage[o] + (size_t)r * slice;  // owner o's stage, slot r
        c->iframe = F + (int)first;
        int rc = B200PT_OK;
        if (count > 0) {
            rc = b200pt_render_frames(c, (int32_t)count);
            active = r + 1;  // block_of hands the frames to the first ranks: the active ones are 0 .. active-1
        }
        c->scatter_gpo = 0;
        c->iframe = F + nframes;
        if (rc != B200PT_OK) return cfail(g, r, rc, "b200pt_render_frames");
        rc = record(g, r, g->render_done[r]);
        if (rc != B200PT_OK) return rc;
    }
    {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb0, g->ctx[0]->stream));
    }
    for (int r = 0; r < n; r++)
        for (int q = 0; q < n; q++)
            if (q != r) {
                const int rc = stream_wait(g, r, g->render_done[q]);
                if (rc != B200PT_OK) return rc;
            }
    for (int r = 0; r < n; r++) {
        const size_t begin = (size_t)r * slice;
        if (begin >= nfl) continue;
        StagedCombineParams sc{};
        sc.stage = reinterpret_cast<const float4*>(g->stage[r]);
        sc.image = reinterpret_cast<float4*>(g->ctx[0]->d_target + begin);
        sc.slot4 = slice / 4;
        sc.count4 = ((begin + slice <= nfl) ? slice : nfl - begin) / 4;
        sc.nsrc = active;
        sc.prev_factor = (float)F + 1.f;  // rank 0 holds A_F = (A_0 + sum of F samples) / (F + 1)
        sc.scale = 1.0f / ((float)(F + nframes) + 1.f);
        DeviceGuard guard(g->device[r]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, launch_staged_combine(sc, g->ctx[r]->sm_count, g->ctx[r]->stream));
        g->launches++;
        GROUP_CUDA(g, cudaEventRecord(g->combine_done[r], g->ctx[r]->stream));
    }
    // the next call's render kernels store into these stages again: every rank waits for every owner's combine
    for (int r = 0; r < n; r++)
        for (int q = 0; q < n; q++)
            if (q != r) {
                const int rc = stream_wait(g, r, g->combine_done[q]);
                if (rc != B200PT_OK) return rc;
            }
    {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb1, g->ctx[0]->stream));
        g->combine_timing_pending = true;
    }
    g->iframe = F + nframes;
    return B200PT_OK;
}

int render_spp(b200pt_group* g, int nframes)
{
    const int F = g->iframe, n = g->n;
    const size_t nfl = image_floats(g);
    if (g->combine == B200PT_COMBINE_FUSED && n > 1) return render_spp_fused(g, nframes);
    const int bands = n > 1 ? auto_bands(g) : 1;
    const size_t per_tile_row = (size_t)g->width * (size_t)(g->height / g->nty) * 3;
    // rank 0 holds the running average after F calls, A_F = (A_0 + sum of the F samples) / (F + 1) (the reference's
    // blend factor is 1/(iFrame + 1)).  A_F * (F + 1) is the sum the N new samples are added to.
    if (F > 0) GROUP_CTX(g, 0, b200pt_scale_target(g->ctx[0], (float)F + 1.f));
    for (int r = 1; r < n; r++) {
        b200pt_context* c = g->ctx[r];
        DeviceGuard guard(g->device[r]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaMemsetAsync(c->d_target, 0, nfl * sizeof(float), c->stream));
    }
    const float scale = 1.0f / ((float)(F + nframes) + 1.f);  // == b200pt_finalize_sum(F + nframes)
    for (int b = 0; b < bands; b++) {
        const int row0 = (int)((long long)g->nty * b / bands), row1 = (int)((long long)g->nty * (b + 1) / bands);
        const size_t off = (size_t)row0 * per_tile_row, cnt = (size_t)(row1 - row0) * per_tile_row;
        for (int r = 0; r < n; r++) {
            size_t first, count;
            block_of((size_t)nframes, n, r, &first, &count);
            b200pt_context* c = g->ctx[r];
            c->first_tile = bands > 1 ? row0 * g->ntx : 0;
            c->num_tiles = bands > 1 ? (row1 - row0) * g->ntx : 0;
            c->iframe = F + (int)first;
            if (count > 0) GROUP_CTX(g, r, b200pt_render_frames(c, (int32_t)count));
            c->first_tile = c->num_tiles = 0;
            c->iframe = F + nframes;
            if (n > 1) {
                const int rc = record(g, r, g->render_done[r]);
                if (rc != B200PT_OK) return rc;
            }
        }
        if (n == 1) continue;
        // the band's combine, on the second streams
        for (int r = 0; r < n; r++) {
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            if (g->combine == B200PT_COMBINE_NCCL) {
                GROUP_CUDA(g, cudaStreamWaitEvent(g->comm_stream[r], g->render_done[r], 0));
            } else {
                for (int q = 0; q < n; q++) GROUP_CUDA(g, cudaStreamWaitEvent(g->comm_stream[r], g->render_done[q], 0));
            }
        }
        if (g->combine == B200PT_COMBINE_NCCL) {
            ncclResult_t nr = g_nccl.GroupStart();
            for (int r = 0; r < n && nr == ncclSuccess; r++)
                nr = g_nccl.Reduce(g->ctx[r]->d_target + off, g->ctx[r]->d_target + off, cnt, ncclFloat, ncclSum, 0, g->comm[r], g->comm_stream[r]);
            const ncclResult_t ne = g_nccl.GroupEnd();
            if (nr == ncclSuccess) nr = ne;
            if (nr != ncclSuccess) return gfail(g, B200PT_ERR_CUDA, std::string("ncclReduce: ") + g_nccl.GetErrorString(nr));
            DeviceGuard guard(g->device[0]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, launch_scale(g->ctx[0]->d_target + off, cnt, scale, g->comm_stream[0]));
            g->launches++;
        } else {
            for (int r = 0; r < n; r++) {
                PeerCombineParams pc{};
                pc.nsrc = n;
                for (int q = 0; q < n; q++) pc.src[q] = reinterpret_cast<const float4*>(g->ctx[q]->d_target + off);
                pc.dst = reinterpret_cast<float4*>(g->ctx[0]->d_target + off);
                block_of(cnt / 4, n, r, &pc.begin, &pc.count);  // a tile row holds a multiple of 24 floats
                pc.scale = scale;
                DeviceGuard guard(g->device[r]);
                GROUP_CUDA(g, guard.status);
                GROUP_CUDA(g, launch_peer_combine(pc, g->ctx[r]->sm_count, g->comm_stream[r]));
                g->launches++;
            }
        }
    }
    {   // combine_ms = what is left of the combine after rank 0's last render launch has finished (the exposed part)
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb0, g->ctx[0]->stream));
    }
    if (n == 1) {
        GROUP_CTX(g, 0, b200pt_finalize_sum(g->ctx[0], F + nframes));
    } else {
        // every rank's next touch of its SUM buffer (and rank 0's readers) must come after every combine that reads it
        for (int r = 0; r < n; r++) {
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, cudaEventRecord(g->combine_done[r], g->comm_stream[r]));
        }
        for (int r = 0; r < n; r++)
            for (int q = 0; q < n; q++) {
                const int rc = stream_wait(g, r, g->combine_done[q]);
                if (rc != B200PT_OK) return rc;
            }
    }
    {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb1, g->ctx[0]->stream));
        g->combine_timing_pending = true;
    }
    g->iframe = F + nframes;
    return B200PT_OK;
}

// ---- B200PT_SHARD_TILES -------------------------------------------------------------------------------------
// float span of rank r's tiles in the tile-major buffer (tiles follow one another in FlatTileIndex order,
// RenderTile, ..._optimization_v4.cpp:1189-1194)
void tile_span(const b200pt_group* g, int r, size_t* offset, size_t* count)
{
    const size_t per_tile = (size_t)(g->width / g->ntx) * (size_t)(g->height / g->nty) * 3;
    *offset = (size_t)g->tile_first[r] * per_tile;
    *count = (size_t)g->tile_count[r] * per_tile;
}

int render_tiles(b200pt_group* g, int nframes, bool gather)
{
    const int F = g->iframe, n = g->n;
    for (int r = 0; r < n; r++) {
        b200pt_context* c = g->ctx[r];
        c->iframe = F;
        if (g->tile_count[r] > 0) GROUP_CTX(g, r, b200pt_render_frames(c, nframes));
        c->iframe = F + nframes;
    }
    {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb0, g->ctx[0]->stream));
    }
    if (gather) {
        const size_t per_tile = (size_t)(g->width / g->ntx) * (size_t)(g->height / g->nty) * 3;
        for (int r = 1; r < n; r++) {
            if (g->tile_count[r] == 0) continue;
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            if (g->tile_stride && g->peer_all) {
                // one kernel per rank: its interleaved tiles stored straight into rank 0's buffer over NVLink peer memory
                GROUP_CUDA(g, launch_tile_gather(g->ctx[r]->d_target, g->ctx[0]->d_target, g->ntx * g->nty, n, r, per_tile, g->ctx[r]->sm_count,
                                                 g->ctx[r]->stream));
                g->launches++;
            } else if (g->tile_stride) {
                for (int t = r; t < g->ntx * g->nty; t += n)
                    GROUP_CUDA(g, cudaMemcpyPeerAsync(g->ctx[0]->d_target + (size_t)t * per_tile, g->device[0], g->ctx[r]->d_target + (size_t)t * per_tile,
                                                      g->device[r], per_tile * sizeof(float), g->ctx[r]->stream));
            } else {
                size_t off, cnt;
                tile_span(g, r, &off, &cnt);
                GROUP_CUDA(g, cudaMemcpyPeerAsync(g->ctx[0]->d_target + off, g->device[0], g->ctx[r]->d_target + off, g->device[r],
                                                  cnt * sizeof(float), g->ctx[r]->stream));
            }
            GROUP_CUDA(g, cudaEventRecord(g->combine_done[r], g->ctx[r]->stream));
            const int rc = stream_wait(g, 0, g->combine_done[r]);
            if (rc != B200PT_OK) return rc;
        }
    }
    {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaEventRecord(g->cb1, g->ctx[0]->stream));
        g->combine_timing_pending = true;
    }
    g->iframe = F + nframes;
    return B200PT_OK;
}

int sync_all(b200pt_group* g)
{
    for (int r = 0; r < g->n; r++) GROUP_CTX(g, r, b200pt_synchronize(g->ctx[r]));
    return collect_combine_timing(g);
}

}  // namespace

extern "C" {

int b200pt_group_create(const b200pt_params* params, const int32_t* devices, int32_t num_devices, int32_t sharding,
                        int32_t combine, b200pt_group** out_group)
{
    if (!params || !devices || !out_group || num_devices < 1 || num_devices > kMaxGroup) return B200PT_ERR_INVALID_ARGUMENT;
    *out_group = nullptr;
    if (sharding != B200PT_SHARD_SPP && sharding != B200PT_SHARD_TILES) return B200PT_ERR_INVALID_ARGUMENT;
    if (combine != B200PT_COMBINE_NCCL && combine != B200PT_COMBINE_PEER && combine != B200PT_COMBINE_FUSED) return B200PT_ERR_INVALID_ARGUMENT;
    if (combine == B200PT_COMBINE_FUSED && num_devices > kMaxScatterRanks) return B200PT_ERR_INVALID_ARGUMENT;
    b200pt_group* g = new (std::nothrow) b200pt_group();
    if (!g) return B200PT_ERR_OUT_OF_MEMORY;
    g->n = num_devices;
    g->sharding = sharding;
    g->combine = combine;
    for (int r = 0; r < g->n; r++) {
        g->device[r] = devices[r];
        for (int q = 0; q < r; q++)
            if (devices[q] == devices[r]) g->distinct_devices = false;
    }
    // Several ranks on one device are legal for the library's own combine and for tile sharding (the ranks are
    // contexts with their own streams and buffers): that is how the sharding logic is tested on a one-GPU box.
    // NCCL wants one device per rank.
    if (!g->distinct_devices && sharding == B200PT_SHARD_SPP && combine == B200PT_COMBINE_NCCL && g->n > 1) {
        delete g;
        return B200PT_ERR_INVALID_ARGUMENT;
    }
    int rc = B200PT_OK;
    for (int r = 0; r < g->n && rc == B200PT_OK; r++) {
        b200pt_params p = *params;
        p.device = devices[r];
        p.accum_mode = (sharding == B200PT_SHARD_SPP) ? B200PT_ACCUM_SUM : B200PT_ACCUM_RUNNING_AVERAGE;
        rc = b200pt_create(&p, &g->ctx[r]);
    }
    for (int r = 0; r < g->n && rc == B200PT_OK; r++) {
        DeviceGuard guard(g->device[r]);
        if (guard.status != cudaSuccess || cudaEventCreateWithFlags(&g->render_done[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->combine_done[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaStreamCreateWithFlags(&g->comm_stream[r], cudaStreamNonBlocking) != cudaSuccess)
            rc = B200PT_ERR_CUDA;
        // peer access both ways: the combine kernel reads every rank's buffer and writes rank 0's
        for (int q = 0; q < g->n && rc == B200PT_OK; q++) {
            if (g->device[q] == g->device[r]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->device[r], g->device[q]) != cudaSuccess) can = 0;
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(g->device[q], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) can = 0;
            }
            if (!can) g->peer_all = false;
            if (!can && sharding == B200PT_SHARD_SPP && (combine == B200PT_COMBINE_PEER || combine == B200PT_COMBINE_FUSED)) rc = B200PT_ERR_CUDA;
        }
    }
    if (rc == B200PT_OK) {
        DeviceGuard guard(g->device[0]);
        if (guard.status != cudaSuccess || cudaEventCreate(&g->cb0) != cudaSuccess || cudaEventCreate(&g->cb1) != cudaSuccess)
            rc = B200PT_ERR_CUDA;
    }
    if (rc == B200PT_OK && sharding == B200PT_SHARD_SPP && combine == B200PT_COMBINE_NCCL && g->n > 1) {
        if (!g_nccl.load()) {
            std::fprintf(stderr, "b200pt: %s\n", g_nccl.error.c_str());
            rc = B200PT_ERR_NOT_READY;
        } else {
            const ncclResult_t nr = g_nccl.CommInitAll(g->comm, g->n, g->device);
            if (nr != ncclSuccess) {
                std::fprintf(stderr, "b200pt: ncclCommInitAll: %s\n", g_nccl.GetErrorString(nr));
                rc = B200PT_ERR_CUDA;
            } else {
                g->have_comm = true;
            }
        }
    }
    if (rc != B200PT_OK) {
        b200pt_group_destroy(g);
        return rc;
    }
    *out_group = g;
    return B200PT_OK;
}

int b200pt_group_destroy(b200pt_group* g)
{
    if (!g) return B200PT_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < g->n; r++)
        if (g->ctx[r]) b200pt_synchronize(g->ctx[r]);
    if (g->have_comm)
        for (int r = 0; r < g->n; r++)
            if (g->comm[r]) g_nccl.CommDestroy(g->comm[r]);
    for (int r = 0; r < g->n; r++) {
        DeviceGuard guard(g->device[r]);
        if (g->render_done[r]) cudaEventDestroy(g->render_done[r]);
        if (g->combine_done[r]) cudaEventDestroy(g->combine_done[r]);
        if (g->comm_stream[r]) {
            cudaStreamSynchronize(g->comm_stream[r]);
            cudaStreamDestroy(g->comm_stream[r]);
        }
    }
    {
        DeviceGuard guard(g->device[0]);
        if (g->cb0) cudaEventDestroy(g->cb0);
        if (g->cb1) cudaEventDestroy(g->cb1);
        if (g->h_pinned) cudaFreeHost(g->h_pinned);
    }
    for (int r = 0; r < g->n; r++) {
        if (g->stage[r]) {
            DeviceGuard guard(g->device[r]);
            cudaFree(g->stage[r]);
        }
        if (g->ctx[r]) b200pt_destroy(g->ctx[r]);
    }
    delete g;
    return B200PT_OK;
}

int b200pt_group_size(b200pt_group* g) { return g ? g->n : 0; }

b200pt_context* b200pt_group_context(b200pt_group* g, int32_t rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

const char* b200pt_group_last_error(b200pt_group* g) { return g ? g->last_error.c_str() : "null group"; }

int b200pt_group_set_env(b200pt_group* g, b200pt_texture tex)
{
    if (!g) return B200PT_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < g->n; r++) GROUP_CTX(g, r, b200pt_set_env(g->ctx[r], tex));
    return B200PT_OK;
}

int b200pt_group_resize(b200pt_group* g, int32_t width, int32_t height, int32_t ntx, int32_t nty)
{
    if (!g) return B200PT_ERR_INVALID_ARGUMENT;
    for (int r = 0; r < g->n; r++) GROUP_CTX(g, r, b200pt_resize(g->ctx[r], width, height, ntx, nty));
    g->width = width;
    g->height = height;
    g->ntx = ntx;
    g->nty = nty;
    g->iframe = 0;
    if (g->sharding == B200PT_SHARD_SPP && g->combine == B200PT_COMBINE_FUSED && g->n > 1) {
        const long long groups = (long long)width * height / 8;
        g->stage_gpo = (int)((groups + g->n - 1) / g->n);
        const size_t need = (size_t)g->stage_gpo * 24 * (size_t)g->n;
        if (need != g->stage_floats) {
            for (int r = 0; r < g->n; r++) {
                DeviceGuard guard(g->device[r]);
                GROUP_CUDA(g, guard.status);
                if (g->stage[r]) cudaFree(g->stage[r]);
                g->stage[r] = nullptr;
            }
            g->stage_floats = 0;
            for (int r = 0; r < g->n; r++) {
                DeviceGuard guard(g->device[r]);
                GROUP_CUDA(g, guard.status);
                GROUP_CUDA(g, cudaMalloc(&g->stage[r], need * sizeof(float)));
            }
            g->stage_floats = need;
        }
    }
    g->tile_stride = false;
    if (g->sharding == B200PT_SHARD_TILES && g->n > 1 && (((width / ntx) / 8) * (height / nty)) % 4 == 0) {
        // Interleaved tiles: rank r renders the tiles with FlatTileIndex % n == r, in one launch (the pull-order table lists
        // them).  Cheap sky tiles and expensive scene tiles spread evenly over the ranks without any cost model.
        g->tile_stride = true;
        for (int r = 0; r < g->n; r++) {
            g->tile_first[r] = r;
            g->tile_count[r] = r < ntx * nty ? (ntx * nty - r + g->n - 1) / g->n : 0;
            GROUP_CTX(g, r, b200pt_set_tile_stride(g->ctx[r], r, g->n));
        }
    } else if (g->sharding == B200PT_SHARD_TILES) {
        // Tiles of another shape (items would straddle tiles): contiguous flat-tile ranges (one span of the buffer per rank) of near-equal COST, not near-equal size: a tile
        // of sky pixels (camera-culled: no scene trace) is ~10x cheaper than a tile looking into the scene, and the
        // reference's images have the sky at the top and the scene in the middle.
        const int ntiles = ntx * nty;
        std::vector<double> cost((size_t)ntiles, 1.0);
        float4 rects[kMaxCullRects];
        const b200pt_context* c0 = g->ctx[0];
        const int nrects = (c0->custom_scene || c0->params.disable_camera_culling) ? -1 : compute_cull_rects(c0->params.profile, width, height, rects);
        if (nrects >= 0) {
            const int tw = width / ntx, th = height / nty;
            for (int t = 0; t < ntiles; t++) {
                const int tx = t % ntx, ty = t / ntx;
                int traced = 0, total = 0;
                for (int ly = 0; ly < th; ly += 4)      // a 4x4 sub-sample of the tile's pixels is plenty for a weight
                    for (int lx = 0; lx < tw; lx += 4) {
                        const float x = (float)(tx * tw + lx), yflip = (float)(height - 1 - (ty * th + ly));
                        bool hit = false;
                        for (int k = 0; k < nrects; k++)
                            if (x + 0.5f >= rects[k].x && x - 0.5f <= rects[k].z && yflip + 0.5f >= rects[k].y && yflip - 0.5f <= rects[k].w) hit = true;
                        traced += hit ? 1 : 0;
                        total++;
                    }
                cost[(size_t)t] = (double)(total - traced) + 10.0 * (double)traced;
            }
        }
        double sum = 0.0;
        for (double v : cost) sum += v;
        int t = 0;
        double acc = 0.0;
        for (int r = 0; r < g->n; r++) {
            const double target = sum * (double)(r + 1) / (double)g->n;
            const int first = t;
            while (t < ntiles && (r == g->n - 1 || acc + 0.5 * cost[(size_t)t] <= target)) acc += cost[(size_t)t++];
            g->tile_first[r] = first;
            g->tile_count[r] = t - first;
            // (0, 0) would mean "all tiles": a rank without tiles simply never launches
            if (t > first) GROUP_CTX(g, r, b200pt_set_tile_range(g->ctx[r], (int32_t)first, (int32_t)(t - first)));
        }
    }
    return B200PT_OK;
}

int b200pt_group_set_bands(b200pt_group* g, int32_t bands)
{
    if (!g || bands < 0) return B200PT_ERR_INVALID_ARGUMENT;
    g->bands = bands;
    return B200PT_OK;
}

int b200pt_group_reset(b200pt_group* g)
{
    if (!g) return B200PT_ERR_INVALID_ARGUMENT;
    if (!g->width) return gfail(g, B200PT_ERR_NOT_READY, "resize first");
    for (int r = 0; r < g->n; r++) GROUP_CTX(g, r, b200pt_reset(g->ctx[r]));
    g->iframe = 0;
    return B200PT_OK;
}

int b200pt_group_set_frame_counter(b200pt_group* g, int32_t iframe)
{
    if (!g || iframe < 0) return B200PT_ERR_INVALID_ARGUMENT;
    g->iframe = iframe;
    for (int r = 0; r < g->n; r++) g->ctx[r]->iframe = iframe;
    return B200PT_OK;
}

int b200pt_group_get_frame_counter(b200pt_group* g, int32_t* iframe)
{
    if (!g || !iframe) return B200PT_ERR_INVALID_ARGUMENT;
    *iframe = g->iframe;
    return B200PT_OK;
}

int b200pt_group_render_frames(b200pt_group* g, int32_t nframes)
{
    if (!g || nframes < 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (!g->width) return gfail(g, B200PT_ERR_NOT_READY, "resize first");
    if (nframes == 0) return B200PT_OK;
    if (collect_combine_timing(g) != B200PT_OK) return B200PT_ERR_CUDA;
    return g->sharding == B200PT_SHARD_SPP ? render_spp(g, nframes) : render_tiles(g, nframes, true);
}

int b200pt_group_synchronize(b200pt_group* g)
{
    if (!g) return B200PT_ERR_INVALID_ARGUMENT;
    return sync_all(g);
}

int b200pt_group_upload_target(b200pt_group* g, const float* host_src)
{
    if (!g || !host_src) return B200PT_ERR_INVALID_ARGUMENT;
    if (!g->width) return gfail(g, B200PT_ERR_NOT_READY, "resize first");
    // spp: the running average lives on rank 0 (the other ranks' SUM buffers are zeroed by every render call);
    // tiles: every rank keeps the running average of its own tiles, so each one needs its span
    if (g->sharding == B200PT_SHARD_SPP) {
        GROUP_CTX(g, 0, b200pt_upload_target(g->ctx[0], host_src));
        return B200PT_OK;
    }
    for (int r = 0; r < g->n; r++) GROUP_CTX(g, r, b200pt_upload_target(g->ctx[r], host_src));
    return B200PT_OK;
}

int b200pt_group_download_target(b200pt_group* g, float* host_dst)
{
    if (!g || !host_dst) return B200PT_ERR_INVALID_ARGUMENT;
    if (!g->width) return gfail(g, B200PT_ERR_NOT_READY, "resize first");
    int rc = sync_all(g);
    if (rc != B200PT_OK) return rc;
    GROUP_CTX(g, 0, b200pt_download_target(g->ctx[0], host_dst));
    return B200PT_OK;
}

int b200pt_group_render_host(b200pt_group* g, float* BufferOut, int32_t W, int32_t H, int32_t NumTilesX, int32_t NumTilesY,
                             int32_t TileWidth, int32_t TileHeight, int32_t NumChannels, b200pt_texture Texture,
                             void* ScreenBufferData, int32_t nframes)
{
    if (!g || !BufferOut || nframes < 0) return B200PT_ERR_INVALID_ARGUMENT;
    if (NumChannels != 3) return gfail(g, B200PT_ERR_INVALID_ARGUMENT, "NumChannels must be 3");
    if (NumTilesX <= 0 || NumTilesY <= 0 || TileWidth * NumTilesX != W || TileHeight * NumTilesY != H)
        return gfail(g, B200PT_ERR_INVALID_ARGUMENT, "tiles must cover the buffer exactly");
    if (g->width != W || g->height != H || g->ntx != NumTilesX || g->nty != NumTilesY) {
        const int keep = g->iframe;  // the reference's static iFrame survives a resize
        const int rc = b200pt_group_resize(g, W, H, NumTilesX, NumTilesY);
        if (rc != B200PT_OK) return rc;
        b200pt_group_set_frame_counter(g, keep);
    }
    if (Texture.Data) {
        b200pt_context* c0 = g->ctx[0];
        const b200pt_params& pp = c0->params;
        const bool uses = pp.profile == B200PT_PROFILE_SIMT_TEXTURED || pp.profile == B200PT_PROFILE_V3_REDO ||
                          (pp.profile == B200PT_PROFILE_OPT_V4 && pp.env_kind != B200PT_ENV_NONE);
        if (uses && (Texture.Data != c0->last_env_ptr || Texture.Width != c0->env_w || Texture.Height != c0->env_h)) {
            const int rc = b200pt_group_set_env(g, Texture);
            if (rc != B200PT_OK) return rc;
        }
    }
    const size_t nfl = image_floats(g);
    float* src = BufferOut;
    const bool direct = host_pointer_is_pinned(BufferOut);
    if (!direct) {
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        if (g->pinned_floats != nfl) {
            if (g->h_pinned) cudaFreeHost(g->h_pinned);
            g->h_pinned = nullptr;
            g->pinned_floats = 0;
            GROUP_CUDA(g, cudaMallocHost(&g->h_pinned, nfl * sizeof(float)));
            g->pinned_floats = nfl;
        }
        std::memcpy(g->h_pinned, BufferOut, nfl * sizeof(float));
        src = g->h_pinned;
    }
    if (collect_combine_timing(g) != B200PT_OK) return B200PT_ERR_CUDA;
    if (g->sharding == B200PT_SHARD_SPP) {
        b200pt_context* c0 = g->ctx[0];
        {
            DeviceGuard guard(g->device[0]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, cudaMemcpyAsync(c0->d_target, src, nfl * sizeof(float), cudaMemcpyHostToDevice, c0->stream));
        }
        if (nframes > 0) {
            const int rc = render_spp(g, nframes);
            if (rc != B200PT_OK) return rc;
        }
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaMemcpyAsync(src, c0->d_target, nfl * sizeof(float), cudaMemcpyDeviceToHost, c0->stream));
    } else if (g->tile_stride) {
        // interleaved tiles: every rank takes the whole state over its own PCIe link (its tiles are spread all over the
        // buffer), the finished tiles meet on rank 0 (tile_gather_kernel), one copy back
        for (int r = 0; r < g->n; r++) {
            if (g->tile_count[r] == 0) continue;
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, cudaMemcpyAsync(g->ctx[r]->d_target, src, nfl * sizeof(float), cudaMemcpyHostToDevice, g->ctx[r]->stream));
        }
        if (nframes > 0) {
            const int rc = render_tiles(g, nframes, true);
            if (rc != B200PT_OK) return rc;
        }
        DeviceGuard guard(g->device[0]);
        GROUP_CUDA(g, guard.status);
        GROUP_CUDA(g, cudaMemcpyAsync(src, g->ctx[0]->d_target, nfl * sizeof(float), cudaMemcpyDeviceToHost, g->ctx[0]->stream));
    } else {
        // contiguous tile ranges need no GPU-to-GPU traffic at all on this path: every rank moves its own span of the
        // caller's buffer over its own PCIe link, in both directions
        for (int r = 0; r < g->n; r++) {
            size_t off, cnt;
            tile_span(g, r, &off, &cnt);
            if (cnt == 0) continue;
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, cudaMemcpyAsync(g->ctx[r]->d_target + off, src + off, cnt * sizeof(float), cudaMemcpyHostToDevice,
                                          g->ctx[r]->stream));
        }
        if (nframes > 0) {
            const int rc = render_tiles(g, nframes, false);
            if (rc != B200PT_OK) return rc;
        }
        for (int r = 0; r < g->n; r++) {
            size_t off, cnt;
            tile_span(g, r, &off, &cnt);
            if (cnt == 0) continue;
            DeviceGuard guard(g->device[r]);
            GROUP_CUDA(g, guard.status);
            GROUP_CUDA(g, cudaMemcpyAsync(src + off, g->ctx[r]->d_target + off, cnt * sizeof(float), cudaMemcpyDeviceToHost,
                                          g->ctx[r]->stream));
        }
    }
    int rc = sync_all(g);
    if (rc != B200PT_OK) return rc;
    if (!direct) std::memcpy(BufferOut, g->h_pinned, nfl * sizeof(float));
    if (ScreenBufferData && g->ctx[0]->params.output_to_screen) {
        // OUTPUT_TO_SCREEN: the tone-mapped frame of the combined image (the per-rank kernels only saw partial sums)
        if (g->sharding == B200PT_SHARD_TILES && !g->tile_stride) {
            // rank 0 needs the other ranks' spans for the tone map
            GROUP_CTX(g, 0, b200pt_upload_target(g->ctx[0], BufferOut));
        }
        GROUP_CTX(g, 0, b200pt_resolve_ldr(g->ctx[0], static_cast<uint32_t*>(ScreenBufferData), B200PT_LDR_SCREEN_BGRA, 0));
    }
    return B200PT_OK;
}

int b200pt_group_resolve_ldr(b200pt_group* g, uint32_t* host_dst, int32_t mode, int32_t bump_frame_counter)
{
    if (!g || !host_dst) return B200PT_ERR_INVALID_ARGUMENT;
    if (!g->width) return gfail(g, B200PT_ERR_NOT_READY, "resize first");
    int rc = sync_all(g);
    if (rc != B200PT_OK) return rc;
    GROUP_CTX(g, 0, b200pt_resolve_ldr(g->ctx[0], host_dst, mode, 0));
    if (bump_frame_counter) b200pt_group_set_frame_counter(g, g->iframe + 1);  // CopyOutputToFile: iFrame += 1.0f, v4.cpp:1741
    return B200PT_OK;
}

int b200pt_group_get_counters(b200pt_group* g, b200pt_counters* out, double* combine_ms)
{
    if (!g || !out) return B200PT_ERR_INVALID_ARGUMENT;
    int rc = sync_all(g);
    if (rc != B200PT_OK) return rc;
    std::memset(out, 0, sizeof(*out));
    for (int r = 0; r < g->n; r++) {
        b200pt_counters c{};
        GROUP_CTX(g, r, b200pt_get_counters(g->ctx[r], &c));
        out->paths += c.paths;
        out->segments += c.segments;
        out->escapes += c.escapes;
        out->culled_segments += c.culled_segments;
        out->launches += c.launches;
        if (c.last_render_ms > out->last_render_ms) out->last_render_ms = c.last_render_ms;
    }
    out->launches += g->launches;
    if (combine_ms) *combine_ms = g->combine_ms;
    return B200PT_OK;
}

}  // extern "C"
