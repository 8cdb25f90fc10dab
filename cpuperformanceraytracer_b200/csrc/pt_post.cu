// pt_post.cu -- the small HBM-bound kernels around the megakernel.
//   resolve_ldr : OutputToScreen / OutputToFile (demofox_path_tracing_optimization_v4.cpp:1260-1331)
//                 with ACESFilm (:166-176), LinearToSRGB (:178-187) and fast_pow_gamma (:144-155),
//                 fast variants (global_preprocessor_flags.h:62-63), exact rcp like the oracle.
//   scale       : ACCUM_SUM epilogue after the cross-GPU sum (target *= 1/(N+1)).
//   pack_env    : RGB f32 rows -> RGBA32F texels for the env texture object.
//   build_item_order : the pull order of a launch's work items (scene-first, sky-last).
// Compiled with --fmad=false so the tone map rounds like the reference's explicit fmadd sequence.
#include "pt_common.cuh"
#include "pt_tonemap.cuh"

namespace b200pt {

// one thread per pixel; a warp reads 4 SoA8 groups (384 contiguous bytes) and writes 32
// consecutive u32 of one image row
__global__ void resolve_ldr_kernel(const float* __restrict__ target, uint32_t* __restrict__ out, int width, int height,
                                   int tile_w, int tile_h, int num_tiles_x, int mode)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)width * height;
    if (gid >= npix) return;
    const long long g = gid >> 3;
    const int l = (int)(gid & 7);
    const int groups_per_tile_row = tile_w / 8, groups_per_tile = groups_per_tile_row * tile_h;
    const int t = (int)(g / groups_per_tile), r = (int)(g - (long long)t * groups_per_tile);
    const int ty = t / num_tiles_x, tx = t - ty * num_tiles_x;
    const int ly = r / groups_per_tile_row, gx = r - ly * groups_per_tile_row;
    const int x = tx * tile_w + gx * 8 + l, y = ty * tile_h + ly;
    const float* px = target + g * 24 + l;
    out[(size_t)y * width + x] = tonemap::pack<true>(px[0], px[8], px[16], mode);
}

__global__ void scale_kernel(float* __restrict__ t, size_t n, float scale)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n; k += stride) t[k] = t[k] * scale;
}

__global__ void pack_env_kernel(const float* __restrict__ rgb, float4* __restrict__ rgba, size_t texels)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < texels; k += stride) rgba[k] = make_float4(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2], 0.f);
}

// Pull order of a launch's work items: the items whose pixel block touches a culling rectangle (they trace the
// scene) first, in buffer order, then the sky-only ones, from the end of the table backwards.  Two small launches:
// per-block counts, then every block sums the counts of the blocks before it and places its items (deterministic).
__device__ __forceinline__ bool item_touches_scene(const RenderParams& rp, int item)
{
    // pixel bounds of the item: an 8x4 block, or 4 consecutive groups of a tile row (pt_render_kernel's mapping)
    int x0, x1, y0, y1;
    auto pix = [&](int g, int& x, int& y) {
        const int tile = g / rp.groups_per_tile, r = g - tile * rp.groups_per_tile;
        const int ty = tile / rp.num_tiles_x, tx = tile - ty * rp.num_tiles_x;
        const int ly = r / rp.groups_per_tile_row, gx = r - ly * rp.groups_per_tile_row;
        x = tx * rp.tile_w + gx * 8;
        y = ty * rp.tile_h + ly;
    };
    if (rp.block_items) {
        const int per_tile = rp.groups_per_tile >> 2;
        const int t = item / per_tile, it = item - t * per_tile;
        const int band = it / rp.groups_per_tile_row, gx = it - band * rp.groups_per_tile_row;
        pix(rp.group_offset + t * rp.groups_per_tile + band * 4 * rp.groups_per_tile_row + gx, x0, y0);
        x1 = x0 + 7;
        y1 = y0 + 3;
    } else {
        int gl1 = item * 4 + 3;
        if (gl1 >= rp.num_groups) gl1 = rp.num_groups - 1;
        int xa, ya, xb, yb;
        pix(rp.group_offset + item * 4, xa, ya);
        pix(rp.group_offset + gl1, xb, yb);
        if (ya == yb) { x0 = xa; x1 = xb + 7; y0 = y1 = ya; }
        else { x0 = 0; x1 = rp.width - 1; y0 = min(ya, yb); y1 = max(ya, yb); }  // wraps a row / a tile: be conservative
    }
    const float fx0 = (float)x0 - 0.5f, fx1 = (float)x1 + 0.5f;
    const float fy0 = (float)(rp.height - 1 - y1) - 0.5f, fy1 = (float)(rp.height - 1 - y0) + 0.5f;  // flipped rows
    bool hit = false;
    for (int k = 0; k < rp.num_cull_rects; k++)
        hit = hit || (fx1 >= rp.cull_rect[k].x && fx0 <= rp.cull_rect[k].z && fy1 >= rp.cull_rect[k].y && fy0 <= rp.cull_rect[k].w);
    return hit;
}

constexpr int kOrderBlock = 1024;

// does the launch render this item at all (tile stride), and does it touch the scene
__device__ __forceinline__ void classify_item(const RenderParams& rp, int item, bool& mine, bool& scene)
{
    mine = item < rp.order_domain_items;
    if (mine && rp.tile_mod > 1) mine = ((item / (rp.groups_per_tile >> 2)) % rp.tile_mod) == rp.tile_rem;  // items never straddle tiles here
    scene = mine && rp.num_cull_rects > 0 && item_touches_scene(rp, item);
}

// per block: scene items in the low half, all items of the launch in the high half
__global__ void __launch_bounds__(kOrderBlock) count_scene_items_kernel(const __grid_constant__ RenderParams rp, int* __restrict__ block_counts)
{
    bool mine, scene;
    classify_item(rp, blockIdx.x * kOrderBlock + threadIdx.x, mine, scene);
    const int ns = __syncthreads_count(scene), nm = __syncthreads_count(mine);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = ns | (nm << 16);
}

__global__ void __launch_bounds__(kOrderBlock) place_items_kernel(const __grid_constant__ RenderParams rp, const int* __restrict__ block_counts,
                                                                  int* __restrict__ order)
{
    __shared__ int warp_a[kOrderBlock / 32], warp_b[kOrderBlock / 32], warp_c[kOrderBlock / 32];
    __shared__ int s_scene_before, s_sky_before, s_scene_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // scene / sky items of the blocks before this one, and the scene items of the whole launch
    int sb = 0, kb = 0, st = 0;
    for (int b = tid; b < (int)gridDim.x; b += kOrderBlock) {
        const int v = block_counts[b], ns = v & 0xffff, nm = v >> 16;
        st += ns;
        if (b < (int)blockIdx.x) { sb += ns; kb += nm - ns; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        kb += __shfl_xor_sync(0xffffffffu, kb, o);
        st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if (lane == 0) { warp_a[warp] = sb; warp_b[warp] = kb; warp_c[warp] = st; }
    __syncthreads();
    if (tid == 0) {
        int a = 0, b = 0, c = 0;
        for (int w = 0; w < kOrderBlock / 32; w++) { a += warp_a[w]; b += warp_b[w]; c += warp_c[w]; }
        s_scene_before = a; s_sky_before = b; s_scene_total = c;
    }
    __syncthreads();
    const int item = blockIdx.x * kOrderBlock + tid;
    bool mine, scene;
    classify_item(rp, item, mine, scene);
    const unsigned bt = __ballot_sync(0xffffffffu, scene), bs = __ballot_sync(0xffffffffu, mine && !scene);
    __syncthreads();
    if (lane == 0) warp_a[warp] = __popc(bt) | (__popc(bs) << 16);
    __syncthreads();
    int before_t = 0, before_s = 0;
    for (int w = 0; w < warp; w++) {
        before_t += warp_a[w] & 0xffff;
        before_s += warp_a[w] >> 16;
    }
    const unsigned lt = (1u << lane) - 1u;
    if (scene) order[s_scene_before + before_t + __popc(bt & lt)] = item;
    else if (mine) order[s_scene_total + s_sky_before + before_s + __popc(bs & lt)] = item;
}

// `order` holds num_items ints followed by ceil(order_domain_items / 1024) ints of scratch
cudaError_t launch_build_item_order(const RenderParams& rp, int* order, cudaStream_t stream)
{
    const int blocks = (rp.order_domain_items + kOrderBlock - 1) / kOrderBlock;
    int* counts = order + rp.num_items;
    count_scene_items_kernel<<<blocks, kOrderBlock, 0, stream>>>(rp, counts);
    place_items_kernel<<<blocks, kOrderBlock, 0, stream>>>(rp, counts, order);
    return cudaGetLastError();
}

// One rank's tiles of a tile-strided render (FlatTileIndex % mod == rem) copied into another buffer -- rank 0's, over
// NVLink peer memory: the gather of b200pt_group's tile sharding in one launch per rank.  16 B in, 16 B out per float4.
__global__ void tile_gather_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int num_tiles, int mod, int rem, int quads_per_tile)
{
    const int my_tiles = (num_tiles - rem + mod - 1) / mod;
    const size_t total = (size_t)my_tiles * quads_per_tile, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const size_t j = k / quads_per_tile, q = k - j * quads_per_tile;
        const size_t i = ((size_t)rem + j * mod) * quads_per_tile + q;
        dst[i] = __ldg(src + i);
    }
}

cudaError_t launch_tile_gather(const float* src, float* dst, int num_tiles, int mod, int rem, size_t floats_per_tile, int sm_count, cudaStream_t stream)
{
    if (floats_per_tile % 4) return cudaErrorInvalidValue;
    tile_gather_kernel<<<sm_count * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), num_tiles, mod, rem,
                                                         (int)(floats_per_tile / 4));
    return cudaGetLastError();
}

// FFMA micro-benchmark: the measured FP32-pipe peak the roofline of K1 is quoted against (SURVEY.md 8d asks for the real
// sustained FP32 rate instead of lanes x nominal clock).  16 independent fused multiply-add chains per thread, 4 CTAs of 256
// threads per SM, `iters` trips of 16 x 16 FFMA: 2 flop each.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* __restrict__ out, int iters, float seed)
{
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = seed + (float)(threadIdx.x + k);
    const float m = 1.0000001f, c = seed * 1e-9f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = __fmaf_rn(a[k], m, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) s += a[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

cudaError_t launch_ffma_peak(float* scratch, int blocks, int iters, cudaStream_t stream)
{
    ffma_peak_kernel<<<blocks, 256, 0, stream>>>(scratch, iters, 1.0f);
    return cudaGetLastError();
}

cudaError_t launch_resolve_ldr(const float* target, uint32_t* out, int width, int height, int tile_w, int tile_h,
                               int num_tiles_x, int mode, cudaStream_t stream)
{
    const long long npix = (long long)width * height;
    const int block = 256;
    const unsigned grid = (unsigned)((npix + block - 1) / block);
    resolve_ldr_kernel<<<grid, block, 0, stream>>>(target, out, width, height, tile_w, tile_h, num_tiles_x, mode);
    return cudaGetLastError();
}

cudaError_t launch_scale(float* target, size_t n, float scale, cudaStream_t stream)
{
    scale_kernel<<<148 * 8, 256, 0, stream>>>(target, n, scale);
    return cudaGetLastError();
}

cudaError_t launch_pack_env(const float* rgb, float4* rgba, size_t texels, cudaStream_t stream)
{
    pack_env_kernel<<<148 * 4, 256, 0, stream>>>(rgb, rgba, texels);
    return cudaGetLastError();
}

}  // namespace b200pt
