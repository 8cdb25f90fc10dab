// pt_post.cu -- the small HBM-bound kernels around the megakernel.
//   resolve_ldr : OutputToScreen / OutputToFile (demofox_path_tracing_optimization_v4.cpp:1260-1331)
//                 with ACESFilm (:166-176), LinearToSRGB (:178-187) and fast_pow_gamma (:144-155),
//                 fast variants (global_preprocessor_flags.h:62-63), exact rcp like the oracle.
//   scale       : ACCUM_SUM epilogue after the cross-GPU sum (target *= 1/(N+1)).
//   pack_env    : RGB f32 rows -> RGBA32F texels for the env texture object.
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
    out[(size_t)y * width + x] = tonemap::pack(px[0], px[8], px[16], mode);
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
