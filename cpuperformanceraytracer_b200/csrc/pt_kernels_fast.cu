// pt_kernels_fast.cu -- instantiates the megakernel for the FastMath policy.
// This translation unit is compiled with --fmad=true: free FMA contraction, MUFU approximations via the FastMath policy.
#include "pt_device.cuh"

namespace b200pt {

cudaError_t launch_render_fast(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return dispatch_config<FastMath>(lc, [&](auto kernel) -> cudaError_t {
        using KernelT = decltype(kernel);
        if constexpr (std::is_same<KernelT, void (*)(RenderParams, V4Scene)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v4);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v3redo);
        } else if constexpr (std::is_same<KernelT, void (*)(RenderParams, V3RedoScene0)>::value) {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v3redo0);
        } else {
            kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.cornell);
        }
        return cudaGetLastError();
    });
}

cudaError_t occupancy_fast(const LaunchConfig& lc, int* blocks_per_sm)
{
    return dispatch_config<FastMath>(lc, [&](auto kernel) -> cudaError_t {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, lc.block, 0);
    });
}

}  // namespace b200pt
