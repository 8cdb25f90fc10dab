// pt_kernels_parity_v4sw.cu -- the scene-specialised OPT_V4 kernels with the reference's non-default shading switches compiled in
// (USE_FAST_APPROXIMATE_EXP 0 and / or USE_UNIT_VECTOR_REJECTION_SAMPLING 0, global_preprocessor_flags.h:64-65), ParityMath policy.
// A translation unit of its own (--fmad=false -prec-div=true -prec-sqrt=true -ftz=false, like pt_kernels_parity.cu) so that it builds next to the default kernels, not after them.
#include "pt_device.cuh"

namespace b200pt {

cudaError_t launch_render_parity_v4sw(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return dispatch_config_v4sw<ParityMath>(lc, [&](auto kernel) -> cudaError_t {
        kernel<<<lc.grid, lc.block, 0, stream>>>(rp, scenes.v4);
        return cudaGetLastError();
    });
}

cudaError_t occupancy_parity_v4sw(const LaunchConfig& lc, int* blocks_per_sm)
{
    return dispatch_config_v4sw<ParityMath>(lc, [&](auto kernel) -> cudaError_t {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, lc.block, 0);
    });
}

}  // namespace b200pt
