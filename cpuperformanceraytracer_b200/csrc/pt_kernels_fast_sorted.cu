// pt_kernels_fast_sorted.cu -- instantiates the CTA-sorted kernel (pt_wavefront.cuh) for the FastMath policy (--fmad=true).
#include "pt_wavefront.cuh"

namespace b200pt {

cudaError_t launch_render_sorted_fast(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return launch_sorted<FastMath>(lc, rp, scenes, stream);
}

cudaError_t occupancy_sorted_fast(const LaunchConfig& lc, int* blocks_per_sm) { return occupancy_sorted<FastMath>(lc, blocks_per_sm); }

}  // namespace b200pt
