// pt_kernels_parity_sorted.cu -- instantiates the CTA-sorted kernel (pt_wavefront.cuh) for the ParityMath policy.
// Compiled like pt_kernels_parity.cu: --fmad=false -prec-div=true -prec-sqrt=true -ftz=false.
#include "pt_wavefront.cuh"

namespace b200pt {

cudaError_t launch_render_sorted_parity(const LaunchConfig& lc, const RenderParams& rp, const SceneSet& scenes, cudaStream_t stream)
{
    return launch_sorted<ParityMath>(lc, rp, scenes, stream);
}

cudaError_t occupancy_sorted_parity(const LaunchConfig& lc, int* blocks_per_sm) { return occupancy_sorted<ParityMath>(lc, blocks_per_sm); }

}  // namespace b200pt
