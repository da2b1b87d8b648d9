// surface_warp_kernel instantiations: MODE 0, SHADE false (see wt_launch.cuh)
#include "wt_launch.cuh"

namespace wt {
cudaError_t launch_production_exact(wt_ctx* ctx, EvalArgs& a, unsigned, cudaStream_t stream) {
    return launch_production<0, false>(ctx, a, stream);
}
cudaError_t launch_production_batch_exact(wt_ctx* ctx, std::vector<EvalArgs>& frames, cudaStream_t stream) {
    return launch_production_batch<0>(ctx, frames, stream);
}
}  // namespace wt
