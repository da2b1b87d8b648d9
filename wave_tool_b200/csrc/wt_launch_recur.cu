// surface_warp_kernel instantiations: MODE 1, SHADE false (see wt_launch.cuh)
#include "wt_launch.cuh"

namespace wt {
cudaError_t launch_production_recur(wt_ctx* ctx, EvalArgs& a, unsigned, cudaStream_t stream) {
    return launch_production<1, false>(ctx, a, stream);
}
cudaError_t launch_production_batch_recur(wt_ctx* ctx, std::vector<EvalArgs>& frames, cudaStream_t stream) {
    return launch_production_batch<1>(ctx, frames, stream);
}
}  // namespace wt
