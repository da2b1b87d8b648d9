// surface_warp_kernel instantiations: MODE 1, SHADE true (see wt_launch.cuh)
#include "wt_launch.cuh"

namespace wt {
cudaError_t launch_production_recur_shade(wt_ctx* ctx, EvalArgs& a, unsigned, cudaStream_t stream) {
    return launch_production<1, true>(ctx, a, stream);
}
}  // namespace wt
