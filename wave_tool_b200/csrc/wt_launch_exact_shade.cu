// surface_warp_kernel instantiations: MODE 0, SHADE true (see wt_launch.cuh)
#include "wt_launch.cuh"

namespace wt {
cudaError_t launch_production_exact_shade(wt_ctx* ctx, EvalArgs& a, unsigned, cudaStream_t stream) {
    return launch_production<0, true>(ctx, a, stream);
}
}  // namespace wt
