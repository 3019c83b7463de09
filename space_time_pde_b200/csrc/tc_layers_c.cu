// CTA-pair kernel with the fused layer-0 generator (opt-in, STPDE_TC_FUSE0=1), instantiations K = 1..10.
#define STPDE_TC_LAUNCH_IMPL
#include "tc_launch.cuh"

namespace stpde {
int tc_launch_layer_pair_gen(int kc, const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    int rc = STPDE_OK;
    STPDE_TC_DISPATCH_KC(kc, (rc = launch_layer_pair<KC, true>(tc, L, spec, a, st)));
    return rc;
}
}  // namespace stpde
