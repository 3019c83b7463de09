// CTA-pair tensor-core layer kernel (cta_group::2, M = 256), instantiations K = 1..10.
#define STPDE_TC_LAUNCH_IMPL
#include "tc_launch.cuh"

namespace stpde {
int tc_launch_layer_pair(int kc, const TcContext& tc, const TcLayerPlan& L, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    int rc = STPDE_OK;
    if (spec_is_rb2(spec)) return launch_layer_pair<6, tc::kSpecRb2>(tc, L, spec, a, st);
    STPDE_TC_DISPATCH_KC(kc, (rc = launch_layer_pair<KC>(tc, L, spec, a, st)));
    return rc;
}
}  // namespace stpde
