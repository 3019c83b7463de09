// CTA-pair kernel, training mode tc::kModeBwd0, instantiations K = 1..10.
#define STPDE_TC_LAUNCH_IMPL
#include "tc_bwd.h"
#include "tc_launch.cuh"

namespace stpde {
int tc_launch_pair_bwd0(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                        const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    int rc = STPDE_OK;
    if (spec_is_rb2(spec)) return launch_layer_pair_mode<6, tc::kModeBwd0, tc::kSpecRb2>(num_sms, w_hi, w_lo, a_hi, a_lo, spec, a, st);
    STPDE_TC_DISPATCH_KC(kc, (rc = launch_layer_pair_mode<KC, tc::kModeBwd0>(num_sms, w_hi, w_lo, a_hi, a_lo, spec, a, st)));
    return rc;
}

int tc_launch_single_bwd0(int kc, int num_sms, const CUtensorMap& w_hi, const CUtensorMap& w_lo, const CUtensorMap& a_hi,
                          const CUtensorMap& a_lo, const JetSpec& spec, const tc::LayerArgs& a, cudaStream_t st) {
    int rc = STPDE_OK;
    if (spec_is_rb2(spec)) return launch_layer_mode<6, tc::kModeBwd0, tc::kSpecRb2>(num_sms, w_hi, w_lo, a_hi, a_lo, spec, a, st);
    STPDE_TC_DISPATCH_KC(kc, (rc = launch_layer_mode<KC, tc::kModeBwd0>(num_sms, w_hi, w_lo, a_hi, a_lo, spec, a, st)));
    return rc;
}
}  // namespace stpde
