// tc_bwd_b_pair: K = 1..5 (see tc_bwd_b_pair.inc)
#define STPDE_KC_HALF 0
#include "tc_bwd_b_pair.inc"
