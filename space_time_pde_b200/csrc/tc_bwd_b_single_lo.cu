// tc_bwd_b_single: K = 1..5 (see tc_bwd_b_single.inc)
#define STPDE_KC_HALF 0
#include "tc_bwd_b_single.inc"
