// tc_bwd_a_single: K = 1..5 (see tc_bwd_a_single.inc)
#define STPDE_KC_HALF 0
#include "tc_bwd_a_single.inc"
