// tc_bwd_a_single: K = 6..10, Rayleigh-Benard specialisation, dispatcher (see tc_bwd_a_single.inc)
#define STPDE_KC_HALF 1
#include "tc_bwd_a_single.inc"
