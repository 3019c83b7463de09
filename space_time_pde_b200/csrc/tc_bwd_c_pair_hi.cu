// tc_bwd_c_pair: K = 6..10, Rayleigh-Benard specialisation, dispatcher (see tc_bwd_c_pair.inc)
#define STPDE_KC_HALF 1
#include "tc_bwd_c_pair.inc"
