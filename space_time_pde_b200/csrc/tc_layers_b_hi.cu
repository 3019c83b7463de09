// tc_layers_b: K = 6..10, Rayleigh-Benard specialisation, dispatcher (see tc_layers_b.inc)
#define STPDE_KC_HALF 1
#include "tc_layers_b.inc"
