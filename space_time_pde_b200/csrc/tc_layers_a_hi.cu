// tc_layers_a: K = 6..10, Rayleigh-Benard specialisation, dispatcher (see tc_layers_a.inc)
#define STPDE_KC_HALF 1
#include "tc_layers_a.inc"
