// tc_layers_a: K = 1..5 (see tc_layers_a.inc)
#define STPDE_KC_HALF 0
#include "tc_layers_a.inc"
