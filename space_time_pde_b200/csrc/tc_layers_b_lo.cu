// tc_layers_b: K = 1..5 (see tc_layers_b.inc)
#define STPDE_KC_HALF 0
#include "tc_layers_b.inc"
