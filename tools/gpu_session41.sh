#!/bin/bash
# the documented opt-out switches still give a green suite (one row group per tile, no PDL, fp32 pre-activation planes,
# round-1 wgrad order, 8 rows per warp in layer 0, no setup cache)
O=gpurun_out/s41; mkdir -p $O
STPDE_PACK=0 STPDE_PDL=0 STPDE_Z_HALF=0 STPDE_WGRAD_ORDER=0 STPDE_L0_RPW=8 STPDE_SETUP_CACHE=0 timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_configs.py::test_row_group_packing_is_bitwise_neutral --deselect tests/test_gpu_configs.py::test_setup_cache_for_inference_loops > $O/pytest_switches_off.log 2>&1; echo "pytest (switches off) rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest_switches_off.log | tail -8
