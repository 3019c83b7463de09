#!/bin/bash
mkdir -p gpurun_out/s8
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 600 python tools/debug_e2.py 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/s8/debug.log
timeout 300 python tools/quick_parity.py 2>&1 | tail -12 | tee gpurun_out/s8/parity.log
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x -k "rb2_spec_smooth or stash_is_reused or fused_vs_torch" 2>&1 | tail -15 | tee gpurun_out/s8/pytest.log
for prec in fp16 fp16x3; do
  timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s8/breakdown.log
