#!/bin/bash
O=gpurun_out/s21; mkdir -p $O
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 300 python tools/quick_parity.py 2>&1 | tail -12 | tee $O/parity.log
for l0 in 1 0; do
  export STPDE_L0_TMA=$l0
  echo "== STPDE_L0_TMA=$l0"
  for prec in fp16 fp16x3; do
    timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
    timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
  done
done 2>&1 | tee $O/breakdown.log
