#!/bin/bash
mkdir -p gpurun_out/s4
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 300 python tools/quick_parity.py 2>&1 | tail -14 | tee gpurun_out/s4/parity.log
for prec in fp16 fp16x3; do
  timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s4/breakdown.log
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 2 -f -o gpurun_out/s4/nf32_fp16 python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s4/ncu.log 2>&1; echo "ncu rc=$?"
du -sh gpurun_out/s4
