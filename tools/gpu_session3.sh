#!/bin/bash
# epilogue variants: drain (LDS/STG), tma1 (TMA store, 1 buffer), tma2 (TMA store, 2 half-size buffers)
mkdir -p gpurun_out/s3
for v in drain tma1 tma2; do
  export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_$v.so
  for prec in fp16 fp16x3; do
    echo "== $v $prec"
    timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
    timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/s3/breakdown.log
for v in drain tma1; do
  export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_$v.so
  timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 2 -f -o gpurun_out/s3/nf32_fp16_$v python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s3/ncu_$v.log 2>&1; echo "ncu rc=$?"
done
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_drain.so
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"layer0_jets|final_blend" --launch-skip 8 -c 2 -f -o gpurun_out/s3/nf32_l0_fb python tools/breakdown.py fp16x3 32 128 32 262144 > gpurun_out/s3/ncu3.log 2>&1; echo "ncu rc=$?"
du -sh gpurun_out/s3; ls -la gpurun_out/s3
