#!/bin/bash
mkdir -p gpurun_out/s6
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section MemoryWorkloadAnalysis --section Occupancy --clock-control none --import-source on -k regex:"layer0_jets|final_blend" --launch-skip 8 -c 2 -f -o gpurun_out/s6/nf32_l0_fb_fp16 python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s6/ncu3.log 2>&1; echo "ncu rc=$?"
du -sh gpurun_out/s6
