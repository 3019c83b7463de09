#!/bin/bash
# full build with the persistent-loop epilogue: tests (+ parity report), breakdowns, bench with all legs, L0/final_blend profile
mkdir -p gpurun_out/s5
export STPDE_PARITY_REPORT=$PWD/gpurun_out/s5/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s5/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/s5/pytest.log
unset STPDE_PARITY_REPORT
timeout 300 python tools/quick_parity.py 2>&1 | tail -14 | tee gpurun_out/s5/parity.log
for prec in fp16 fp16x3; do
  timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s5/breakdown.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s5/bench.json 2> gpurun_out/s5/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s5/bench.err
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 2 -f -o gpurun_out/s5/nf32_fp16 python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s5/ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"layer0_jets|final_blend" --launch-skip 8 -c 2 -f -o gpurun_out/s5/nf32_l0_fb python tools/breakdown.py fp16x3 32 128 32 262144 > gpurun_out/s5/ncu3.log 2>&1; echo "ncu rc=$?"
du -sh gpurun_out/s5
