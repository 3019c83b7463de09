#!/bin/bash
mkdir -p gpurun_out/s7
timeout 600 python tools/debug_e2.py 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/s7/debug.log
export STPDE_PARITY_REPORT=$PWD/gpurun_out/s7/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s7/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/s7/pytest.log | tail -30
unset STPDE_PARITY_REPORT
for prec in fp16 fp16x3; do
  timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s7/breakdown.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s7/bench.json 2> gpurun_out/s7/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/s7/bench.err
du -sh gpurun_out/s7
