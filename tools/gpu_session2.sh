#!/bin/bash
# E1 (staged TMA-store forward epilogue): tests, breakdowns, ncu instruction mix
mkdir -p gpurun_out/s2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s2/pytest.log
for prec in fp16 fp16x3; do
  timeout 300 python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  timeout 300 python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done | tee gpurun_out/s2/breakdown.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 4 -f -o gpurun_out/s2/nf32_fp16 python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s2/ncu1.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 2 -f -o gpurun_out/s2/nf32_fp16x3 python tools/breakdown.py fp16x3 32 128 32 262144 > gpurun_out/s2/ncu2.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layer0_jets|final_blend" --launch-skip 8 -c 2 -f -o gpurun_out/s2/nf32_l0_fb python tools/breakdown.py fp16x3 32 128 32 262144 > gpurun_out/s2/ncu3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/s2
