#!/bin/bash
# baseline session of round 2: tests, breakdowns, ncu of the narrow-layer kernels (fp16, nf=32)
mkdir -p gpurun_out/s1
python -m pytest tests -m gpu -x -q > gpurun_out/s1/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/s1/pytest.log
for prec in fp16 fp16x3; do
  python tools/breakdown.py $prec 32 128 32 1000000 2>&1 | tail -1
  python tools/breakdown.py $prec 128 32 16 1000000 2>&1 | tail -1
done | tee gpurun_out/s1/breakdown.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer --launch-skip 16 -c 4 -f -o gpurun_out/s1/nf32_fp16 python tools/breakdown.py fp16 32 128 32 262144 > gpurun_out/s1/ncu1.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_layer_pair --launch-skip 16 -c 2 -f -o gpurun_out/s1/nf128_fp16 python tools/breakdown.py fp16 128 32 16 131072 > gpurun_out/s1/ncu2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/s1
