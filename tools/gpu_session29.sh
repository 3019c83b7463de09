#!/bin/bash
# narrow-layer row-group packing (STPDE_PACK) + layer-0 rows per warp 32: tests, A/B timings
O=gpurun_out/s29; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
for pk in 0 1; do
  echo "== STPDE_PACK=$pk"
  STPDE_PACK=$pk timeout 300 python tools/breakdown.py fp16 32 128 32 1000000 2>&1 | tail -1
  STPDE_PACK=$pk timeout 300 python tools/breakdown.py fp16x3 32 128 32 1000000 2>&1 | tail -1
  STPDE_PACK=$pk timeout 300 python tools/breakdown.py fp16x3 32 128 32 10000 2>&1 | tail -1
  STPDE_PACK=$pk timeout 300 python tools/breakdown.py fp16x3 64 32 16 262144 2>&1 | tail -1
  STPDE_PACK=$pk timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2
  STPDE_PACK=$pk timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-700
done 2>&1 | tee $O/pack_ab.log
timeout 600 python tools/quick_parity.py 2>&1 | tail -12 | tee $O/parity.log
