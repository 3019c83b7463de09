#!/bin/bash
# dgrad of narrow layers: 2 row groups per tile (block-diagonal W^T); tests + A/B
O=gpurun_out/s31; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
for pk in 0 1; do
  echo "== STPDE_PACK=$pk"
  STPDE_PACK=$pk timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2
  STPDE_PACK=$pk NF=16 STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16x3 65536 3 2>&1 | grep -E "^\{|^ok"
  STPDE_PACK=$pk timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-700
done 2>&1 | tee $O/pack_ab.log
