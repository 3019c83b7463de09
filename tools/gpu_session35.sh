#!/bin/bash
# programmatic dependent launch of the tensor-core kernels (STPDE_PDL): tests, A/B
O=gpurun_out/s35; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
for pdl in 0 1 0 1; do
  echo "== STPDE_PDL=$pdl"
  STPDE_PDL=$pdl timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2 | head -1
  STPDE_PDL=$pdl timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-200
  STPDE_PDL=$pdl timeout 300 python tools/sweep.py fp16x3 2>&1 | head -3 | cut -c1-120
done 2>&1 | tee $O/pdl_ab.log
for pdl in 0 1; do
STPDE_PDL=$pdl timeout 600 python - <<'PY' 2>&1 | tail -3 | tee -a $O/pdl_ab.log
import os, sys
sys.path.insert(0, os.getcwd())
import bench, torch
r = bench.reference_size_train_step(torch.device("cuda:0"), False)
print("PDL", os.environ.get("STPDE_PDL"), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k != "config"})
PY
done
