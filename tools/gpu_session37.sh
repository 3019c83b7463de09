#!/bin/bash
# fp16 pre-activation planes in the single-pass training mode (STPDE_Z_HALF): tests, A/B, gradient parity of the fp16 mode
O=gpurun_out/s37; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
for zh in 0 1 0 1; do
  echo "== STPDE_Z_HALF=$zh"
  STPDE_Z_HALF=$zh timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2
done 2>&1 | tee $O/zhalf_ab.log
for zh in 0 1; do
STPDE_Z_HALF=$zh timeout 600 python - <<'PY' 2>&1 | tail -4 | tee -a $O/zhalf_ab.log
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from tests.test_gpu_backward import run_case, RB2
dev = torch.device("cuda:0")
for nf, seed in ((8, 3), (32, 6)):
    errs = run_case(dev, 3, (4, 6, 6), 32, 4, nf, "softplus", *RB2, p=2048, precision="fp16", seed=seed)
    print("Z_HALF", os.environ.get("STPDE_Z_HALF"), "nf", nf, "fp16-mode gradient errors vs fp64 autograd:", {k: f"{v:.2e}" for k, v in errs.items()})
PY
done
