#!/bin/bash
# setup cache across the chunks of a training step: backward tests + the two training legs of the bench
O=gpurun_out/s40; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_backward.py tests/test_reference_loop.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
for sc in 0 1; do
  echo "== STPDE_SETUP_CACHE=$sc"
  STPDE_SETUP_CACHE=$sc timeout 600 python bench.py --steps 2 --warmup 1 --legs config2_train,config3 --no-cpu-baseline > $O/bench_sc$sc.json 2> $O/bench_sc$sc.err
  python - $sc <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/s40/bench_sc{sys.argv[1]}.json').read().strip().splitlines()[-1])
t=d['train_step']; print('train', round(t['ms_per_step'],1), 'setup', round(t['kernel_ms_per_step']['setup'],2), 'launches', t['gpu_launches_per_step'])
c=d['configs']['config3']; print('config3', round(c['ms_per_step'],1), 'setup', c['kernel_ms_per_step']['setup'], 'launches', c['gpu_launches_per_step'], 'loss', c['loss_reg'], c['loss_pde'])
PY
done 2>&1 | tee $O/setup_cache_ab.log
