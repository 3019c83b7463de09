#!/bin/bash
O=gpurun_out/s17; mkdir -p $O
for ch in 32768 131072 524288; do
  timeout 600 python bench.py --steps 2 --warmup 3 --legs config3 --no-cpu-baseline --config3-chunk $ch > $O/bench_$ch.json 2> $O/bench_$ch.err; echo "chunk $ch rc=$?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/s17/bench_$ch.json') if x.startswith('{')][-1]
c=json.loads(l)['configs']['config3']
print('chunk', $ch, 'ms', c.get('ms_per_step'), c.get('error'), c.get('gpu_launches_per_step'))
print(c.get('kernel_ms_per_step'))
PY
done
# single crop, same decoder, inference breakdown for comparison (per 1 M points)
timeout 300 python tools/breakdown.py fp16 32 32 16 1000000 2>&1 | tail -1
