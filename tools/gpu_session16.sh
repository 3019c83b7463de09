#!/bin/bash
O=gpurun_out/s16; mkdir -p $O
timeout 900 python bench.py --steps 2 --warmup 3 --legs config3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 300 $O/bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/s16/bench.json') if x.startswith('{')][-1]
d=json.loads(l)
c=d['configs']['config3']
print('config3', c.get('ms_per_step'), c.get('error'), c.get('gpu_launches_per_step'))
print(c.get('kernel_ms_per_step'))
PY
