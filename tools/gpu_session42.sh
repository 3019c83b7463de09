#!/bin/bash
# chunks of the config-3 training step replayed from CUDA graphs (bench leg): eager loop vs graphs
O=gpurun_out/s42; mkdir -p $O
timeout 900 python bench.py --steps 2 --warmup 1 --legs config3,config2_train --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -5 $O/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s42/bench.json').read().strip().splitlines()[-1])
for k in ('config3','config2_train'):
    c=d['configs'][k]; print(k, 'ms', c.get('ms_per_step'), 'eager', c.get('eager_chunk_loop_ms_per_step'), 'graph', c.get('cuda_graph_chunks'), 'value', c.get('value'), 'loss', c.get('loss_reg'), c.get('loss_pde'))
PY
