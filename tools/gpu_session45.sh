#!/bin/bash
# 2 x B200: the two-phase (capture locally, then all ranks agree) graph path of the training legs
O=gpurun_out/s45; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --legs config2_train,config3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s45/bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], d['n_gpus'])
for k in ('config2_train','config3'):
    c=d['configs'][k]; print(k, c['ms_per_step'], c.get('eager_chunk_loop_ms_per_step'), c.get('cuda_graph_chunks'), c.get('loss_reg'), c.get('loss_pde'))
PY
