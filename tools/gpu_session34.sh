#!/bin/bash
# 2 x B200: default bench under torchrun (weak-scaling headline, strong-scaling config 3 / 5 legs, NCCL all-reduce)
O=gpurun_out/s34; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s34/bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'n',d['n_gpus'],'ms',d['ms_per_step'],'clk',d['clocks'])
t=d['train_step']; print('train',t['value'],t['ms_per_step'])
for k,v in d['configs'].items():
    print(k,v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),(v.get('roofline_hbm') or {}).get('frac'))
    if k=='config5':
        for s in v['sweep']: print('   ',s['total_points'],s['ms_per_step'],s['value'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"; tail -c 300 $O/bench_ref_n2.json
