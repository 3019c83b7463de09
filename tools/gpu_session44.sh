#!/bin/bash
# final bench lines (default bench, reference arm) + the graph-replay test
O=gpurun_out/s44; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x -k "cuda_graph or setup_cache or single_pass_stash" 2>&1 | tail -3
T0=$(date +%s)
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s44/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'step_frac',d['roofline']['step_frac'],'frac',d['roofline']['frac'],'clk',d['clocks'])
t=d['train_step']; print('train',t['value'],t['ms_per_step'],'eager',t.get('eager_chunk_loop_ms_per_step'),t.get('cuda_graph_chunks'))
print('small',t.get('reference_size_step'))
for k,v in d['configs'].items():
    print(k,v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),(v.get('roofline_hbm') or {}).get('frac'),v.get('cuda_graph_chunks'))
r=json.loads(open('gpurun_out/s44/bench_ref.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline']['kind'],r['cpu_baseline']['cores'])
PY
