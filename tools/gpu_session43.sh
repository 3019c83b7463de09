#!/bin/bash
# final validation after the CUDA-graph chunk replay in the training legs: GPU tests, smoke, default bench, reference arm
O=gpurun_out/s43; mkdir -p $O
export STPDE_PARITY_REPORT=$PWD/$O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
unset STPDE_PARITY_REPORT
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
/usr/bin/time -v timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s43/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'step_frac',d['roofline']['step_frac'],'frac',d['roofline']['frac'],'clk',d['clocks'])
t=d['train_step']; print('train',t['value'],t['ms_per_step'],'eager',t.get('eager_chunk_loop_ms_per_step'),t.get('cuda_graph_chunks'))
print('small',t.get('reference_size_step'))
for k,v in d['configs'].items():
    print(k,v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),(v.get('roofline_hbm') or {}).get('frac'),v.get('cuda_graph_chunks'))
r=json.loads(open('gpurun_out/s43/bench_ref.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline']['kind'],r['cpu_baseline']['cores'])
PY
