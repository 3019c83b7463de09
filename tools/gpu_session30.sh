#!/bin/bash
# blend_backward rewrite (register accumulators, RB2 specialisation) + deferred status checks in the chunk loops
O=gpurun_out/s30; mkdir -p $O
export STPDE_PARITY_REPORT=$PWD/$O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
unset STPDE_PARITY_REPORT
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
timeout 300 python tools/train_chunk_probe.py 8192 40960 2>&1 | tail -2 | tee $O/probe.log
STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16x3 65536 3 2>&1 | grep -E "^\{" | tee -a $O/probe.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s30/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'step_frac',d['roofline']['step_frac'],'clk',d['clocks'])
print('kernel_ms',d['roofline']['kernel_ms'])
t=d['train_step']; print('train',t['value'],t['ms_per_step'],t['kernel_ms_per_step'])
print('small',t.get('reference_size_step'))
for k,v in d['configs'].items():
    print(k,v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),v.get('kernel_ms_per_step'))
PY
