#!/bin/bash
O=gpurun_out/s12; mkdir -p $O
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "rb2_spec_smooth or stash_is_reused or fused_vs_torch or kinked or tiny_cotangents or multi_chunk or fused_loss or swish_beta or encoder_gradients or chunked_training" 2>&1 | tail -12 | tee $O/pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --legs config2_train,config3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 300 $O/bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/s12/bench.json') if x.startswith('{')][-1]
d=json.loads(l)
t=d['configs']['config2_train']
print('train', t.get('ms_per_step'), {k:round(v,1) for k,v in t.get('kernel_ms_per_step',{}).items()})
print('config3', d['configs']['config3'].get('ms_per_step'), d['configs']['config3'].get('error'))
PY
