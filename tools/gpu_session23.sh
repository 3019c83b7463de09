#!/bin/bash
# CUDA-graph replay of the training step, merged zero/scale launches, refreshed ncu summary of the reverse-mode kernels
O=gpurun_out/s23; mkdir -p $O /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x 2>&1 | tail -6 | tee $O/pytest.log
timeout 600 python - <<'PY' 2>&1 | tee $O/small_step.log
import json, torch, bench
out = bench.reference_size_train_step(torch.device("cuda:0"), with_eager=False)
print(json.dumps(out))
PY
timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-900 | tee -a $O/small_step.log
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"tc_layer|tc_wgrad" --launch-skip 12 -c 12 -f -o /tmp/ncu/train python tools/profile_bwd.py fp16x3 16384 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/ncu/train.ncu-rep "" $O/train_summary.json > $O/train_summary.txt 2>&1
ncu -i /tmp/ncu/train.ncu-rep --page source --csv --print-source sass > /tmp/ncu/train.src.csv 2>/dev/null
python tools/sass_profile.py /tmp/ncu/train.src.csv 24 12 > $O/train_mix.txt 2>&1
python - <<'PY' > $O/train_hot.txt 2>&1
import csv
rows=list(csv.reader(open('/tmp/ncu/train.src.csv')))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]+[len(rows)]
for b in range(len(starts)-1):
    blk=rows[starts[b]:starts[b+1]]
    print("==", blk[0][1][:80] if len(blk[0])>1 else blk[0])
    hdr=blk[1]; ix={h:i for i,h in enumerate(hdr)}
    recs=[]
    for r in blk[2:]:
        if len(r)<len(hdr): continue
        g=lambda k:int(r[ix[k]] or 0) if k in ix else 0
        recs.append((g('# Samples'), g('stall_long_sb'), g('stall_lg'), g('stall_short_sb'), g('stall_wait'), g('stall_membar'), g('stall_mio'), r[ix['Source']].strip()[:90]))
    print("total samples", sum(x[0] for x in recs))
    for x in sorted(recs,key=lambda t:-t[0])[:22]: print(x)
PY
du -sh $O
