#!/bin/bash
# variant t: z staging + GEMM-style vertex kernels.  Tests, A/B against the main library, launch list of the small step,
# stall hot-list of the reverse kernels
O=gpurun_out/s26; mkdir -p $O /tmp/ncu
T=$PWD/space_time_pde_b200/libstpde_t.so
STPDE_LIB_PATH=$T timeout 900 python -m pytest tests/test_gpu_backward.py -q -x -k "rb2_spec_smooth or stash_is_reused or fused_vs_torch or kinked or tiny_cotangents or multi_chunk or fused_loss or swish_beta or encoder_gradients or chunked_training or cuda_graph or golden_gradients and rb2" 2>&1 | tail -5 | tee $O/pytest.log
STPDE_LIB_PATH=$T timeout 600 python tools/quick_parity.py 2>&1 | tail -12 | tee $O/parity.log
for lib in main t; do
  if [ $lib = t ]; then export STPDE_LIB_PATH=$T; else unset STPDE_LIB_PATH; fi
  echo "== lib $lib"
  echo "nf128 fp16x3"; STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16x3 65536 3 2>&1 | grep -E "^\{"
  echo "nf32 fp16"; NF=32 STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16 262144 3 2>&1 | grep -E "^\{"
  timeout 300 python tools/profile_small_step.py 2>&1 | head -2 | cut -c1-900
done 2>&1 | tee $O/ab.log
export STPDE_LIB_PATH=$T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 140 --csv --log-file $O/small_launches.csv python tools/profile_small_step.py > /dev/null 2>&1
python - <<'PY' | tee $O/small_launches.txt
import csv
rows=[r for r in csv.reader(open('gpurun_out/s26/small_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
for r in rows[1:]:
    v=float(r[iv].replace(',',''))
    if r[iu]=='ns': v/=1000
    elif r[iu]=='ms': v*=1000
    if v > 20: print("%8.1f us  %s" % (v, r[ik][:100]))
PY
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none --import-source on -k regex:"tc_layer" --launch-skip 16 -c 4 -f -o /tmp/ncu/dgrad python tools/profile_bwd.py fp16x3 16384 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/ncu/dgrad.ncu-rep "" $O/dgrad_summary.json > $O/dgrad_summary.txt 2>&1
ncu -i /tmp/ncu/dgrad.ncu-rep --page source --csv --print-source sass > /tmp/ncu/dgrad.src.csv 2>/dev/null
python - <<'PY' > $O/dgrad_hot.txt 2>&1
import csv
rows=list(csv.reader(open('/tmp/ncu/dgrad.src.csv')))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]+[len(rows)]
for b in range(len(starts)-1):
    blk=rows[starts[b]:starts[b+1]]
    print("==", blk[0][1][:80] if len(blk[0])>1 else blk[0])
    hdr=blk[1]; ix={h:i for i,h in enumerate(hdr)}
    recs=[]
    for r in blk[2:]:
        if len(r)<len(hdr): continue
        g=lambda k:int(r[ix[k]] or 0) if k in ix else 0
        recs.append((g('# Samples'), g('stall_long_sb'), g('stall_lg'), g('stall_short_sb'), g('stall_wait'), g('stall_membar'), g('stall_mio'), g('stall_barrier'), r[ix['Source']].strip()[:90]))
    print("total samples", sum(x[0] for x in recs))
    for x in sorted(recs,key=lambda t:-t[0])[:28]: print(x)
PY
du -sh $O
