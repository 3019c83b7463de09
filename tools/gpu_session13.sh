#!/bin/bash
O=gpurun_out/s13; mkdir -p $O /tmp/ncu
export STPDE_LIB_PATH=$PWD/space_time_pde_b200/libstpde_t.so
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:tc_layer_pair_kernel --launch-skip 10 -c 3 -f -o /tmp/ncu/dgrad python tools/profile_bwd.py fp16x3 16384 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/ncu/dgrad.ncu-rep "" $O/dgrad_summary.json > $O/dgrad_summary.txt 2>&1
ncu -i /tmp/ncu/dgrad.ncu-rep --page source --csv --print-source sass > /tmp/ncu/dgrad.src.csv 2>/dev/null
python tools/sass_profile.py /tmp/ncu/dgrad.src.csv 40 2 > $O/dgrad_mix.txt 2>&1
python - <<'PY' > gpurun_out/s13/dgrad_hot.txt 2>&1
import csv
rows=list(csv.reader(open('/tmp/ncu/dgrad.src.csv')))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]+[len(rows)]
blk=rows[starts[2]:starts[3]]
hdr=blk[1]; ix={h:i for i,h in enumerate(hdr)}
recs=[]
for r in blk[2:]:
    if len(r)<len(hdr): continue
    recs.append((int(r[ix['# Samples']] or 0), int(r[ix['stall_long_sb']] or 0), int(r[ix['stall_lg']] or 0), int(r[ix['stall_short_sb']] or 0), int(r[ix['stall_wait']] or 0), int(r[ix['stall_membar']] or 0), int(r[ix['stall_mio']] or 0), r[ix['Source']].strip()[:90]))
print("total samples", sum(x[0] for x in recs))
print("(samples, long_sb, lg, short_sb, wait, membar, mio, instr)")
for x in sorted(recs,key=lambda t:-t[0])[:40]: print(x)
PY
du -sh $O
