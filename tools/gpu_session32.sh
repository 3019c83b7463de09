#!/bin/bash
# where do the nf=32 single-pass forward epilogues wait?  ncu source counters of the four hidden-layer kernels
O=gpurun_out/s32; mkdir -p $O /tmp/ncu
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section InstructionStats --section SchedulerStats --clock-control none --import-source on -k regex:"tc_layer" --launch-skip 8 -c 4 -f -o /tmp/ncu/nf32 python tools/breakdown.py fp16 32 128 32 200000 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/ncu/nf32.ncu-rep "" $O/nf32_summary.json > $O/nf32_summary.txt 2>&1
ncu -i /tmp/ncu/nf32.ncu-rep --page source --csv --print-source sass > /tmp/ncu/nf32.src.csv 2>/dev/null
for k in 0 1 2 3; do python tools/ncu_top_stalls.py $k < /tmp/ncu/nf32.src.csv > $O/stalls_$k.txt 2>&1; done
gzip -c /tmp/ncu/nf32.src.csv > $O/nf32.src.csv.gz
ls -la $O
