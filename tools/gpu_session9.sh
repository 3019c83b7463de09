#!/bin/bash
# evidence: launch list of the bench command, ncu full-set summaries, gather coalescing counters, racecheck, small-step profile
O=gpurun_out/s9; mkdir -p $O /tmp/ncu
# (a) launch list of the bench command (headline only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --legs '' --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
gzip -f $O/launches_bench.csv
# (b) full-set captures, summarised on the box (the .ncu-rep files stay in /tmp)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_layer|layer0_jets|prep_points|vertex_bias|residual" --launch-skip 40 -c 12 -f -o /tmp/ncu/cfg2_fp16x3 python tools/breakdown.py fp16x3 128 32 16 131072 > $O/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
python tools/ncu_summary.py /tmp/ncu/cfg2_fp16x3.ncu-rep "" $O/ncu_cfg2_fp16x3_summary.json > $O/ncu_cfg2_fp16x3_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_layer|layer0_jets" --launch-skip 20 -c 6 -f -o /tmp/ncu/cfg2_fp16 python tools/breakdown.py fp16 128 32 16 131072 > $O/ncu_cfg2_fp16.log 2>&1; echo "ncu cfg2 fp16 rc=$?"
python tools/ncu_summary.py /tmp/ncu/cfg2_fp16.ncu-rep "" $O/ncu_cfg2_fp16_summary.json > $O/ncu_cfg2_fp16_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_layer|layer0_jets" --launch-skip 20 -c 6 -f -o /tmp/ncu/nf32_fp16 python tools/breakdown.py fp16 32 128 32 262144 > $O/ncu_nf32.log 2>&1; echo "ncu nf32 rc=$?"
python tools/ncu_summary.py /tmp/ncu/nf32_fp16.ncu-rep "" $O/ncu_nf32_fp16_summary.json > $O/ncu_nf32_fp16_summary.txt 2>&1
for f in cfg2_fp16x3 nf32_fp16; do
  ncu -i /tmp/ncu/$f.ncu-rep --page source --csv --print-source sass > /tmp/ncu/$f.src.csv 2>/dev/null
  for b in 0 1 2 3 4 5; do python tools/sass_profile.py /tmp/ncu/$f.src.csv 16 $b; done > $O/sass_mix_$f.txt 2>&1
done
# training-step kernels (forward-save, wgrad, dgrad)
timeout 900 ncu --set full --clock-control none -k regex:"tc_layer|tc_wgrad|blend_backward" --launch-skip 30 -c 14 -f -o /tmp/ncu/train python tools/profile_bwd.py > $O/ncu_train.log 2>&1; echo "ncu train rc=$?"
python tools/ncu_summary.py /tmp/ncu/train.ncu-rep "" $O/ncu_train_summary.json > $O/ncu_train_summary.txt 2>&1
# (c) racecheck + memcheck on a tiny forward / training step
for prec in fp16x3 fp16; do
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_small.py $prec > $O/racecheck_$prec.txt 2>&1; echo "racecheck $prec rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python tools/racecheck_small.py fp16x3 > $O/memcheck.txt 2>&1; echo "memcheck rc=$?"
tail -3 $O/racecheck_fp16x3.txt $O/memcheck.txt
# (d) host profile of the reference-size training step
timeout 300 python tools/profile_small_step.py > $O/small_step.txt 2>&1; head -4 $O/small_step.txt
du -sh $O
