#!/bin/bash
O=gpurun_out/s27; mkdir -p $O
export STPDE_PARITY_REPORT=$PWD/$O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" $O/pytest.log | tail -10
unset STPDE_PARITY_REPORT
# racecheck restricted to the forward layer kernels (new smem staging / scratch protocol); barrier-word reports filtered
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 100000 --kernel-regex kns=tc_layer python tools/racecheck_small.py fp16x3 > /tmp/rc_fwd.txt 2>&1; echo "racecheck rc=$?"
grep -c "hazard detected" /tmp/rc_fwd.txt; grep "hazard detected" /tmp/rc_fwd.txt | sed 's/ at __shared__.*//' | sort | uniq -c > $O/racecheck_fwd_kinds.txt; cat $O/racecheck_fwd_kinds.txt
grep -A3 "hazard detected" /tmp/rc_fwd.txt | grep -v "CUDA barrier operation" | grep -A3 "hazard detected" | head -60 > $O/racecheck_fwd_non_barrier.txt
grep "RACECHECK SUMMARY" /tmp/rc_fwd.txt | tee -a $O/racecheck_fwd_kinds.txt
python tools/sweep.py fp16x3 2>&1 | tail -6 | tee $O/sweep_fp16x3.log
python tools/sweep.py fp16 2>&1 | tail -6 | tee $O/sweep_fp16.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; tail -c 300 $O/bench_ref.json
du -sh $O
for p in 16384 8192; do python tools/train_chunk_probe.py $p 40960 2>&1 | tail -2; done | tee $O/chunk_probe.log
