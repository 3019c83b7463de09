#!/bin/bash
# compute-sanitizer on the final library: memcheck over every kernel family (packed layers, PDL launches, new
# blend_backward), racecheck restricted to the SIMT reverse kernels (shared-memory accumulators of blend_backward)
O=gpurun_out/s36; mkdir -p $O
for prec in fp16x3 fp16; do
  timeout 900 compute-sanitizer --tool memcheck --print-limit 50 python tools/racecheck_small.py $prec > $O/memcheck_$prec.txt 2>&1; echo "memcheck $prec rc=$?"
  grep -E "ERROR SUMMARY|^nf" $O/memcheck_$prec.txt
done
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 --kernel-regex kns=blend_backward python tools/racecheck_small.py fp16x3 > $O/racecheck_blend.txt 2>&1; echo "racecheck blend rc=$?"
grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_blend.txt | sort | uniq -c | head
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 --kernel-regex kns=vertex_backward python tools/racecheck_small.py fp16x3 > $O/racecheck_vertex.txt 2>&1; echo "racecheck vertex rc=$?"
grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_vertex.txt | sort | uniq -c | head
timeout 600 compute-sanitizer --tool memcheck --print-limit 50 python tools/pack_probe.py /tmp/pp.npz > $O/memcheck_pack_probe.txt 2>&1; echo "memcheck pack_probe rc=$?"
grep -E "ERROR SUMMARY|saved" $O/memcheck_pack_probe.txt
