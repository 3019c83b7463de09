#!/bin/bash
# A/B on one box: main library (z planes loaded straight into registers) vs variant t (z chunks staged by cp.async)
O=gpurun_out/s25; mkdir -p $O
T=$PWD/space_time_pde_b200/libstpde_t.so
for lib in main t; do
  if [ $lib = t ]; then export STPDE_LIB_PATH=$T; else unset STPDE_LIB_PATH; fi
  echo "== lib $lib"
  for prec in fp16x3 fp16; do
    echo "nf128 $prec"; STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py $prec 65536 3 2>&1 | grep -E "^\{"
    echo "nf32 $prec"; NF=32 STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py $prec 262144 3 2>&1 | grep -E "^\{"
  done
done 2>&1 | tee $O/ab.log
export STPDE_LIB_PATH=$T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 140 --csv --log-file $O/small_launches.csv python tools/profile_small_step.py > /dev/null 2>&1
python - <<'PY' | tee $O/small_launches.txt
import csv
rows=[r for r in csv.reader(open('gpurun_out/s25/small_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
tot=0
for r in rows[1:]:
    v=float(r[iv].replace(',','')); 
    if r[iu]=='ns': v/=1000
    elif r[iu]=='ms': v*=1000
    tot+=v
    print("%8.1f us  %s" % (v, r[ik][:100]))
print("total us", tot, "launches", len(rows)-1)
PY
