#!/bin/bash
# wgrad unit order (tile-fastest vs slice-fastest) and layer-0 rows per warp: tests, A/B timings, DRAM bytes of the wgrad kernel
O=gpurun_out/s28; mkdir -p $O /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" $O/pytest.log | tail -5
for wo in 0 1; do
  echo "== STPDE_WGRAD_ORDER=$wo nf128 fp16x3 65536"
  STPDE_WGRAD_ORDER=$wo STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16x3 65536 3 2>&1 | grep -E "^\{"
  echo "== STPDE_WGRAD_ORDER=$wo nf32 fp16 262144"
  STPDE_WGRAD_ORDER=$wo NF=32 STPDE_PRINT_PROFILE=1 timeout 300 python tools/profile_bwd.py fp16 262144 3 2>&1 | grep -E "^\{"
done 2>&1 | tee $O/wgrad_ab.log
for rpw in 8 16 32 64; do
  echo "== STPDE_L0_RPW=$rpw"
  STPDE_L0_RPW=$rpw timeout 300 python tools/breakdown.py fp16x3 128 32 16 262144 2>&1 | tail -1
  STPDE_L0_RPW=$rpw timeout 300 python tools/breakdown.py fp16 32 128 32 1000000 2>&1 | tail -1
  STPDE_L0_RPW=$rpw timeout 300 python tools/breakdown.py fp16x3 32 128 32 10000 2>&1 | tail -1
done 2>&1 | tee $O/l0_rpw.log
for wo in 0 1; do
  STPDE_WGRAD_ORDER=$wo timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tc_wgrad --launch-skip 8 -c 4 --csv --log-file $O/wgrad_order$wo.csv python tools/profile_bwd.py fp16x3 32768 3 > /dev/null 2>&1; echo "ncu rc=$?"
done
python - <<'PY' | tee $O/wgrad_dram.txt
import csv
for wo in (0,1):
    rows=[r for r in csv.reader(open(f'gpurun_out/s28/wgrad_order{wo}.csv')) if len(r)>5]
    hdr=rows[0]; ii=hdr.index('ID'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
    d={}
    for r in rows[1:]:
        d.setdefault(r[ii],{})[r[im]]=r[iv]+' '+r[iu]
    print('order',wo)
    for k,v in d.items(): print(' ',k,v)
PY
