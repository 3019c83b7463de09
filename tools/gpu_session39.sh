#!/bin/bash
# final validation of round 2: GPU tests, smoke, default bench, reference arm, ncu launch list of the bench command,
# per-kernel DRAM / tensor / issue figures of one config-3-shaped training chunk
O=gpurun_out/s39; mkdir -p $O
export STPDE_PARITY_REPORT=$PWD/$O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
unset STPDE_PARITY_REPORT
grep -E "^FAILED|passed|failed|Error" $O/pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s39/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'step_frac',d['roofline']['step_frac'],'frac',d['roofline']['frac'],'clk',d['clocks'])
print('kernel_ms',d['roofline']['kernel_ms'])
t=d['train_step']; print('train',t['value'],t['ms_per_step'])
print('small',t.get('reference_size_step'))
for k,v in d['configs'].items():
    print(k,v.get('value'),v.get('ms_per_step'),v.get('roofline',{}).get('frac'),(v.get('roofline_hbm') or {}).get('frac'))
r=json.loads(open('gpurun_out/s39/bench_ref.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['cpu_baseline']['kind'],r['cpu_baseline']['cores'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --legs '' --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
gzip -f $O/launches_bench.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
NF=32 timeout 900 ncu --metrics $M --clock-control none --launch-skip 110 -c 110 --csv --log-file $O/train_chunk_nf32_fp16.csv python tools/profile_bwd.py fp16 65536 2 > /dev/null 2>&1; echo "ncu chunk rc=$?"
python - <<'PY' | tee gpurun_out/s39/train_chunk_nf32_fp16.txt
import csv
rows=[r for r in csv.reader(open('gpurun_out/s39/train_chunk_nf32_fp16.csv')) if len(r)>5]
hdr=rows[0]; ii=hdr.index('ID'); ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
d={}; names={}
for r in rows[1:]:
    d.setdefault(r[ii],{})[r[im]]=(r[iv],r[iu]); names[r[ii]]=r[ik]
print("# one training step (forward-save + reverse sweep) of ImNet nf=32, single-pass fp16, 65536 points = 524288 rows, ncu per launch (cold, serialised)")
print("%-58s %9s %8s %8s %7s %7s %7s %6s" % ("kernel","us","rd MB","wr MB","dram%","tens%","issue%","L2hit"))
for k in sorted(d,key=int):
    m=d[k]; n=names[k]
    if not any(t in n for t in ('stpde','tc::')): continue
    def g(name,scale=1.0):
        v,u=m.get(name,('0',''))
        x=float(v.replace(',',''))
        if u=='ns': x/=1000
        elif u=='ms': x*=1000
        elif u=='s': x*=1e6
        elif u=='Kbyte': x/=1e3
        elif u=='byte': x/=1e6
        elif u=='Gbyte': x*=1e3
        return x*scale
    t=g('gpu__time_duration.sum')
    if t<15: continue
    print("%-58s %9.1f %8.1f %8.1f %7.1f %7.1f %7.1f %6.1f" % (n.replace('stpde::','').replace('void ','')[:58], t, g('dram__bytes_read.sum'), g('dram__bytes_write.sum'), g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), g('smsp__issue_active.avg.pct_of_peak_sustained_active'), g('lts__t_sector_hit_rate.pct')))
PY
