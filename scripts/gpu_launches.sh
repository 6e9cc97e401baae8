#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'score_kernel|exact_rows|bands_kernel|compact_|fold_redo' -s 24 -c 36 --csv --log-file gpurun_out/launches_iter.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_iter.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_iter.csv')) if len(r)>5]
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    k=d['Kernel Name'][:60]; v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,a in agg.items(): print(f"{a[0]:4d} x {a[1]/a[0]:10.1f} us  {k}")
PY
