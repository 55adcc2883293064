#!/usr/bin/env bash
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_gpu.log | tail -1)"; grep -E "^FAILED|^E  |worst" $out/pytest_gpu.log | head -20
grep -E "dense shared adaptation" -r $out/pytest_gpu.log | head -2
python scripts/nuts_loop_diag.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_nuts.csv python scripts/bench_nuts.py 65536 128 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_nuts.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; rows=rows[hdr+1:]
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[]
for r in rows:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    seq.append((r[ki].split('(')[0][:50],v))
idx=[i for i,(k,v) in enumerate(seq) if 'k_nuts_init' in k]
print([round(v,1) for k,v in seq[idx[-1]:idx[-1]+9]])
PY
