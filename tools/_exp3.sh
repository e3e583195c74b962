set -e
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv
RSDET_TUNING=1 python -m rs_detection_b200.build --force > /dev/null
python tools/roi_sweep.py --paths 9,1 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/exp3_launches.csv python tools/roi_sweep.py --paths 9,1 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/exp3_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    d[r[h.index('Kernel Name')][:60]].append(float(r[h.index('Metric Value')]))
for k,v in d.items(): print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v)/1000:8.1f} us")
PY
