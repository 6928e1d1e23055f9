#!/bin/bash
# ncu launch list (per-launch device time, cold-cache & serialised: compare SHARES) of one bench step.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-600
python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/launches.csv') as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = row['Kernel Name'][:70]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit in ('ns', 'nsecond'): v /= 1e3
    elif unit in ('ms', 'msecond'): v *= 1e3
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% x{c:<4d} {n}")
PY
