#!/bin/bash
# per-kernel durations of the call-by-call chain on a small bank (1,024 streams x 42 calls)
set -e
mkdir -p gpurun_out
python tools/small_bank_bench.py > gpurun_out/small_bank.log 2>&1
cat gpurun_out/small_bank.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/small_launches.csv \
    python tools/small_bank_bench.py 1024 > gpurun_out/small_ncu.log 2>&1 || true
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/small_launches.csv') if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size'); bi = h.index('Block Size')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[(r[ki][:60], r[gi], r[bi])].append(float(r[vi].replace(',', '')))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[0]:60s} grid {k[1]:>14s} block {k[2]:>12s} n={len(v):4d} mean {sum(v)/len(v)/1e3:8.2f} us  min {min(v)/1e3:8.2f}")
PY
