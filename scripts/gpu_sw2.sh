#!/bin/bash
mkdir -p gpurun_out
DLWP_TC_DEBUG=1 timeout 120 python scripts/prof_tc.py --batch 256 --iters 3 2>&1 | sort | uniq -c | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_sw.csv python scripts/prof_tc.py --batch 256 --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/launches_sw.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
for nb in 1 2 4 8; do DLWP_TC_BANDS=$nb timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1; done
