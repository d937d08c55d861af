#!/bin/bash
# where a training step spends its time: ncu launch list of scripts/bench_train.py (batch 4, 1 step after warm-up)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2000 --csv --log-file gpurun_out/r02_launches_train.csv python scripts/bench_train.py --batch 4 --steps 1 > gpurun_out/r02_ncu_train.log 2>&1
tail -2 gpurun_out/r02_ncu_train.log | cut -c1-300
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_train.csv')) if len(r)>10]
hdr=rows[0]; ni=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ni].split('(')[0][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%-62s %5d %10.1f us %5.1f%%' % (k,a[0],a[1]/1000,100*a[1]/tot))
PY
timeout 300 python scripts/bench_net_b.py --batch 64 --steps 20 --per-op 2>&1 | tail -1 | cut -c1-500
