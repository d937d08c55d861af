#!/bin/bash
# round 2: full GPU test pass + bench + issuing-warp timing of the two Net A kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/f_pytest.log | tail -15
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/f_bench50.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/f_bench50.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks']); print([ (r['kernel'],round(r['ms_per_launch'],4)) for r in d['roofline']['layers']])
else:
    print(open('gpurun_out/f_bench50.log').read()[-2000:])
PY
for b in 256 64; do for dbg in 4 5 6 7; do timeout 120 python scripts/prof_tc.py --batch $b --opt tc_debug=$dbg 2>&1 | tail -2; done; done > gpurun_out/f_triage.txt 2>&1
cat gpurun_out/f_triage.txt
