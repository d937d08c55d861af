#!/bin/bash
# round 2: recurrent block tests first (new), then the full GPU pass, bench, issuing-warp timing with the balanced schedule
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_recurrent_gpu.py -q -x > gpurun_out/g_recurrent.log 2>&1
echo "exit $?" >> gpurun_out/g_recurrent.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/g_recurrent.log | head -30
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/g_pytest.log | tail -15
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/g_bench50.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/g_bench50.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks']); print([ (r['kernel'],round(r['ms_per_launch'],4)) for r in d['roofline']['layers']])
else:
    print(open('gpurun_out/g_bench50.log').read()[-2000:])
PY
for b in 256; do for dbg in 0 4; do timeout 120 python scripts/prof_tc.py --batch $b --opt tc_debug=$dbg 2>&1 | tail -2; done; done > gpurun_out/g_triage.txt 2>&1
for mr in 4 16 32; do timeout 120 python scripts/prof_tc.py --batch 256 --opt tc_bands=$mr 2>&1 | tail -1; done >> gpurun_out/g_triage.txt 2>&1
for b in 1 8 32; do timeout 120 python scripts/prof_tc.py --batch $b 2>&1 | tail -1; done >> gpurun_out/g_triage.txt 2>&1
cat gpurun_out/g_triage.txt
