#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/c_pytest.log | tail -15
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/c_bench50.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/c_bench50.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks']); print([ (r['kernel'],round(r['ms_per_launch'],4)) for r in d['roofline']['layers']])
else:
    print(open('gpurun_out/c_bench50.log').read()[-2000:])
PY
