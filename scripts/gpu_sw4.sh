#!/bin/bash
for it in 10 100 400; do timeout 120 python scripts/prof_tc.py --batch 256 --iters $it 2>&1 | tail -1; done
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step', d['ms_per_step'], d['clocks'])"
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench(50) ms/step', d['ms_per_step'], d['clocks'])"
