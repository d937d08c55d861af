#!/bin/bash
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench steps', d['steps'], 'ms/step', d['ms_per_step'], d['clocks'], 'checksum', d['config']['checksum_mean_abs_last_state'])"
