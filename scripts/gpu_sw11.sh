#!/bin/bash
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_rollout_gpu.py tests/test_latband_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
for st in 50 200; do timeout 300 python bench.py --steps $st --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench steps', d['steps'], 'ms/step', d['ms_per_step'], d['clocks'])"; done
timeout 300 python scripts/bench_net_b.py --batch 16 --steps 20 --math tc --per-op 2>&1 | tail -1 | cut -c1-600
