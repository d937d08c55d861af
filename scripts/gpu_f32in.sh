#!/bin/bash
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -6
DLWP_SW_F32IN=1 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_SW_F32IN=1 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('F32IN bench steps', d['steps'], 'ms/step', d['ms_per_step'], d['clocks'])"
DLWP_SW_F32IN=1 timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('F32IN bench steps', d['steps'], 'ms/step', d['ms_per_step'], d['clocks'])"
