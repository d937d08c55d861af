#!/bin/bash
for nb in 1 2 4; do DLWP_TC_BANDS=$nb timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1; done
for nb in 2 4; do DLWP_TC_BANDS=$nb timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bands $nb bench steps', d['steps'], 'ms/step', d['ms_per_step'], d['clocks'])"; done
