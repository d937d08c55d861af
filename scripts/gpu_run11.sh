#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_rollout_gpu.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_TC_TAPS_IN_K=0 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_TC_ROUT=4 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_TC_ROUT=3 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-200
