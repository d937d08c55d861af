#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_tc.log
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_TC_ROUT=4 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
echo "== bench tc"
timeout 900 python bench.py --steps 50 --warmup 3 --math tc --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_tc.log | cut -c1-400
