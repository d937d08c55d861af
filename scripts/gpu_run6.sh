#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_tc.log
echo "== bench tc"
timeout 900 python bench.py --steps 50 --warmup 3 --math tc 2>&1 | tail -1 | tee gpurun_out/bench_tc.log
echo "== bench ffma"
timeout 900 python bench.py --steps 50 --warmup 3 --math ffma --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_ffma.log
