#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/pytest_tc.log
