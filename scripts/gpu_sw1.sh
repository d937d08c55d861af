#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
DLWP_TC_GENERIC=1 timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
