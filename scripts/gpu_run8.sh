#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
for r in 8 2 1; do DLWP_TC_ROUT=$r timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1; done
timeout 120 python scripts/prof_tc.py --batch 256 --math ffma 2>&1 | tail -1
echo "== ncu full tc kernels (N=64)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 0 -c 2 -o gpurun_out/prof_tc2 python scripts/prof_tc.py --batch 64 --iters 1 2>&1 | tail -2
