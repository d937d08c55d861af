#!/bin/bash
mkdir -p gpurun_out
python -m dlwp_b200.build >/dev/null 2>&1
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
echo "== ncu full tc kernels (N=64)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 0 -c 2 -o gpurun_out/prof_tc python scripts/prof_tc.py --batch 64 --iters 1 2>&1 | tail -2
