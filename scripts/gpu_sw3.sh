#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_sw_kernel -s 2 -c 2 -o gpurun_out/prof_sw_b256 python scripts/prof_tc.py --batch 256 --iters 1 2>&1 | tail -2
