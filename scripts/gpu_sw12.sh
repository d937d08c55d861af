#!/bin/bash
mkdir -p gpurun_out
DLWP_SW_DEBUG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_sw_kernel -s 2 -c 1 -o gpurun_out/prof_sw_dbg2 python scripts/prof_tc.py --batch 256 --iters 1 2>&1 | tail -1
