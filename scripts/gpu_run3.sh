#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== conv timing (N=256)"
for impl in ffma ffma_tma; do timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl $impl 2>&1 | tail -2; done
echo "== layer 2 tile sweep (TMA)"
for cfg in "16 3 100" "16 4 200" "16 8 227" "8 4 100" "8 8 200" "12 4 100" "12 8 200" "4 8 100" "16 2 100"; do
  set -- $cfg
  DLWP_TILE_TH=$1 DLWP_TILE_CC=$2 DLWP_TILE_SMEM_KB=$3 timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl ffma_tma --layer 2 2>&1 | tail -1
done
echo "== layer 1 tile sweep (TMA)"
for cfg in "4 4" "8 2" "2 4" "16 1" "8 1" "4 2"; do
  set -- $cfg
  DLWP_TILE_TH=$1 DLWP_TILE_NCG=$2 timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl ffma_tma --layer 1 2>&1 | tail -1
done
echo "== ncu full: conv2 tma"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 2 -c 1 -o gpurun_out/prof_conv2_tma python scripts/prof_conv.py --batch 64 --iters 1 --impl ffma_tma --layer 2 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 2 -c 1 -o gpurun_out/prof_conv1_tma python scripts/prof_conv.py --batch 64 --iters 1 --impl ffma_tma --layer 1 2>&1 | tail -1
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
