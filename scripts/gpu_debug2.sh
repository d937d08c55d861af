#!/bin/bash
mkdir -p gpurun_out
echo "== TMA probe x=-4 (aligned) y=-2"; timeout 60 ./scripts/tma_probe 1 -4 -2 2>&1 | tail -2
echo "== TMA probe x=-2 y=0 (misaligned)"; timeout 60 ./scripts/tma_probe 0 -2 0 2>&1 | tail -2
echo "== FFMA probe"; timeout 120 ./scripts/ffma_probe 2>&1 | tee gpurun_out/ffma_probe.log
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== conv timing variants (N=256)"
for impl in ffma ffma_tma; do
  timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl $impl 2>&1 | tail -2
done
for cfg in "8 4 100" "8 8 200" "16 4 200" "16 2 100" "12 4 100" "4 8 100" "6 6 100"; do
  set -- $cfg
  DLWP_TILE_TH=$1 DLWP_TILE_CC=$2 DLWP_TILE_SMEM_KB=$3 timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl ffma --layer 2 2>&1 | tail -1
done
for cfg in "4 4" "8 4" "8 2" "2 4" "16 1" "8 1"; do
  set -- $cfg
  DLWP_TILE_TH=$1 DLWP_TILE_NCG=$2 timeout 120 python scripts/prof_conv.py --batch 256 --iters 10 --impl ffma --layer 1 2>&1 | tail -1
done
echo "== ncu full: conv2 ffma"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 2 -c 1 -o gpurun_out/prof_conv2_ffma python scripts/prof_conv.py --batch 64 --iters 1 --impl ffma --layer 2 2>&1 | tail -2
echo "== ncu full: conv1 ffma"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 2 -c 1 -o gpurun_out/prof_conv1_ffma python scripts/prof_conv.py --batch 64 --iters 1 --impl ffma --layer 1 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
