#!/bin/bash
mkdir -p gpurun_out
echo "== TMA probe (prefetch, negative coords)"; timeout 60 ./scripts/tma_probe 1 2>&1 | tail -8
echo "== TMA probe (no prefetch, in-bounds coords)"; timeout 60 ./scripts/tma_probe 0 0 0 2>&1 | tail -8
echo "== TMA probe under compute-sanitizer"; timeout 120 compute-sanitizer ./scripts/tma_probe 1 2>&1 | tail -15
echo "== pytest (no TMA)"
DLWP_NO_TMA=1 timeout 1200 python -m pytest tests -q -m gpu -k "not ffma_tma" 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_notma.log
echo "== pytest (TMA only)"
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "ffma_tma" 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_tma.log
echo "== bench (no TMA)"
DLWP_NO_TMA=1 timeout 900 python bench.py --steps 50 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_notma.log
echo "== ncu launch list"
DLWP_NO_TMA=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_notma.csv python bench.py --steps 4 --warmup 3 --batch 64 --no-cpu --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
