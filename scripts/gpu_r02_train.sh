#!/bin/bash
# round 2: training path with the tiled backward-weight kernel and dX through the forward kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q 2>&1 | tail -12
timeout 600 python scripts/bench_train.py --batch 8 --steps 3 2>&1 | tail -1 | cut -c1-500
timeout 600 python scripts/bench_train.py --batch 16 --steps 3 2>&1 | tail -1 | cut -c1-500
for o in "" "--opt tc_taps_in_k=1"; do DLWP_PRECISION=bf16 timeout 300 python scripts/bench_net_b.py --batch 64 --steps 20 --per-op $o 2>&1 | tail -1 | cut -c1-700; done
