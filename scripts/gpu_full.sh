#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== bench"
timeout 900 python bench.py --steps 200 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
