#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
