#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== bench"
timeout 900 python bench.py --steps 200 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-200
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-200
echo "== ncu full sw kernels (N=256)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_sw_kernel -s 2 -c 2 -o gpurun_out/prof_sw_final python scripts/prof_tc.py --batch 256 --iters 1 2>&1 | tail -1
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
tail -c 150 gpurun_out/ncu_bench.log
