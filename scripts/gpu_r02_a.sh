#!/bin/bash
# Round 2, GPU call A: full GPU test suite after the scaled-split rewrite, a bench line, a launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/a_bench50.log 2>&1
tail -2 gpurun_out/a_bench50.log
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu > gpurun_out/a_bench200.log 2>&1
tail -1 gpurun_out/a_bench200.log
