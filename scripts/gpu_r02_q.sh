#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/q_pytest.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/q_pytest.log | tail -25
timeout 300 python scripts/bench_net_b.py --batch 64 --steps 20 --per-op 2>&1 | tail -1 | cut -c1-600
DLWP_PRECISION=bf16 timeout 300 python scripts/bench_net_b.py --batch 64 --steps 20 --per-op 2>&1 | tail -1 | cut -c1-600
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/q_bench20.log 2>&1; tail -1 gpurun_out/q_bench20.log | cut -c1-250
bash scripts/gpu_r02_sanitize.sh
