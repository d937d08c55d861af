#!/bin/bash
# round 2: bf16 mode tests, full GPU pass, Net B bench lines (fp32-equivalent and bf16), Net A bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bf16_gpu.py -q -x -s > gpurun_out/j_bf16.log 2>&1
echo "exit $?" >> gpurun_out/j_bf16.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit|bf16" gpurun_out/j_bf16.log | head -30
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/j_pytest.log | tail -15
timeout 600 python bench.py --workload net_b --steps 20 --warmup 3 > gpurun_out/j_netb_fp32.log 2>&1; tail -1 gpurun_out/j_netb_fp32.log | cut -c1-700
timeout 600 python bench.py --workload net_b --precision bf16 --steps 20 --warmup 3 --no-cpu > gpurun_out/j_netb_bf16.log 2>&1; tail -1 gpurun_out/j_netb_bf16.log | cut -c1-700
timeout 600 python bench.py --workload net_b --precision bf16 --batch 64 --steps 20 --warmup 3 --no-cpu > gpurun_out/j_netb_bf16_b64.log 2>&1; tail -1 gpurun_out/j_netb_bf16_b64.log | cut -c1-500
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/j_bench20.log 2>&1; tail -1 gpurun_out/j_bench20.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --precision bf16 > gpurun_out/j_bench20_bf16.log 2>&1; tail -1 gpurun_out/j_bench20_bf16.log | cut -c1-300
