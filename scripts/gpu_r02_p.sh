#!/bin/bash
# round 2: full GPU pass after the peers-flavour / probe removal, per-layer times, bench, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/p_pytest.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/p_pytest.log | tail -25
for rep in 1 2; do timeout 120 python scripts/prof_tc.py --batch 256 --iters 20 2>&1 | tail -1; done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench20.log 2>&1; tail -1 gpurun_out/p_bench20.log | cut -c1-250
bash scripts/gpu_r02_sanitize.sh
