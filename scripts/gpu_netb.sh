#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/bench_net_b.py --batch 16 --steps 20 --math tc --per-op 2>&1 | tail -1 | tee -a gpurun_out/net_b.jsonl
timeout 300 python scripts/bench_net_b.py --batch 64 --steps 10 --math tc 2>&1 | tail -1 | tee -a gpurun_out/net_b.jsonl
for b in 1 32; do timeout 300 python bench.py --batch $b --steps 200 --warmup 3 --no-cpu --e2e-steps 8 2>&1 | tail -1 | tee gpurun_out/bench_b$b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('net A batch', d['config']['batch_per_gpu'], 'ms/step', d['ms_per_step'], 'forecast-steps/s', d['value'], 'e2e', d['e2e']['value'])"; done
