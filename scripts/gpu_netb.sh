#!/bin/bash
mkdir -p gpurun_out
for m in tc ffma; do timeout 300 python scripts/bench_net_b.py --batch 16 --steps 20 --math $m --per-op 2>&1 | tail -1 | tee -a gpurun_out/net_b.jsonl; done
timeout 300 python scripts/bench_net_b.py --batch 64 --steps 10 --math tc 2>&1 | tail -1 | tee -a gpurun_out/net_b.jsonl
