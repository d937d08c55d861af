#!/bin/bash
# round 2: where do the issuing warp's untimed clocks go (phantom-slot waits / between segments / epilogue) + Net B numbers
mkdir -p gpurun_out
for dbg in 4 5; do timeout 120 python scripts/prof_tc.py --batch 256 --opt tc_debug=$dbg 2>&1 | tail -2; done > gpurun_out/h_triage.txt 2>&1
cat gpurun_out/h_triage.txt
for b in 16 64; do timeout 300 python scripts/bench_net_b.py --batch $b --steps 20 --per-op 2>&1 | tail -1; done > gpurun_out/h_net_b.jsonl
cat gpurun_out/h_net_b.jsonl
