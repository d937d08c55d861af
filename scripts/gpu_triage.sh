#!/bin/bash
# Bottleneck triage of the Net A tensor-core kernels: plan option tc_debug = 1 epilogue only waits/arrives,
# 2 issuer only commits, 3 neither (producer + barrier handshakes only), +4 the issuing warp times itself;
# fuse=0 (default) keeps one kernel per layer, fuse=1 runs conv1 -> conv2 as one kernel.
for fuse in 0 1; do
for dbg in 0 1 2 3 4 5; do timeout 120 python scripts/prof_tc.py --batch 256 --opt fuse=$fuse --opt tc_debug=$dbg 2>&1 | tail -2; done
done
timeout 120 python scripts/prof_tc.py --batch 256 --opt fuse=0 --opt tc_generic=1 2>&1 | tail -1
for nb in 1 2 3 4 6; do timeout 120 python scripts/prof_tc.py --batch 256 --opt fuse=1 --opt tc_bands=$nb 2>&1 | tail -1; done
