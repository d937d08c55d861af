#!/bin/bash
# Bottleneck triage of the two Net A tensor-core layers (profiles/r01_mma_ctx_probe.txt):
# DLWP_SW_DEBUG=1 epilogue only waits/arrives, 2 issuer only commits, 3 neither (producer + barrier handshakes only).
for dbg in 0 1 2 3; do DLWP_SW_DEBUG=$dbg timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1; done
