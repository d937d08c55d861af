#!/bin/bash
for dbg in 1 2; do DLWP_SW_DEBUG=$dbg timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1; done
