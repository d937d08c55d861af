#!/bin/bash
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
