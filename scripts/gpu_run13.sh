#!/bin/bash
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_training_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|Error|assert [0-9]" | tail -8
timeout 120 python scripts/prof_tc.py --batch 256 2>&1 | tail -1
