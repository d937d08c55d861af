#!/bin/bash
timeout 200 python -m pytest tests/test_conv_tc_gpu.py tests/test_rollout_gpu.py -x -q -m gpu 2>&1 | tail -3
