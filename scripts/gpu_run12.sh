#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/pytest_train.log
