#!/bin/bash
# round 2: peer-memory halo emulated on one GPU, channels_last, step_sequence on device, conv backward ABI, full pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_latband_gpu.py -q -x > gpurun_out/l_latband.log 2>&1
echo "exit $?" >> gpurun_out/l_latband.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/l_latband.log | head -30
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/l_pytest.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/l_pytest.log | tail -25
