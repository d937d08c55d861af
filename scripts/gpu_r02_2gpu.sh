#!/bin/bash
# round 2, 2 GPUs: lat-band Net A (weak + strong scaling) with the fused halo kernels, data-parallel training (configs[4])
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/i_lat2_weak.log 2>&1
grep '^{' gpurun_out/i_lat2_weak.log | tail -1 | cut -c1-900; tail -3 gpurun_out/i_lat2_weak.log | grep -v '^{' | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/i_lat2_strong.log 2>&1
grep '^{' gpurun_out/i_lat2_strong.log | tail -1 | cut -c1-400
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --parallel batch --no-cpu > gpurun_out/i_batch2.log 2>&1
grep '^{' gpurun_out/i_batch2.log | tail -1 | cut -c1-300
timeout 600 python scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/i_train1.log 2>&1; tail -1 gpurun_out/i_train1.log | cut -c1-600
timeout 600 $TR scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/i_train2.log 2>&1; grep '^{' gpurun_out/i_train2.log | tail -1 | cut -c1-600
timeout 600 $TR scripts/train_check.py --small --batch 8 --steps 3 > gpurun_out/i_traincheck.log 2>&1; grep 'rank' gpurun_out/i_traincheck.log | tail -2
