#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 50 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_4gpu_latband.log | cut -c1-800
