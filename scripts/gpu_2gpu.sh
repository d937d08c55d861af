#!/bin/bash
echo "== latband check"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/latband_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
echo "== bench latband"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 50 --warmup 3 2>&1 | tail -1 | cut -c1-330
echo "== bench batch"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 50 --warmup 3 --parallel batch 2>&1 | tail -1 | cut -c1-330
