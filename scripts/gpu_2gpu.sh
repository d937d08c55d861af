#!/bin/bash
echo "== DP training check"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/train_check.py --batch 8 --steps 3 --unroll 2 --small 2>&1 | grep -E "rank|Error|error" | tail -6
echo "== training throughput (1 GPU of the box)"
timeout 300 python scripts/train_check.py --batch 16 --steps 2 --unroll 6 2>&1 | tail -2
