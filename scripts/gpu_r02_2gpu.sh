#!/bin/bash
# round 2, 2 GPUs, final build: halo over peer memory (default) vs NCCL; Net A and Net B; multi_gpu_model test; DP training
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f scaling %s gb %s bitwise %s halo %s launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), c.get('halo',{}).get('exchange'), d['gpu_launches']))
" 2>/dev/null || tail -8 $1 | cut -c1-300; }
timeout 300 python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/w_bench20.log 2>&1; show gpurun_out/w_bench20.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/w_lat2_weak_p2p.log 2>&1; show gpurun_out/w_lat2_weak_p2p.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --halo nccl > gpurun_out/w_lat2_weak_nccl.log 2>&1; show gpurun_out/w_lat2_weak_nccl.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/w_lat2_strong_p2p.log 2>&1; show gpurun_out/w_lat2_strong_p2p.log
timeout 600 $TR bench.py --gpus 2 --workload net_b --precision bf16 --steps 20 --warmup 3 > gpurun_out/w_netb2_bf16_p2p.log 2>&1; show gpurun_out/w_netb2_bf16_p2p.log
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-200
timeout 600 $TR scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/w_train2.log 2>&1; grep '^{' gpurun_out/w_train2.log | tail -1 | cut -c1-500
timeout 600 $TR scripts/train_check.py --small --batch 8 --steps 3 2>&1 | grep rank | tail -2
