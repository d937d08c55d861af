#!/bin/bash
# round 2, 2 GPUs: lat-band with the exchange overlapped (default) vs serialized, Net A and Net B (fp32-equivalent, bf16)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f scaling %s gb %s bitwise %s launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), d['gpu_launches']))
" 2>/dev/null || tail -5 $1 | cut -c1-300; }
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k_lat2_weak.log 2>&1; show gpurun_out/k_lat2_weak.log
DLWP_LATBAND_SPARE_SMS=-1 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k_lat2_weak_serial.log 2>&1; show gpurun_out/k_lat2_weak_serial.log
DLWP_LATBAND_SPARE_SMS=16 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k_lat2_weak_s16.log 2>&1; show gpurun_out/k_lat2_weak_s16.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/k_lat2_strong.log 2>&1; show gpurun_out/k_lat2_strong.log
timeout 600 $TR bench.py --gpus 2 --workload net_b --steps 20 --warmup 3 > gpurun_out/k_netb2_fp32.log 2>&1; show gpurun_out/k_netb2_fp32.log
timeout 600 $TR bench.py --gpus 2 --workload net_b --precision bf16 --steps 20 --warmup 3 > gpurun_out/k_netb2_bf16.log 2>&1; show gpurun_out/k_netb2_bf16.log
timeout 600 $TR bench.py --gpus 2 --workload net_b --precision bf16 --steps 20 --warmup 3 --scaling strong > gpurun_out/k_netb2_bf16_strong.log 2>&1; show gpurun_out/k_netb2_bf16_strong.log
