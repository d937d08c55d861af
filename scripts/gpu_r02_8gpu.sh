#!/bin/bash
# round 2, 8 GPUs: lat-band Net A (weak), Net B bf16, data-parallel training (BASELINE.json configs[3-4]), multi-rank check
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f scaling %s gb %s bitwise %s halo %s launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), c.get('halo',{}).get('exchange'), d['gpu_launches']))
" 2>/dev/null || tail -8 $1 | cut -c1-300; }
timeout 300 $TR scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/x8_train.log 2>&1; grep '^{' gpurun_out/x8_train.log | tail -1 | cut -c1-500
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/x8_lat_weak_p2p.log 2>&1; show gpurun_out/x8_lat_weak_p2p.log
timeout 300 $TR bench.py --gpus $N --workload net_b --precision bf16 --steps 20 --warmup 3 > gpurun_out/x8_netb_bf16_p2p.log 2>&1; show gpurun_out/x8_netb_bf16_p2p.log
timeout 300 $TR scripts/latband_check.py 2>&1 | grep -E "LATBAND|predict_timeseries" | tail -2
