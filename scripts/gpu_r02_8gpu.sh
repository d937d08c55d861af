#!/bin/bash
# round 2, 8 GPUs, final build (trimmed: 8 GPUs are charged 8x): lat-band weak (peer-memory halo, pipelined host path in e2e),
# Net B bf16 lat-band, data-parallel training with the tiled backward kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f (%.4f s) scaling %s gb %s bitwise %s halo %s cpus %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), c.get('halo',{}).get('exchange'), c.get('host_cpus_bound_near_gpu')))
" 2>/dev/null || tail -12 $1 | cut -c1-300; }
nvidia-smi topo -m 2>&1 | head -10 | cut -c1-160
timeout 100 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/x8_lat_weak_p2p.log 2>&1; show gpurun_out/x8_lat_weak_p2p.log
timeout 100 $TR bench.py --gpus 8 --workload net_b --precision bf16 --steps 20 --warmup 3 > gpurun_out/x8_netb_bf16_p2p.log 2>&1; show gpurun_out/x8_netb_bf16_p2p.log
timeout 100 $TR scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/x8_train.log 2>&1; grep '^{' gpurun_out/x8_train.log | tail -1 | cut -c1-500
