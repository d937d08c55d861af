#!/bin/bash
# round 2, 2 GPUs: the host-buffer lat-band path (dlwp_rollout_latband_host) -- parity on real ranks, then bench e2e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f (%.4f s) scaling %s gb %s bitwise %s halo %s cpus %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), c.get('halo',{}).get('exchange'), c.get('host_cpus_bound_near_gpu')))
" 2>/dev/null || tail -12 $1 | cut -c1-300; }
nvidia-smi topo -m 2>&1 | head -8
timeout 300 python -m pytest tests/test_latband_gpu.py -q -x 2>&1 | tail -3
timeout 300 $TR scripts/latband_check.py 2>&1 | grep -v "^\[" | tail -10
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/y_lat2_weak_p2p.log 2>&1; show gpurun_out/y_lat2_weak_p2p.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --halo nccl > gpurun_out/y_lat2_weak_nccl.log 2>&1; show gpurun_out/y_lat2_weak_nccl.log
timeout 600 $TR bench.py --gpus 2 --workload net_b --precision bf16 --steps 20 --warmup 3 > gpurun_out/y_netb2_bf16_p2p.log 2>&1; show gpurun_out/y_netb2_bf16_p2p.log
