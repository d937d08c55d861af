#!/bin/bash
# round 2, 2 GPUs: the host-buffer lat-band path (dlwp_rollout_latband_host: band-only x0 upload + 4-byte max|x0| all-reduce,
# strided D2H pipelined behind the steps) and the in-library gradient all-reduce -- parity on real ranks, then numbers
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
show() { grep '^{' $1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$1', 'value %.0f ms/step %.4f e2e %.0f (%.4f s, h2d/step %.0f) scaling %s gb %s bitwise %s halo %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['seconds'], d['e2e']['h2d_bytes_per_step'], d['scaling'], c.get('global_batch'), c.get('bands_equal_single_domain_bitwise'), c.get('halo',{}).get('exchange')))
" 2>/dev/null || tail -12 $1 | cut -c1-300; }
timeout 200 $TR scripts/latband_check.py 2>&1 | grep -a "rank\|LATBAND\|Error\|error" | tail -12
timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/y_lat2_weak_p2p.log 2>&1; show gpurun_out/y_lat2_weak_p2p.log
timeout 200 $TR bench.py --gpus 2 --workload net_b --steps 20 --warmup 3 > gpurun_out/y_netb2_fp32_p2p.log 2>&1; show gpurun_out/y_netb2_fp32_p2p.log
timeout 200 $TR scripts/train_check.py --small --batch 8 --steps 3 2>&1 | grep rank | tail -2
timeout 200 $TR scripts/bench_train.py --batch 8 --steps 3 > gpurun_out/y_train2.log 2>&1; grep '^{' gpurun_out/y_train2.log | tail -1 | cut -c1-600
