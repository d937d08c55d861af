"""BASELINE.json configs[4]: data-parallel training throughput -- skip U-Net (examples/train_functional.py:248-275) unrolled
x6 with shared weights, 12 x 180 x 360, MSE with loss_weights 1/6, Adam; one process per GPU, gradients all-reduced over
NCCL between backward and the Adam update (dlwp_b200/training.py).  Under torchrun every rank trains on its own --batch
samples (weak scaling).  Prints one JSON line: samples/s over all GPUs (device time, max over ranks), the all-reduce's own
time, and its share of the step.
    python -m torch.distributed.run --nproc-per-node N scripts/bench_train.py [--batch 8] [--steps 5] [--unroll 6]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import build_functional_pair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8, help='samples per GPU')
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--unroll', type=int, default=6)
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cs = (12, 180, 360)
rng = np.random.RandomState(2 + rank)
B = args.batch
X = rng.standard_normal((B,) + cs).astype(np.float32)
Y = [rng.standard_normal((B,) + cs).astype(np.float32) for _ in range(args.unroll)]
dlwp, _ = build_functional_pair(cs, skip=True, integration_steps=args.unroll, seed=1, bias_scale=0.0)
for _ in range(2):
    dlwp.model.train_on_batch(X, Y)       # warm-up: plan creation, NCCL channels
torch.cuda.synchronize()


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier()
e0.record()
for _ in range(args.steps):
    loss = dlwp.model.train_on_batch(X, Y)
e1.record()
barrier()
ms = e0.elapsed_time(e1) / args.steps
# the all-reduce alone, on the real gradient buffer
g = dlwp.model._train_engine.grad_tensor()
ar_ms = 0.0
if dist is not None:
    from dlwp_b200 import training
    eng = dlwp.model._train_engine
    for _ in range(3):
        training._allreduce_gradients(eng, dist)     # the library's ncclAllReduce(ncclAvg) on its own communicator
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(20):
        training._allreduce_gradients(eng, dist)
    a1.record()
    barrier()
    ar_ms = a0.elapsed_time(a1) / 20
    t = torch.tensor([ms, ar_ms], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0].item()), float(t[1].item())
if rank == 0:
    print(json.dumps({
        'workload': 'train skip U-Net x%d unrolled 12x180x360, Adam, MSE (BASELINE.json configs[4])' % args.unroll,
        'n_gpus': world, 'batch_per_gpu': B, 'global_batch': B * world, 'steps': args.steps, 'ms_per_step': ms,
        'samples_per_sec': B * world / (ms * 1e-3), 'scaling': 'weak', 'gradient_bytes': int(g.numel()) * 4,
        'allreduce_ms': ar_ms, 'allreduce': 'dlwp_train_allreduce: ncclAllReduce(ncclAvg) inside the library', 'allreduce_fraction_of_step': ar_ms / ms if ms else None,
        'loss': float(loss[0] if isinstance(loss, list) else loss), 'math': 'fp32 FFMA forward + backward kernels'}), flush=True)
if dist is not None:
    dist.barrier()
os._exit(0)
