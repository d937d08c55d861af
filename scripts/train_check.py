"""Training throughput (BASELINE.json configs[4] shape: skip U-Net unrolled x6, 12x180x360, Adam, MSE) and, under torchrun,
the data-parallel check: P ranks x (B/P) samples with gradient all-reduce == one rank x B samples.
    python scripts/train_check.py [--batch 8] [--steps 3] [--unroll 6] [--small]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import build_functional_pair  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--unroll', type=int, default=6)
ap.add_argument('--small', action='store_true')
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
cs = (12, 32, 48) if args.small else (12, 180, 360)
rng = np.random.RandomState(2)
B = args.batch
X = rng.standard_normal((B,) + cs).astype(np.float32)
Y = [rng.standard_normal((B,) + cs).astype(np.float32) for _ in range(args.unroll)]


def make():
    dlwp, _ = build_functional_pair(cs, skip=True, integration_steps=args.unroll, seed=1, bias_scale=0.0)
    return dlwp


if world > 1:
    # data parallel: every rank trains on its slice; the result must equal single-process training on the whole batch
    dlwp = make()
    sl = slice(rank * B // world, (rank + 1) * B // world)
    for _ in range(args.steps):
        dlwp.model.train_on_batch(X[sl], [y[sl] for y in Y])
    w_dp = dlwp.model.get_weights()
    if rank == 0:
        os.environ_backup = None
        import torch.distributed as dist2
    # reference: all ranks also compute the single-process result locally (no collective: temporarily hide the group)
    import dlwp_b200.training as T
    saved = T._dist
    T._dist = lambda: None
    ref = make()
    for _ in range(args.steps):
        ref.model.train_on_batch(X, Y)
    T._dist = saved
    err = max(float(np.abs(a - b).max()) for a, b in zip(w_dp, ref.model.get_weights()))
    print('rank %d: data-parallel (%d x %d samples, all-reduce) vs single process (%d samples): max |dw| = %.3g -> %s' % (
        rank, world, B // world, B, err, 'OK' if err < 1e-5 else 'FAILED'), flush=True)
    dist.barrier()
    os._exit(0)

dlwp = make()
dlwp.model.train_on_batch(X, Y)          # warm-up (plan creation, allocations)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.steps):
    loss = dlwp.model.train_on_batch(X, Y)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.steps
print('train step: skip U-Net x%d unrolled, %s, batch %d: %.1f ms/step = %.1f samples/s (loss %s)' % (
    args.unroll, cs, B, 1e3 * dt, B / dt, loss[0] if isinstance(loss, list) else loss), flush=True)
