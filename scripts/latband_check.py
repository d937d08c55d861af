"""torchrun --nproc-per-node P scripts/latband_check.py : lat-band rollout (native NCCL, CUDA graph) vs single domain."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dlwp_b200.parallel import LatBandEngine  # noqa: E402

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
B, K = int(os.environ.get('B', '32')), int(os.environ.get('K', '6'))
dlwp = bench.build_model()
eng = LatBandEngine(dlwp.model, B, rank, world, dist=dist)
x0 = bench.make_inputs(B)
xd = torch.from_numpy(x0).cuda()
series = torch.zeros((K, B) + bench.STATE, device='cuda')
for use_graph in (0, 1):
    series.zero_()
    eng.rollout_device(xd, K, out=series, use_graph=bool(use_graph))
    torch.cuda.synchronize()
    lo, hi = eng.me.band
    band = series[:, :, :, lo:hi].contiguous()
    parts = [torch.empty((K, B, 6, p.band[1] - p.band[0], 180), device='cuda') for p in eng.planners]
    for r in range(world):
        t = band if r == rank else parts[r]
        dist.broadcast(t, src=r)
        if r == rank:
            parts[r] = band
    if rank == 0:
        full = torch.cat(parts, dim=3).cpu().numpy()
        ref = dlwp.predict_timeseries(x0, K)
        print('graph=%d world=%d: bands == single domain: %s (max abs diff %.3g)' % (
            use_graph, world, np.array_equal(full, ref), np.abs(full - ref).max()), flush=True)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    eng.rollout_device(xd, K, out=series, use_graph=True)
torch.cuda.synchronize()
if rank == 0:
    print('graph replay: %.3f ms per step' % (1e3 * (time.perf_counter() - t0) / 3 / K), flush=True)
eng.close()
dist.destroy_process_group()
