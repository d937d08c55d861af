"""torchrun --nproc-per-node P scripts/latband_check.py
Every rank rolls its latitude band forward with the in-library NCCL halo exchange (eager and CUDA-graph), then computes
the single-domain rollout of the same small batch LOCALLY and compares its own band rows bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dlwp_b200.parallel import LatBandEngine  # noqa: E402

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
B, K = 4, 6
dlwp = bench.build_model()
eng = LatBandEngine(dlwp.model, B, rank, world, dist=dist)
x0 = bench.make_inputs(B)
xd = torch.from_numpy(x0).cuda()
ref = dlwp.model.engine(B).rollout_device(xd, K, use_graph=False)
torch.cuda.synchronize()
lo, hi = eng.me.band
ok = True
for use_graph in (False, True):
    series = torch.full((K, B) + bench.STATE, float('nan'), device='cuda')
    eng.rollout_device(xd, K, out=series, use_graph=use_graph)
    torch.cuda.synchronize()
    same = bool(torch.equal(series[:, :, :, lo:hi], ref[:, :, :, lo:hi]))
    print('rank %d/%d band [%d,%d) graph=%d: band rows == single-domain rows: %s' % (rank, world, lo, hi, use_graph, same),
          flush=True)
    ok = ok and same
# the reference-facing entry: numpy in, complete series on rank 0 (bands gathered with one collective)
full = eng.predict_timeseries(x0, K)
if rank == 0:
    same = bool(np.array_equal(full, ref.cpu().numpy()))
    print('rank 0: LatBandEngine.predict_timeseries (halo %s) == single-domain series: %s' % (eng.halo, same), flush=True)
    ok = ok and same
else:
    assert full is None
# host buffers in, this rank's band out, strided D2H pipelined behind the steps (groups of 2 of the 6 steps)
band = eng.rollout_host(x0, K, d2h_group=2)
same = bool(np.array_equal(band, ref[:, :, :, lo:hi].cpu().numpy()))
print('rank %d: LatBandEngine.rollout_host band == single-domain rows: %s' % (rank, same), flush=True)
ok = ok and same
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('LATBAND CHECK', 'PASSED' if int(flag.item()) else 'FAILED', flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)   # skip communicator teardown
