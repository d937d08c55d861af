"""Net A as a tensor-core chain: per-op device times (plan profile hook) and a short rollout for ncu captures.
    python scripts/prof_tc.py [--batch 64] [--iters 10] [--math tc|ffma]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--math', default='tc')
args = ap.parse_args()
os.environ['DLWP_MATH'] = args.math
import bench  # noqa: E402
from dlwp_b200 import _native as nat  # noqa: E402

dlwp = bench.build_model()
eng = dlwp.model.engine(args.batch)
x = torch.from_numpy(bench.make_inputs(args.batch)).cuda()
s = eng.rollout_device(x, 2, use_graph=False)
torch.cuda.synchronize()
t = [eng.profile_op(args.batch, i, args.iters) for i in range(2)]
print('math=%s tc=%s batch=%d: conv1 %.3f ms (%.2f us/sample)  conv2 %.3f ms (%.2f us/sample)  flags=%d env=%s' % (
    args.math, eng.uses_tensor_cores(), args.batch, t[0], 1e3 * t[0] / args.batch, t[1], 1e3 * t[1] / args.batch,
    nat.lib().dlwp_debug_flags(), {k: v for k, v in os.environ.items() if k.startswith('DLWP_T')}), flush=True)
eng.close()
dlwp.model._engine = None
