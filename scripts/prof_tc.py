"""Net A as a tensor-core chain: per-op device times (plan profile hook) and a short rollout for ncu captures.
    python scripts/prof_tc.py [--batch 64] [--iters 10] [--math tc|ffma] [--opt tc_debug=1 --opt fuse=1 ...]
--opt sets DlwpPlanOptions fields (tc_debug: 1 = epilogue only waits/arrives, 2 = issuer only commits, 3 = both)."""
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
ap.add_argument('--opt', action='append', default=[])
args = ap.parse_args()
os.environ['DLWP_MATH'] = args.math
opts = {k: int(v) for k, v in (o.split('=') for o in args.opt)}
import bench  # noqa: E402
from dlwp_b200 import _native as nat  # noqa: E402

from dlwp_b200.engine import CompiledNet  # noqa: E402
dlwp = bench.build_model()
eng = CompiledNet(dlwp.model, args.batch, options=opts)
x = torch.from_numpy(bench.make_inputs(args.batch)).cuda()
s = eng.rollout_device(x, 2, use_graph=False)
torch.cuda.synchronize()
import ctypes  # noqa: E402
cnt = (ctypes.c_int64 * 12)()
nat.lib().dlwp_debug_counters(cnt, 12)
t, per_row = [], []
for i in range(2):
    t.append(eng.profile_op(args.batch, i, args.iters))
    nat.lib().dlwp_debug_counters(cnt, 12)
    if cnt[3]:
        per_row.append('op%d clk/row: wait-acc %.0f wait-stage %.0f issue %.0f (rows %d)' % (
            i, cnt[0] / cnt[3], cnt[1] / cnt[3], cnt[2] / cnt[3], cnt[3]))
    if cnt[10]:
        per_row.append('op%d per CTA: span %.0f clk = %.1f us (%.3f GHz), first row after %.0f clk, rows/CTA %.1f' % (
            i, cnt[8] / cnt[10], 1e-3 * cnt[9] / cnt[10], cnt[8] / max(cnt[9], 1), cnt[11] / cnt[10], cnt[3] / cnt[10]))
    if cnt[10] and eng.fused_pair() < 0:
        per_row.append('op%d per CTA: phantom-slot waits %.0f clk, between segments %.0f clk; epilogue warp 0: waiting %.0f clk, '
                       'working %.0f clk' % (i, cnt[4] / cnt[10], cnt[5] / cnt[10], cnt[6] / cnt[10], cnt[7] / cnt[10]))
    elif cnt[7]:
        per_row.append('fused conv2 clk/row: wait-acc %.0f wait-intermediate %.0f issue %.0f (rows %d)' % (
            cnt[4] / cnt[7], cnt[5] / cnt[7], cnt[6] / cnt[7], cnt[7]))
if per_row:
    print(' | '.join(per_row))
print('math=%s tc=%s batch=%d: conv1 %.3f ms (%.2f us/sample)  conv2 %.3f ms (%.2f us/sample)  flags=%d opts=%s' % (
    args.math, eng.uses_tensor_cores(), args.batch, t[0], 1e3 * t[0] / args.batch, t[1], 1e3 * t[1] / args.batch,
    nat.lib().dlwp_debug_flags(), opts), flush=True)
eng.close()
