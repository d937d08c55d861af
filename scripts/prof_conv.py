"""Launch the two Net A conv layers a few times (for ncu captures / quick A-B timing of tiling overrides).

    python scripts/prof_conv.py [--batch 64] [--iters 5] [--impl auto|ffma|ffma_tma] [--layer 1|2|0]
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_b200 import _native as nat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--iters', type=int, default=5)
ap.add_argument('--impl', default='auto')
ap.add_argument('--layer', type=int, default=0)
args = ap.parse_args()
lib = nat.lib()
B = args.batch
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
LAYERS = {1: (6, 32, 3, 2, 1), 2: (32, 6, 5, 1, 0)}
for lid, (cin, cout, k, d, act) in LAYERS.items():
    if args.layer and args.layer != lid:
        continue
    rng = np.random.RandomState(3)
    x = torch.from_numpy(rng.standard_normal((B, cin, 91, 180)).astype(np.float32)).cuda()
    w = torch.from_numpy((0.05 * rng.standard_normal((k, k, cin, cout))).astype(np.float32)).cuda()
    b = torch.zeros(cout, device='cuda')
    y = torch.empty((B, cout, 91, 180), device='cuda')
    desc = nat.ConvDesc(N=B, Cin=cin, H=91, W=180, Cout=cout, kh=k, kw=k, dil_h=d, dil_w=d, pad_t=2, pad_b=2, pad_l=2,
                        pad_r=2, pad_mode_h=0, pad_mode_w=1, act=act, pre_op=0, rowwise=0, impl=nat.IMPLS[args.impl],
                        reserved=0, row_begin=0, row_end=0, x_stride_n=cin * 91 * 180, x_stride_c=91 * 180, x_stride_h=180,
                        y_stride_n=cout * 91 * 180, y_stride_c=91 * 180, y_stride_h=180)
    name = lib.dlwp_conv2d_impl_name(ctypes.byref(desc)).decode()
    for _ in range(2):
        nat.check(lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        lib.dlwp_conv2d_fwd(ctypes.byref(desc), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    fl = 2.0 * B * 91 * 180 * cin * cout * k * k
    print('layer %d (%d->%d k%d d%d) impl=%s batch=%d: %.3f ms  %.1f TFLOP/s  %.2f us/sample  flags=%d  env=%s' % (
        lid, cin, cout, k, d, name, B, ms, fl / ms / 1e9, 1e3 * ms / B, lib.dlwp_debug_flags(),
        {k_: v for k_, v in os.environ.items() if k_.startswith('DLWP_')}))
