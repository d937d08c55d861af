"""Net B (skip U-Net of examples/train_functional.py:248-275, 12 x 180 x 360 state = BASELINE.json configs[2]) rollout
throughput on one GPU: tensor-core chain vs the fp32 FFMA path, per-op device times, algorithmic roofline (SURVEY.md §8d:
62,111,600 B fp32 and 4,678,041,600 FLOP per sample-step).  Prints one JSON line per run.
    python scripts/bench_net_b.py [--batch 16] [--steps 20] [--math tc|ffma]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--math', default='tc')
ap.add_argument('--per-op', action='store_true')
ap.add_argument('--opt', action='append', default=[], help='DlwpPlanOptions field, e.g. --opt tc_taps_in_k=1')
args = ap.parse_args()
os.environ['DLWP_MATH'] = args.math

from dlwp_b200 import keras  # noqa: E402
from dlwp_b200.custom import PeriodicPadding2D, slice_layer  # noqa: E402
from dlwp_b200.engine import CompiledNet  # noqa: E402
from dlwp_b200.keras.layers import Conv2D, Input, MaxPooling2D, UpSampling2D, ZeroPadding2D, concatenate  # noqa: E402

cs = (12, 180, 360)
cf = 'channels_first'
x_in = Input(shape=cs, name='input_0')
pp2, zp2 = PeriodicPadding2D(padding=(0, 2), data_format=cf), ZeroPadding2D(padding=(2, 0), data_format=cf)
pp1, zp1 = PeriodicPadding2D(padding=(0, 1), data_format=cf), ZeroPadding2D(padding=(1, 0), data_format=cf)
pool, up = MaxPooling2D(2, data_format=cf), UpSampling2D(2, data_format=cf)
kw = {'padding': 'valid', 'activation': 'tanh', 'data_format': cf}
c1, c2, c3 = Conv2D(32, 3, dilation_rate=2, **kw), Conv2D(64, 3, **kw), Conv2D(128, 3, **kw)
c4, c5 = Conv2D(32, 3, **kw), Conv2D(16, 3, dilation_rate=2, **kw)
c6 = Conv2D(cs[0], 5, padding='valid', activation='linear', data_format=cf)
x = c1(pp2(zp2(x_in)))
x, x1 = slice_layer(0, 16, axis=1)(x), slice_layer(16, 32, axis=1)(x)
x = c2(pp1(zp1(pool(x))))
x, x2 = slice_layer(0, 32, axis=1)(x), slice_layer(32, 64, axis=1)(x)
x = c3(pp1(zp1(pool(x))))
x = c4(pp1(zp1(up(x))))
x = up(concatenate([x, x2], axis=1))
x = c5(pp2(zp2(x)))
x = c6(pp2(zp2(concatenate([x, x1], axis=1))))
model = keras.Model(inputs=x_in, outputs=[x])
rng = np.random.RandomState(1)
for layer in model.layers:  # glorot-uniform kernels, small biases (synthetic weights; there is no checkpoint offline)
    ws = layer.get_weights() if hasattr(layer, 'get_weights') else []
    if len(ws) == 2:
        kh, kw_, ci, co = ws[0].shape
        lim = np.sqrt(6.0 / (kh * kw_ * (ci + co)))
        layer.set_weights([rng.uniform(-lim, lim, ws[0].shape).astype(np.float32),
                           (0.02 * rng.standard_normal(co)).astype(np.float32)])

N, K = args.batch, args.steps
opts = {k: int(v) for k, v in (o.split('=') for o in args.opt)}
eng = CompiledNet(model, N, impl='tc' if args.math == 'tc' else None, force_ffma=args.math != 'tc', options=opts)
x0 = torch.from_numpy(np.random.RandomState(0).standard_normal((N,) + cs).astype(np.float32)).cuda()
series = eng.rollout_device(x0, K, use_graph=True)   # warm-up + graph build
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e30
for _ in range(3):
    e0.record()
    eng.rollout_device(x0, K, use_graph=True, out=series)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
ms_step = best / K
alg_bytes, flops = 62111600.0 * N, 4678041600.0 * N
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6538.0}
line = {'workload': 'net_b_skip_unet_12x180x360_rollout (BASELINE.json configs[2])', 'math': args.math,
        'tensor_cores': bool(eng.uses_tensor_cores()), 'options': opts, 'precision': os.environ.get('DLWP_PRECISION', 'fp32'), 'batch': N, 'steps': K, 'ms_per_step': ms_step,
        'forecast_steps_per_sec': N / (ms_step * 1e-3), 'algorithmic_gbs': alg_bytes / (ms_step * 1e-3) / 1e9,
        'hbm_frac': alg_bytes / (ms_step * 1e-3) / 1e9 / peaks['hbm_gbs'], 'useful_tflops': flops / (ms_step * 1e-3) / 1e12,
        'finite': bool(torch.isfinite(series[-1]).all().item())}
if args.per_op:
    n_ops = 12
    line['per_op_ms'] = [round(eng.profile_op(N, i, 5), 4) for i in range(n_ops)]
print(json.dumps(line), flush=True)
eng.close()
