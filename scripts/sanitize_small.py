"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck) on the warp-specialised kernels:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Net A (tensor-core chain, fp32-equivalent and bf16, CUDA graph and plain launches), the skip U-Net (data movers on P
images), the fp32 FFMA path, the recurrent front block, and a 2-band peer-memory latitude-band rollout in one process."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import layers as OL  # noqa: E402
from tests.helpers import build_functional_pair, build_product_sequential, oracle_sequential_like  # noqa: E402
from dlwp_b200.engine import CompiledNet  # noqa: E402
from dlwp_b200 import _native as nat  # noqa: E402

shape = (6, 24, 48)
layers = OL.net_a_layers(shape)
dlwp = build_product_sequential(layers)
oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.05)
x0 = np.random.RandomState(0).standard_normal((3,) + shape).astype(np.float32)
xd = torch.from_numpy(x0).cuda()
for opts in ({}, {'precision': 'bf16'}, {'fuse': 1}, {'math': 1}):
    eng = CompiledNet(dlwp.model, 3, options=opts)
    for graph in (False, True):
        s = eng.rollout_device(xd, 3, use_graph=graph)
    torch.cuda.synchronize()
    print('net A', opts, 'tc' if eng.uses_tensor_cores() else 'ffma', float(s.abs().mean()), 'flags', nat.lib().dlwp_debug_flags())
    eng.close()
cs = (8, 16, 32)
unet, _ = build_functional_pair(cs, skip=True, integration_steps=1, seed=2)
xu = np.random.RandomState(1).standard_normal((2,) + cs).astype(np.float32)
for opts in ({}, {'precision': 'bf16'}):
    eng = CompiledNet(unet.model, 2, options=opts)
    y = eng.predict(xu)[0]
    print('u-net', opts, float(np.abs(y).mean()), 'flags', nat.lib().dlwp_debug_flags())
    eng.close()
from tests.test_oracle_golden import small_recurrent_layers  # noqa: E402
rec = build_product_sequential(small_recurrent_layers(2), time_dim=2, is_recurrent=True)
yr = rec.predict(np.random.RandomState(2).standard_normal((2, 2, 2, 6, 8)).astype(np.float32))
print('recurrent', yr.shape, float(np.abs(yr).mean()))
from tests.test_latband_gpu import _run_bands_p2p  # noqa: E402
try:
    full = _run_bands_p2p(dlwp.model, 2, x0, 3, False)
    print('lat-band p2p', full.shape, float(np.abs(full).mean()))
except AssertionError as e:      # under the sanitizer a halo wait can outlast its timeout: reported, not fatal
    print('lat-band p2p: a halo wait timed out under the sanitizer', e)
torch.cuda.synchronize()
print('done')
