"""
The plain-bf16 mode of the tensor-core chain (DlwpPlanOptions.precision = 1; BASELINE.json configs[2]: "U-Net, 12-chan
1 degree grid, bf16"): activations and weights are stored as bf16, one MMA pass, fp32 accumulation, bias / tanh in fp32.

Two checks per net: (1) against a float64 oracle that applies the SAME roundings (inputs, weights and every stored
intermediate rounded to bf16) -- the kernels must match it to accumulation-order level; (2) against the unrounded float64
oracle -- the precision the mode delivers, reported and loosely bounded (bf16 parity is not gated at 1e-4).
"""

import numpy as np
import pytest

from oracle import layers as OL
from oracle import ops as OO
from tests.helpers import bf16_round, build_functional_pair, build_product_sequential, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    import torch
    from dlwp_b200 import _native
    _native.lib()
    return _native, torch


def _net_a_bf16_oracle(ws, x, steps):
    """Net A with bf16 storage: state image bf16(x) -> conv1 (bf16 weights) -> tanh -> bf16 -> conv2 -> fp32 output."""
    k1, b1, k2, b2 = [np.asarray(w, np.float64) for w in ws]
    k1, k2 = bf16_round(k1), bf16_round(k2)
    out = []
    for _ in range(steps):
        h = OO.pad_conv2d_closed_form(bf16_round(x), k1, b1, (2, 2), (2, 2), (2, 2), 'zero', 'periodic')
        h = bf16_round(np.tanh(h))
        x = OO.pad_conv2d_closed_form(h, k2, b2, (1, 1), (2, 2), (2, 2), 'zero', 'periodic')
        out.append(x)
    return np.stack(out)


def test_net_a_bf16_matches_bf16_rounded_oracle(env):
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    shape = (6, 46, 92)
    layers = OL.net_a_layers(shape)
    dlwp = build_product_sequential(layers)
    net = OL.OSequential(layers)
    OL.init_weights(net.conv_layers, seed=3, bias_scale=0.05)
    dlwp.model.set_weights(net.get_weights())
    x0 = np.random.RandomState(5).standard_normal((3,) + shape).astype(np.float32)
    eng = CompiledNet(dlwp.model, 3, options={'precision': 'bf16'})
    assert eng.uses_tensor_cores()
    got = eng.rollout_device(torch.from_numpy(x0).cuda(), 4, use_graph=True).cpu().numpy()
    assert nat.lib().dlwp_debug_flags() == 0
    ref_q = _net_a_bf16_oracle(net.get_weights(), x0.astype(np.float64), 4)
    # the stored roundings are identical, so only fp32 accumulation order and the tanh polynomial differ; a bf16 tie that
    # flips moves one intermediate by 2^-9 relative: allow a few of those
    assert rel_err(got[0], ref_q[0]) < 2e-4, rel_err(got[0], ref_q[0])
    assert rel_err(got, ref_q) < 2e-3
    ref = np.stack([net.forward(x0.astype(np.float64))])
    err = rel_err(got[0], ref[0])
    print('bf16 Net A, one application vs float64 oracle: %.2e' % err)
    assert 1e-5 < err < 3e-2          # really bf16 (not the fp32-equivalent path), and sane
    eng.close()


@pytest.mark.parametrize('skip', [True, False])
def test_unet_bf16_predict(env, skip):
    """The U-Net (pool / upsample / slice / concatenate as bf16 P-image data movers) against the unrounded oracle."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    cs = (8, 32, 64)
    dlwp, onet = build_functional_pair(cs, skip=skip, integration_steps=1, seed=2, bias_scale=0.05)
    x0 = np.random.RandomState(6).standard_normal((2,) + cs).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2, options={'precision': 'bf16'})
    assert eng.uses_tensor_cores()
    got = eng.predict(x0)[0]
    assert nat.lib().dlwp_debug_flags() == 0
    ref = onet.forward(x0.astype(np.float64))
    e16 = rel_err(got, ref)
    eng32 = CompiledNet(dlwp.model, 2)
    e32 = rel_err(eng32.predict(x0)[0], ref)
    print('U-Net skip=%s: bf16 %.2e, fp32-equivalent %.2e' % (skip, e16, e32))
    assert e32 < 2e-5 and 1e-4 < e16 < 5e-2
    ser = eng.rollout_device(torch.from_numpy(x0).cuda(), 3, use_graph=True).cpu().numpy()
    assert np.isfinite(ser).all() and rel_err(ser[0], ref) < 5e-2
    eng.close()
    eng32.close()


def test_bf16_generic_instances(env):
    """Layers without a folded bf16 instance take the generic bf16 kernels: 6->16 3x3 relu, 16->6 5x5."""
    nat, torch = env
    from dlwp_b200.engine import CompiledNet
    cf = 'channels_first'
    shape = (6, 20, 40)
    layers = (('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': shape}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (16, 3), {'activation': 'relu', 'data_format': cf}),
              ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
              ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
              ('Conv2D', (6, 5), {'activation': 'linear', 'data_format': cf}))
    dlwp = build_product_sequential(layers)
    net = OL.OSequential(layers)
    OL.init_weights(net.conv_layers, seed=4, bias_scale=0.05)
    dlwp.model.set_weights(net.get_weights())
    x0 = np.random.RandomState(7).standard_normal((2,) + shape).astype(np.float32)
    eng = CompiledNet(dlwp.model, 2, options={'precision': 1})
    assert eng.uses_tensor_cores()
    got = eng.predict(x0)[0]
    k1, b1, k2, b2 = [np.asarray(w, np.float64) for w in net.get_weights()]
    h = OO.pad_conv2d_closed_form(bf16_round(x0), bf16_round(k1), b1, (1, 1), (1, 1), (1, 1), 'zero', 'periodic')
    h = bf16_round(np.maximum(h, 0))
    ref_q = OO.pad_conv2d_closed_form(h, bf16_round(k2), b2, (1, 1), (2, 2), (2, 2), 'zero', 'periodic')
    assert rel_err(got, ref_q) < 2e-4, rel_err(got, ref_q)
    eng.close()
