"""
GPU parity of the rollout path through the reference-facing API (DLWPNeuralNet / DLWPFunctional) against the oracle,
the reference-generated golden series, and size-independent properties at full size.

Gate (BASELINE.json): max|gpu - oracle64| / max|oracle64| <= 1e-4 after 50 feedback steps, fp32.
"""

import os

import numpy as np
import pytest

from oracle import layers as OL
from oracle import rollout as OR
from tests.helpers import (build_functional_pair, build_product_sequential, oracle_rollout64, oracle_sequential_like,
                           rel_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch_cuda():
    import torch
    from dlwp_b200 import _native
    _native.lib()
    return torch


def _small_layers(C, H, W):
    cf = 'channels_first'
    return (('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': (C, H, W)}),
            ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
            ('Conv2D', (8, 3), {'activation': 'tanh', 'data_format': cf}),
            ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
            ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
            ('Conv2D', (C, 3), {'dilation_rate': 2, 'activation': 'linear', 'data_format': cf}))


def test_net_a_50_step_rollout_meets_the_1e4_gate(torch_cuda):
    """BASELINE.json configs[1] net and state shape, N=2, 50 steps, vs the float64 oracle; also reports fp32-CPU drift."""
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.0)
    x0 = np.random.RandomState(0).standard_normal((2, 6, 91, 180)).astype(np.float32)
    got = dlwp.predict_timeseries(x0, 50)
    assert got.shape == (50, 2, 6, 91, 180) and got.dtype == np.float32
    ref = oracle_rollout64(net, x0, 50)
    per_step = [rel_err(got[t], ref[t]) for t in range(50)]
    assert max(per_step) <= 1e-4, per_step
    assert per_step[-1] <= 1e-4


def test_net_a_matches_series_produced_by_the_reference_torch_twin(torch_cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, 'torchnn_net_a.npz'))
    for tag in ('small', 'full'):
        x0 = g[tag + '_x0']
        layers = OL.net_a_layers(x0.shape[1:])
        dlwp = build_product_sequential(layers)
        dlwp.model.set_weights([g[tag + '_k1'], g[tag + '_b1'], g[tag + '_k2'], g[tag + '_b2']])
        sub = int(g[tag + '_sub'])
        y = dlwp.predict_timeseries(x0, int(g[tag + '_steps']))
        assert rel_err(y[:, :, :, ::sub, ::sub], g[tag + '_y'].astype(np.float64)) < 1e-5, tag


@pytest.mark.parametrize('time_dim', [1, 2, 3])
def test_neuralnet_flags_match_reference_loop_goldens(torch_cuda, golden_dir, time_dim):
    """Same nets / inputs as tests/golden/rollout_neuralnet.npz (produced by the reference's own loop code)."""
    g = np.load(os.path.join(golden_dir, 'rollout_neuralnet.npz'))
    layers = _small_layers(2 * time_dim, 6, 8)
    dlwp = build_product_sequential(layers, time_dim=time_dim)
    dlwp.model.set_weights([g['w_td%d_%d' % (time_dim, k)] for k in range(4)])
    x0 = g['x0_td%d' % time_dim]
    for steps in (1, 5):
        for ss in (False, True):
            for ktd in (False, True):
                key = 'y_td%d_s%d_ss%d_k%d' % (time_dim, steps, ss, ktd)
                y = dlwp.predict_timeseries(x0, steps, step_sequence=ss, keep_time_dim=ktd)
                assert y.shape == g[key].shape, key
                assert rel_err(y, g[key].astype(np.float64)) < 1e-5, key
    x_before = x0.copy()
    dlwp.predict_timeseries(x0, 3)
    np.testing.assert_array_equal(x0, x_before)          # inputs are never mutated
    with pytest.raises(ValueError):
        dlwp.predict_timeseries(x0, 0)


@pytest.mark.parametrize('n_steps', [1, 3])
def test_functional_multi_output_rollout_matches_reference_loop_goldens(torch_cuda, golden_dir, n_steps):
    from dlwp_b200 import keras
    from dlwp_b200.model import DLWPFunctional
    g = np.load(os.path.join(golden_dir, 'rollout_functional.npz'))
    for time_dim in (1, 2):
        seq = build_product_sequential(_small_layers(2 * time_dim, 6, 8), time_dim=time_dim).model
        seq.set_weights([g['w_td%d_%d' % (time_dim, k)] for k in range(4)])
        x_in = keras.Input(shape=(2 * time_dim, 6, 8))

        def apply(t):
            for layer in seq.layers:
                t = layer(t)
            return t
        outs = [apply(x_in)]
        for _ in range(1, n_steps):
            outs.append(apply(outs[-1]))
        dlwp = DLWPFunctional(time_dim=time_dim)
        dlwp.build_model(keras.Model(inputs=x_in, outputs=outs if n_steps > 1 else outs[0]), loss='mse',
                         optimizer='adam')
        assert dlwp._n_steps == n_steps
        x0 = g['x0_td%d' % time_dim]
        for steps in (1, 4, 7):
            for ktd in (False, True):
                key = 'y_td%d_n%d_s%d_k%d' % (time_dim, n_steps, steps, ktd)
                y = dlwp.predict_timeseries(x0, steps, keep_time_dim=ktd)
                assert y.shape == g[key].shape, key
                assert rel_err(y, g[key].astype(np.float64)) < 1e-5, key


@pytest.mark.parametrize('skip', [True, False])
def test_unet_predict_and_rollout_match_oracle(torch_cuda, skip):
    """Net B (examples/train_functional.py skip_model / basic_model) on a reduced 12x24x48 grid, 2 unrolled steps."""
    cs = (12, 24, 48)
    dlwp, onet = build_functional_pair(cs, skip=skip, integration_steps=2, seed=3)
    x0 = np.random.RandomState(4).standard_normal((3,) + cs).astype(np.float32)
    outs = dlwp.predict(x0)
    refs = onet.forward(x0.astype(np.float64))
    assert isinstance(outs, list) and len(outs) == 2
    for o, r in zip(outs, refs):
        assert rel_err(o, r) < 5e-5
    y = dlwp.predict_timeseries(x0, 6)
    ref = OR.functional_predict_timeseries(lambda p: onet.forward(p), x0.astype(np.float64), 6, n_steps=2,
                                           dtype=np.float64)
    assert y.shape == ref.shape == (6, 3) + cs
    assert rel_err(y, ref) < 1e-4


def test_unet_full_grid_single_application(torch_cuda):
    """Net B at BASELINE.json configs[2] shape (12, 180, 360), N=1, one application, vs torch-CPU tier-1 (fp32)."""
    import torch
    cs = (12, 180, 360)
    dlwp, onet = build_functional_pair(cs, skip=True, integration_steps=1, seed=5)
    x0 = np.random.RandomState(6).standard_normal((1,) + cs).astype(np.float32)
    y = dlwp.predict(x0)
    with torch.no_grad():
        ref = onet.forward(torch.from_numpy(x0)).numpy()
    assert rel_err(y, ref.astype(np.float64)) < 5e-5


def test_row_connected_last_layer(torch_cuda):
    cs = (12, 24, 48)
    dlwp, onet = build_functional_pair(cs, skip=True, integration_steps=1, seed=7, latitude_dependent=True)
    x0 = np.random.RandomState(8).standard_normal((2,) + cs).astype(np.float32)
    assert rel_err(dlwp.predict(x0), onet.forward(x0.astype(np.float64))) < 5e-5


def test_full_size_properties(torch_cuda):
    """
    Size-independent properties at the benchmark size (N=16, 6x91x180, 20 steps), no oracle needed:
    * longitude-shift equivariance: rolling the input by k columns rolls every forecast by k columns, BIT EXACT
      (the wrap is exact data movement and every pixel sums its taps in the same order);
    * sample independence: each sample's series equals the series of that sample run alone, bit exact;
    * device path == host path, graph == no graph.
    """
    torch = torch_cuda
    layers = OL.net_a_layers()
    dlwp = build_product_sequential(layers)
    oracle_sequential_like(dlwp, layers, seed=1, bias_scale=0.02)
    x0 = np.random.RandomState(0).standard_normal((16, 6, 91, 180)).astype(np.float32)
    y = dlwp.predict_timeseries(x0, 20)
    assert np.isfinite(y).all()
    y_roll = dlwp.predict_timeseries(np.roll(x0, 37, axis=3), 20)
    np.testing.assert_array_equal(y_roll, np.roll(y, 37, axis=4))
    y_one = dlwp.predict_timeseries(x0[5:6], 20)
    np.testing.assert_array_equal(y_one[:, 0], y[:, 5])
    eng = dlwp.model.engine(16)
    xd = torch.from_numpy(x0).cuda()
    s_graph = eng.rollout_device(xd, 20, use_graph=True).cpu().numpy()
    s_plain = eng.rollout_device(xd, 20, use_graph=False).cpu().numpy()
    np.testing.assert_array_equal(s_graph, s_plain)
    np.testing.assert_array_equal(s_graph, y)


def test_linear_net_is_linear(torch_cuda):
    """With linear activations and zero bias the rollout is a linear map: f(a*x + b*z) = a*f(x) + b*f(z)."""
    cf = 'channels_first'
    layers = (('PeriodicPadding2D', ((0, 2),), {'data_format': cf, 'input_shape': (6, 91, 180)}),
              ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
              ('Conv2D', (6, 5), {'activation': 'linear', 'data_format': cf, 'use_bias': False}))
    dlwp = build_product_sequential(layers)
    rng = np.random.RandomState(11)
    x, z = (rng.standard_normal((2, 6, 91, 180)).astype(np.float32) for _ in range(2))
    fx, fz = dlwp.predict_timeseries(x, 3), dlwp.predict_timeseries(z, 3)
    fxz = dlwp.predict_timeseries(2.0 * x - 0.5 * z, 3)
    assert rel_err(fxz, 2.0 * fx.astype(np.float64) - 0.5 * fz) < 1e-5


def test_predict_chunks_batches_larger_than_plan_capacity(torch_cuda):
    from dlwp_b200.engine import CompiledNet
    layers = _small_layers(4, 6, 8)
    dlwp = build_product_sequential(layers)
    net = oracle_sequential_like(dlwp, layers, seed=2)
    x = np.random.RandomState(1).standard_normal((70, 4, 6, 8)).astype(np.float32)
    eng = CompiledNet(dlwp.model, 16)            # a 16-sample plan fed 70 samples: 5 chunks
    assert eng.max_batch == 16
    assert rel_err(eng.predict(x)[0], net.forward(x.astype(np.float64))) < 2e-5
    assert rel_err(eng.rollout_host(x, 3), oracle_rollout64(net, x, 3)) < 5e-5
    eng.close()
