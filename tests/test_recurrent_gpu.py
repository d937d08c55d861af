"""
GPU parity of the recurrent front block (PeriodicPadding3D + ZeroPadding3D + ConvLSTM2D,
examples/train.py:144-157) and of the stand-alone padding modes (FillPadding2D / TFPadding2D), through the C ABI and
through the reference-facing API, against the oracle and the reference-generated goldens.
"""

import ctypes
import os

import numpy as np
import pytest

from oracle import layers as OL
from oracle import ops as OO
from oracle import rollout as OR
from tests.helpers import build_product_sequential, rel_err
from tests.test_oracle_golden import small_recurrent_layers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    import torch
    from dlwp_b200 import _native
    _native.lib()
    return _native, torch


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize('ract', ['hard_sigmoid', 'sigmoid'])
@pytest.mark.parametrize('first', [True, False])
def test_convlstm_gates_match_cell_equations(env, ract, first):
    """dlwp_convlstm_gates vs keras ConvLSTM2DCell.call restated in float64."""
    nat, torch = env
    rng = np.random.RandomState(5)
    N, F, H, W = 3, 5, 7, 9
    z = (2 * rng.standard_normal((N, 4 * F, H, W))).astype(np.float32)
    r = (2 * rng.standard_normal((N, 4 * F, H, W))).astype(np.float32)
    c = rng.standard_normal((N, F, H, W)).astype(np.float32)
    s = OO.hard_sigmoid if ract == 'hard_sigmoid' else (lambda v: 1 / (1 + np.exp(-v)))
    zz = z.astype(np.float64) + (0 if first else r.astype(np.float64))
    cp = 0 if first else c.astype(np.float64)
    c_ref = s(zz[:, F:2 * F]) * cp + s(zz[:, :F]) * np.tanh(zz[:, 2 * F:3 * F])
    h_ref = s(zz[:, 3 * F:]) * np.tanh(c_ref)
    zd, rd, cd = (torch.from_numpy(a).cuda() for a in (z, r, c))
    hd = torch.full((N, F, H, W), float('nan'), device='cuda')
    rc = nat.lib().dlwp_convlstm_gates(zd.data_ptr(), None if first else rd.data_ptr(), None if first else cd.data_ptr(),
                                       cd.data_ptr(), hd.data_ptr(), N, F, H, W, 4 * F * H * W, 4 * F * H * W, F * H * W,
                                       F * H * W, nat.ACT_TANH, nat.RECURRENT_ACTIVATIONS[ract], 0, 0, _stream(torch))
    nat.check(rc, 'dlwp_convlstm_gates')
    torch.cuda.synchronize()
    assert np.abs(cd.cpu().numpy() - c_ref).max() < 2e-6
    assert np.abs(hd.cpu().numpy() - h_ref).max() < 2e-6


def _recurrent_pair(layers, time_dim, seed=7):
    dlwp = build_product_sequential(layers, time_dim=time_dim, is_recurrent=True)
    net = OL.OSequential(layers)
    rng = np.random.RandomState(seed)
    for layer in net.weight_layers:
        if isinstance(layer, OL.OConvLSTM2D):
            layer.randomize(rng, bias_scale=0.1)
    OL.init_weights(net.conv_layers, seed=seed + 1, bias_scale=0.05)
    dlwp.model.set_weights(net.get_weights())
    return dlwp, net


@pytest.mark.parametrize('time_dim', [2, 3])
def test_recurrent_rollout_flags_match_reference_loop_goldens(env, golden_dir, time_dim):
    """Same nets / inputs as tests/golden/rollout_recurrent.npz: the reference's own is_recurrent loop (models.py:270-301)."""
    g = np.load(os.path.join(golden_dir, 'rollout_recurrent.npz'))
    dlwp = build_product_sequential(small_recurrent_layers(time_dim), time_dim=time_dim, is_recurrent=True)
    dlwp.model.set_weights([g['w_td%d_%d' % (time_dim, k)] for k in range(5)])
    x0 = g['x0_td%d' % time_dim]
    for steps in (1, 5):
        for ss in (False, True):
            for ktd in (False, True):
                key = 'y_td%d_s%d_ss%d_k%d' % (time_dim, steps, ss, ktd)
                y = dlwp.predict_timeseries(x0, steps, step_sequence=ss, keep_time_dim=ktd)
                assert y.shape == g[key].shape, key
                assert rel_err(y, g[key].astype(np.float64)) < 2e-5, key


def test_example_recurrent_net_predict_and_rollout(env):
    """The layer tuples of examples/train.py:140-220 (ConvLSTM2D front block + up-sampling conv stack), small grid."""
    cf = 'channels_first'
    cs = cso = (2, 3, 16, 32)
    conv = lambda f, k, d, act='tanh': ('Conv2D', (f, k), {'dilation_rate': d, 'padding': 'valid', 'activation': act,
                                                           'data_format': cf})
    pp = lambda p: ('PeriodicPadding2D', ((0, p),), {'data_format': cf})
    zp = lambda p: ('ZeroPadding2D', ((p, 0),), {'data_format': cf})
    layers = (
        ('PeriodicPadding3D', ((0, 0, 2),), {'data_format': cf, 'input_shape': cs}),
        ('ZeroPadding3D', ((0, 2, 0),), {'data_format': cf}),
        ('ConvLSTM2D', (4 * cs[1], 3), {'dilation_rate': 2, 'padding': 'valid', 'data_format': cf, 'activation': 'tanh',
                                        'return_sequences': True}),
        ('Reshape', ((4 * cs[0] * cs[1], cs[2], cs[3]),), None),
        pp(2), zp(2), conv(32, 3, 2),
        ('MaxPooling2D', (2,), {'data_format': cf}),
        pp(1), zp(1), conv(64, 3, 1),
        ('UpSampling2D', (2,), {'data_format': cf}),
        pp(2), zp(2), conv(32, 3, 2),
        pp(2), zp(2), conv(cso[0] * cso[1], 5, 1, 'linear'),
        ('Reshape', (cso,), None))
    dlwp, net = _recurrent_pair(layers, time_dim=2)
    assert dlwp.model.output_shape == (None,) + cso
    x0 = np.random.RandomState(8).standard_normal((3,) + cs).astype(np.float32)
    y = dlwp.predict(x0)
    ref = net.forward(x0.astype(np.float64))
    assert y.shape == ref.shape == (3,) + cso
    assert rel_err(y, ref) < 2e-5
    got = dlwp.predict_timeseries(x0, 8)          # 4 applications of a time_dim = 2 model, device-resident rollout
    ref_ts = OR.neuralnet_predict_timeseries(lambda p: net.forward(p), x0.astype(np.float64), 8, time_dim=2,
                                             is_recurrent=True, dtype=np.float64)
    assert got.shape == ref_ts.shape == (8, 3, 3, 16, 32)
    assert rel_err(got, ref_ts) < 1e-4


def test_convlstm_last_state_only_and_same_padding(env):
    """return_sequences=False (the layer's last h only) and padding='same' on the input convolution."""
    cf = 'channels_first'
    cs = (3, 4, 10, 12)
    layers = (('ConvLSTM2D', (6, 3), {'padding': 'same', 'data_format': cf, 'activation': 'tanh', 'input_shape': cs,
                                      'recurrent_activation': 'sigmoid', 'return_sequences': False}),
              ('PeriodicPadding2D', ((0, 1),), {'data_format': cf}),
              ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
              ('Conv2D', (4, 3), {'activation': 'linear', 'data_format': cf}))
    dlwp, net = _recurrent_pair(layers, time_dim=3)
    x0 = np.random.RandomState(9).standard_normal((2,) + cs).astype(np.float32)
    y = dlwp.predict(x0)
    ref = net.forward(x0.astype(np.float64))
    assert y.shape == ref.shape == (2, 4, 10, 12)
    assert rel_err(y, ref) < 2e-5


@pytest.mark.parametrize('mode', ['edge', 'reflect', 'symmetric'])
def test_standalone_pad_modes_bit_exact(env, golden_dir, mode):
    """dlwp_pad2d with FillPadding2D / tf.pad REFLECT / SYMMETRIC index maps: pure data movement, bit exact."""
    nat, torch = env
    code = {'edge': nat.PAD_EDGE, 'reflect': nat.PAD_REFLECT, 'symmetric': nat.PAD_SYMMETRIC}[mode]
    g = np.load(os.path.join(golden_dir, 'padding3d_fill2d.npz'))
    x = g['x_channels_first']
    N, C, H, W = x.shape
    xd = torch.from_numpy(x).cuda()
    for k in range(int(g['n2'])):
        (t, b), (l, r) = (tuple(int(v) for v in row) for row in g['pad2_%d' % k])
        yd = torch.full((N, C, H + t + b, W + l + r), float('nan'), device='cuda')
        Ho, Wo = H + t + b, W + l + r
        rc = nat.lib().dlwp_pad2d(xd.data_ptr(), yd.data_ptr(), N, C, H, W, t, b, l, r, code, code, C * H * W, H * W, W,
                                  C * Ho * Wo, Ho * Wo, Wo, _stream(torch))
        nat.check(rc, 'dlwp_pad2d')
        got = yd.cpu().numpy()
        if mode == 'edge':
            np.testing.assert_array_equal(got, g['yfill_%d_channels_first' % k])   # the reference's FillPadding2D.call
        else:
            np.testing.assert_array_equal(got, OO.tf_pad2d(x, ((t, b), (l, r)), mode.upper()))


def test_model_with_fill_and_tf_padding_layers(env):
    """FillPadding2D / TFPadding2D in a layer-tuple model: the paddings run as pad ops in front of the conv."""
    cf = 'channels_first'
    layers = (('FillPadding2D', ((1, 2),), {'data_format': cf, 'input_shape': (3, 9, 12)}),
              ('Conv2D', (5, 3), {'activation': 'tanh', 'data_format': cf}),
              ('TFPadding2D', ((1, 1),), {'data_format': cf, 'mode': 'REFLECT'}),
              ('PeriodicPadding2D', ((0, 1),), {'data_format': cf}),
              ('Conv2D', (3, 3), {'activation': 'linear', 'data_format': cf}))
    dlwp = build_product_sequential(layers)
    rng = np.random.RandomState(4)
    ws = [OO.glorot_uniform(rng, 3, 3, 3, 5), (0.1 * rng.standard_normal(5)).astype(np.float32),
          OO.glorot_uniform(rng, 3, 3, 5, 3), (0.1 * rng.standard_normal(3)).astype(np.float32)]
    dlwp.model.set_weights(ws)
    x = rng.standard_normal((2, 3, 9, 12)).astype(np.float32)
    y = dlwp.predict(x)
    x64 = x.astype(np.float64)
    h = np.tanh(OO.conv2d_valid(OO.fill_pad2d(x64, ((1, 1), (2, 2))), ws[0].astype(np.float64), ws[1].astype(np.float64)))
    h = OO.periodic_pad2d(OO.tf_pad2d(h, ((1, 1), (1, 1)), 'REFLECT'), ((0, 0), (1, 1)))
    ref = OO.conv2d_valid(h, ws[2].astype(np.float64), ws[3].astype(np.float64))
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 2e-5
