"""
data_format='channels_last' models (the Keras default; DLWP/custom.py:205-213 is that branch of PeriodicPadding2D): the plan
computes channels_first between one layout op at each end.  predict / rollout against the oracle run in channels_last, and
the stand-alone layout operator against numpy.
"""

import ctypes

import numpy as np
import pytest

from oracle import layers as OL
from oracle import rollout as OR
from tests.helpers import build_product_sequential, rel_err

pytestmark = pytest.mark.gpu


def _layers(shape):
    cl = 'channels_last'
    return (('PeriodicPadding2D', ((0, 2),), {'data_format': cl, 'input_shape': shape}),
            ('ZeroPadding2D', ((2, 0),), {'data_format': cl}),
            ('Conv2D', (16, 3), {'dilation_rate': 2, 'activation': 'tanh', 'data_format': cl}),
            ('MaxPooling2D', (2,), {'data_format': cl}),
            ('UpSampling2D', (2,), {'data_format': cl}),
            ('PeriodicPadding2D', ((0, 2),), {}),                     # data_format omitted: the Keras default
            ('ZeroPadding2D', ((2, 0),), None),
            ('Conv2D', (shape[2], 5), {'activation': 'linear'}))


def test_channels_last_model_predict_and_rollout():
    import torch  # noqa: F401
    shape = (12, 20, 5)                                               # (H, W, C)
    layers = _layers(shape)
    dlwp = build_product_sequential(layers)
    net = OL.OSequential(layers)
    OL.init_weights(net.conv_layers, seed=5, bias_scale=0.05)
    dlwp.model.set_weights(net.get_weights())
    assert dlwp.model.output_shape == (None,) + shape
    x0 = np.random.RandomState(3).standard_normal((3,) + shape).astype(np.float32)
    y = dlwp.predict(x0)
    ref = net.forward(x0.astype(np.float64))
    assert y.shape == ref.shape == (3,) + shape
    assert rel_err(y, ref) < 2e-5
    got = dlwp.predict_timeseries(x0, 6)
    ref_ts = OR.neuralnet_predict_timeseries(lambda p: net.forward(p), x0.astype(np.float64), 6, dtype=np.float64)
    assert got.shape == ref_ts.shape and rel_err(got, ref_ts) < 1e-4


def test_layout_operator_round_trip():
    import torch
    from dlwp_b200 import _native as nat
    rng = np.random.RandomState(1)
    N, C, H, W = 3, 5, 7, 11
    x = rng.standard_normal((N, H, W, C)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty((N, C, H, W), device='cuda')
    zd = torch.empty((N, H, W, C), device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(nat.lib().dlwp_layout2d(xd.data_ptr(), yd.data_ptr(), N, C, H, W, 1, st))
    nat.check(nat.lib().dlwp_layout2d(yd.data_ptr(), zd.data_ptr(), N, C, H, W, 0, st))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(yd.cpu().numpy(), np.transpose(x, (0, 3, 1, 2)))
    np.testing.assert_array_equal(zd.cpu().numpy(), x)
