"""
`build_model(..., gpus=N)` (DLWP/model/models.py:104-109: keras multi_gpu_model, a single-process batch split) on a box
with at least two GPUs: results equal the single-device ones bit for bit.  Skipped on one GPU.
"""

import numpy as np
import pytest

from oracle import layers as OL

pytestmark = pytest.mark.gpu


def test_build_model_gpus_2_splits_the_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from dlwp_b200.model import DLWPNeuralNet
    layers = OL.net_a_layers((6, 24, 48))
    one = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    one.build_model(layers, loss='mse', optimizer='adam')
    two = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    two.build_model(layers, gpus=2, loss='mse', optimizer='adam')
    two.model.set_weights(one.model.get_weights())
    assert two.gpus == 2 and two.model is two.base_model and two.model._gpus == 2
    x = np.random.RandomState(0).standard_normal((7, 6, 24, 48)).astype(np.float32)
    np.testing.assert_array_equal(two.predict(x), one.predict(x))
    np.testing.assert_array_equal(two.predict_timeseries(x, 5), one.predict_timeseries(x, 5))
    assert sorted(two.model._replicas) == [0, 1]
    three = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    three.build_model(layers, gpus=torch.cuda.device_count() + 1, loss='mse', optimizer='adam')
    with pytest.raises(ValueError):
        three.predict(x)
