"""
`DLWP.model.DLWPTorchNN` on the GPU engine vs the REFERENCE's own DLWPTorchNN (DLWP/model/models_torch.py run on CPU by
tests/golden/make_golden.py:gen_torchnn -> torchnn_net_a.npz): the same build_model call, the same in-place weight
assignment through `dlwp.layers[i].weight`, `predict_timeseries` within the fp32 rollout tolerance; and two Adam steps of
`fit_generator` vs torch's own autograd + torch.optim.Adam on the same modules.
"""

import copy
import os

import numpy as np
import pytest

from oracle import ops as OO
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _layers(C):
    return (('CircularPad2d', ((2, 2, 0, 0),), None),
            ('ZeroPad2d', ((0, 0, 2, 2),), None),
            ('Conv2d', (C, 32, 3), {'dilation': 2, 'activation': 'tanh'}),
            ('CircularPad2d', ((2, 2, 0, 0),), None),
            ('ZeroPad2d', ((0, 0, 2, 2),), None),
            ('Conv2d', (32, C, 5), None))


def test_predict_timeseries_matches_the_reference_torch_twin():
    import torch
    from dlwp_b200.model import DLWPTorchNN
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'torchnn_net_a.npz'))
    tag = 'small'
    x0 = g[tag + '_x0']
    C = x0.shape[1]
    dlwp = DLWPTorchNN(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None, scale_targets=False)
    dlwp.build_model(_layers(C), 'Adam', 'MSELoss')
    with torch.no_grad():                                    # exactly what make_golden.py does to the reference object
        dlwp.layers[2].weight.copy_(torch.from_numpy(np.transpose(g[tag + '_k1'], (3, 2, 0, 1)).copy()))
        dlwp.layers[2].bias.copy_(torch.from_numpy(g[tag + '_b1']))
        dlwp.layers[5].weight.copy_(torch.from_numpy(np.transpose(g[tag + '_k2'], (3, 2, 0, 1)).copy()))
        dlwp.layers[5].bias.copy_(torch.from_numpy(g[tag + '_b2']))
    steps, sub = int(g[tag + '_steps']), int(g[tag + '_sub'])
    y = dlwp.predict_timeseries(x0, steps)
    assert y.shape == (steps,) + x0.shape
    assert rel_err(y[:, :, :, ::sub, ::sub], g[tag + '_y'].astype(np.float64)) < 1e-5
    one = dlwp.predict(x0)
    assert rel_err(one, y[0].astype(np.float64)) < 1e-6


def test_fit_generator_follows_torch_adam():
    import torch
    from dlwp_b200.model import DLWPTorchNN
    torch.manual_seed(0)
    C, H, W = 4, 10, 16
    dlwp = DLWPTorchNN(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None, scale_targets=False)
    dlwp.build_model(_layers(C), 'Adam', 'MSELoss', optimizer_kwargs={'lr': 1e-3, 'eps': 1e-3})
    rng = np.random.RandomState(1)
    batches = [(rng.standard_normal((3, C, H, W)).astype(np.float32), rng.standard_normal((3, C, H, W)).astype(np.float32))
               for _ in range(2)]
    # torch's own training of a copy of the modules (what the reference's fit_generator does, models_torch.py:248-262)
    mods = [copy.deepcopy(m) for m in dlwp.layers]
    params = [p for m in mods for p in m.parameters()]
    # (eps = 1e-3: weights whose gradient is ~0 get a damped, noise-insensitive update, and the placement of eps -- where
    # torch's and Keras' Adam differ -- changes the result by far more than the tolerance below)
    opt = torch.optim.Adam(params, lr=1e-3, eps=1e-3)
    ref_losses = []
    for p, t in batches:
        opt.zero_grad()
        x = torch.from_numpy(p)
        for m, act in zip(mods, dlwp.activations):
            x = m(x)
            if act is not None:
                x = act(x)
        loss = torch.nn.functional.mse_loss(x, torch.from_numpy(t))
        loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    hist = dlwp.fit_generator(batches, epochs=1)
    assert abs(hist['loss'][0] - np.mean(ref_losses)) < 1e-5 * max(1.0, abs(np.mean(ref_losses)))
    assert len(hist['error']) == 1 and np.isfinite(hist['error'][0])
    for mine, ref in zip([p for m in dlwp.layers for p in m.parameters()], params):
        a, b = mine.detach().numpy(), ref.detach().numpy()
        assert np.abs(a - b).max() < 2e-5, np.abs(a - b).max()    # two steps of lr 1e-3: a wrong update would be off by ~1e-3
    loss, err = dlwp.evaluate(*batches[0])
    assert np.isfinite(loss) and np.isfinite(err)
