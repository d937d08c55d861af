"""
CPU tests of `DLWP.model.DLWPTorchNN` (reference DLWP/model/models_torch.py): the translation of torch.nn layer tuples into
the Keras-style stack the engine runs, the weight layouts, the argument validation, and the torch-Adam <-> Keras-Adam
epsilon mapping.  The arithmetic itself is the DLWPNeuralNet path (tests/test_rollout_gpu.py compares it with the
reference's own DLWPTorchNN output, tests/golden/torchnn_net_a.npz).
"""

import numpy as np
import pytest

from oracle import layers as OL


def _net_a_torch_layers(C=6):
    return (
        ('CircularPad2d', ((2, 2, 0, 0),), None),
        ('ZeroPad2d', ((0, 0, 2, 2),), None),
        ('Conv2d', (C, 32, 3), {'dilation': 2, 'activation': 'tanh'}),
        ('CircularPad2d', ((2, 2, 0, 0),), None),
        ('ZeroPad2d', ((0, 0, 2, 2),), None),
        ('Conv2d', (32, C, 5), None),
    )


def _build(layers, **kw):
    from dlwp_b200.model import DLWPTorchNN
    dlwp = DLWPTorchNN(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None, scale_targets=False)
    dlwp.build_model(layers, 'Adam', 'MSELoss', **kw)
    return dlwp


def _torch_forward(dlwp, x):
    """What the reference's DLWPTorchNN._forward does (models_torch.py:155-160), on the CPU parameter holders."""
    import torch
    with torch.no_grad():
        x = torch.from_numpy(x)
        for layer, act in zip(dlwp.layers, dlwp.activations):
            x = layer(x)
            if act is not None:
                x = act(x)
    return x.numpy()


def test_net_a_translates_to_the_keras_stack_of_the_benchmark():
    from dlwp_b200.engine import Lowering
    from dlwp_b200.model import DLWPNeuralNet
    dlwp = _build(_net_a_torch_layers())
    got = dlwp.keras_layers((6, 23, 36))
    names = [l[0] for l in got]
    assert names == ['PeriodicPadding2D', 'ZeroPadding2D', 'Conv2D', 'PeriodicPadding2D', 'ZeroPadding2D', 'Conv2D']
    assert got[0][1] == (((0, 0), (2, 2)),) and got[0][2]['input_shape'] == (6, 23, 36)
    assert got[1][1] == (((2, 2), (0, 0)),)
    assert got[2][1] == (32, (3, 3)) and got[2][2]['dilation_rate'] == (2, 2) and got[2][2]['activation'] == 'tanh'
    assert got[5][1] == (6, (5, 5)) and got[5][2]['activation'] == 'linear'
    # ... and lowers to the same two fused pad + conv ops as the Keras-style Net A
    a = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    a.build_model(got, loss='mse', optimizer='adam')
    b = DLWPNeuralNet(is_convolutional=True, scaler_type=None, scale_targets=False)
    b.build_model(OL.net_a_layers((6, 23, 36)), loss='mse', optimizer='adam')
    assert Lowering(a.model).ops == Lowering(b.model).ops


@pytest.mark.parametrize('layers,shape', [
    (_net_a_torch_layers(4), (4, 12, 16)),
    ((('Conv2d', (3, 8, 3), {'padding': 1, 'padding_mode': 'circular', 'activation': 'relu'}),
      ('MaxPool2d', (2,), None),
      ('Conv2d', (8, 8, 3), {'padding': (1, 1), 'activation': 'tanh', 'bias': False}),
      ('Upsample', (), {'scale_factor': 2}),
      ('ReplicationPad2d', ((1, 1, 0, 0),), None),
      ('ReflectionPad2d', ((0, 0, 1, 1),), None),
      ('Conv2d', (8, 3, 3), None)), (3, 8, 12)),
])
def test_translated_stack_and_weights_equal_the_torch_modules(layers, shape):
    """The torch modules' own forward (CPU) vs the oracle interpreter run on `keras_layers` + `keras_weights`: pins the layer
    mapping (padding orders, circular = periodic in both dimensions, pool / upsample) and the (O,I,kh,kw) -> (kh,kw,I,O)
    weight transposition that `predict` pushes into the engine."""
    import torch
    torch.manual_seed(3)
    dlwp = _build(layers)
    x = np.random.RandomState(0).standard_normal((2,) + shape).astype(np.float32)
    want = _torch_forward(dlwp, x)
    net = OL.OSequential(dlwp.keras_layers(shape))
    ws = dlwp.keras_weights()
    assert [w.shape for w in ws] == [w.shape for w in net.get_weights()]
    net.set_weights(ws)
    got = np.asarray(net.forward(x.astype(np.float64)))
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)


def test_build_model_validation_matches_the_reference_messages():
    from dlwp_b200.model import DLWPTorchNN
    with pytest.raises(ValueError, match="'time_dim' must be >= 1"):
        DLWPTorchNN(time_dim=0)
    dlwp = DLWPTorchNN(is_convolutional=True, scaler_type=None)
    with pytest.raises(TypeError, match="'layers' argument must be a tuple"):
        dlwp.build_model('Conv2d', 'Adam', 'MSELoss')
    with pytest.raises(TypeError, match="each element of 'layers' must be a tuple"):
        dlwp.build_model(('Conv2d',), 'Adam', 'MSELoss')
    with pytest.raises(ValueError, match='three elements'):
        dlwp.build_model((('Conv2d', (1, 1, 1)),), 'Adam', 'MSELoss')
    with pytest.raises(TypeError, match="'args' element of layer 0 must be a tuple"):
        dlwp.build_model((('Conv2d', [1, 1, 1], None),), 'Adam', 'MSELoss')
    with pytest.raises(TypeError, match="'kwargs' element of layer 0 must be a dict"):
        dlwp.build_model((('Conv2d', (1, 1, 1), ()),), 'Adam', 'MSELoss')
    with pytest.raises(TypeError, match="'optimizer_kwargs' must be a dict"):
        dlwp.build_model((('Conv2d', (1, 1, 1), None),), 'Adam', 'MSELoss', optimizer_kwargs=[1])
    for bad in ((('Linear', (4, 4), None),), (('Conv2d', (2, 2, 3), {'stride': 2}),),
                (('Conv2d', (2, 2, 3), {'activation': 'gelu'}),), (('MaxPool2d', (3,), None),)):
        with pytest.raises(NotImplementedError):
            dlwp.build_model(bad, 'Adam', 'MSELoss')
    good = _build(_net_a_torch_layers(), optimizer_kwargs={'lr': 3e-4})
    assert good.optimizer.param_groups[0]['lr'] == 3e-4 and good.activations[2] is not None and good.activations[5] is None
    with pytest.raises(RuntimeError):                      # no CPU forward: the torch modules only hold the parameters
        good.model(None)
    sgd = DLWPTorchNN(is_convolutional=True, scaler_type=None)
    sgd.build_model(_net_a_torch_layers(), 'SGD', 'MSELoss', optimizer_kwargs={'lr': 0.1})
    with pytest.raises(NotImplementedError):
        sgd.fit_generator([])
    with pytest.raises(ValueError, match='time_steps must be an int > 0'):
        good.predict_timeseries(np.zeros((1, 6, 8, 8), np.float32), 0)


@pytest.mark.parametrize('eps', [1e-8, 1e-3])
def test_torch_adam_equals_keras_adam_with_the_mapped_epsilon(eps):
    """torch.optim.Adam vs the Keras 2.2 update dlwp_train_adam implements (training.py / train.cu adam_kernel), with the
    per-step epsilon `DLWPTorchNN._torch_adam_step_constants` hands to it: identical trajectories."""
    import torch
    rng = np.random.RandomState(5)
    w0 = rng.standard_normal(7)
    grads = [rng.standard_normal(7) * s for s in (1.0, 1e-3, 1e-6, 2.0, 1e-8, 0.5)]
    lr, b1, b2 = 1e-2, 0.9, 0.999
    p = torch.tensor(w0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p], lr=lr, betas=(b1, b2), eps=eps)
    w, m, v = w0.copy(), np.zeros(7), np.zeros(7)
    for t, g in enumerate(grads, start=1):
        p.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
        eps_k = eps * np.sqrt(1.0 - b2 ** t)                              # the mapping under test
        lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        w = w - lr_t * m / (np.sqrt(v) + eps_k)
        np.testing.assert_allclose(w, p.detach().numpy(), rtol=1e-10, atol=1e-12)


def test_time_series_estimator_accepts_the_torch_twin():
    from dlwp_b200.model import ArraySeriesGenerator, TimeSeriesEstimator
    dlwp = _build((('Conv2d', (2, 2, 1), None),))
    data = np.zeros((6, 2, 4, 6), np.float32)
    times = np.datetime64('2003-03-01T00:00') + np.arange(6) * np.timedelta64(6, 'h')
    gen = ArraySeriesGenerator(data, times, np.linspace(60, -60, 4), np.arange(0, 360, 60.), ['a', 'b'])
    est = TimeSeriesEstimator(dlwp, gen)
    assert est._input_time_steps == 1 and not est._device_ok()


def test_engine_side_model_follows_the_torch_parameters():
    """`_ensure` builds the engine-side model once per input shape and re-pushes the weights whenever a torch parameter was
    modified in place (`layers[i].weight.copy_`, an optimizer step, `reset`); `_pull` writes engine weights back."""
    import torch
    dlwp = _build(_net_a_torch_layers())
    dlwp._ensure((6, 23, 36))
    model = dlwp._net.model
    for a, b in zip(model.get_weights(), dlwp.keras_weights()):
        np.testing.assert_array_equal(a, b)
    with torch.no_grad():
        dlwp.layers[2].weight.mul_(0.5)
    dlwp._ensure((6, 23, 36))
    assert dlwp._net.model is model                                   # same shape: not rebuilt
    np.testing.assert_array_equal(model.get_weights()[0], dlwp.keras_weights()[0])
    new = [w + 1.0 for w in model.get_weights()]
    model.set_weights(new)
    dlwp._pull()
    for a, b in zip(new, dlwp.keras_weights()):
        np.testing.assert_array_equal(a, b)
    dlwp._ensure((6, 30, 40))
    assert dlwp._net.model is not model and dlwp._net.model.input_shape == (None, 6, 30, 40)
    dlwp.reset()
    dlwp._ensure((6, 30, 40))
    np.testing.assert_array_equal(dlwp._net.model.get_weights()[0], dlwp.keras_weights()[0])
