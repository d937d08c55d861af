"""
The oracle against the golden vectors produced by the reference's own code (tests/golden/make_golden.py), and the
oracle's independent formulations against each other.  CPU only.
"""

import os

import numpy as np
import pytest

from oracle import layers as OL
from oracle import ops as OO
from oracle import rollout as OR


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_periodic_padding_matches_reference_call(golden_dir):
    g = _load(golden_dir, 'periodic_padding2d.npz')
    for k in range(int(g['n_cases'])):
        pad = tuple(tuple(int(v) for v in row) for row in g['pad_%d' % k])
        for fmt in ('channels_first', 'channels_last'):
            y = OO.periodic_pad2d(g['x_' + fmt], pad, fmt)
            assert y.shape == g['y_%d_%s' % (k, fmt)].shape
            np.testing.assert_array_equal(y, g['y_%d_%s' % (k, fmt)])  # pure data movement: bit exact


def test_row_conv2d_matches_reference(golden_dir):
    g = _load(golden_dir, 'row_conv2d.npz')
    y = OO.row_conv2d(g['x'], g['kernel'])
    np.testing.assert_allclose(y, g['y'], rtol=0, atol=1e-12)
    y_cl = OO.row_conv2d(np.moveaxis(g['x'], 1, 3), g['kernel'], data_format='channels_last')
    np.testing.assert_allclose(y_cl, g['y_channels_last'], rtol=0, atol=1e-12)


def _small_seq(time_dim, ws):
    cf = 'channels_first'
    C = time_dim * 2
    net = OL.OSequential((
        ('PeriodicPadding2D', ((0, 1),), {'data_format': cf, 'input_shape': (C, 6, 8)}),
        ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
        ('Conv2D', (8, 3), {'activation': 'tanh', 'data_format': cf}),
        ('PeriodicPadding2D', ((0, 2),), {'data_format': cf}),
        ('ZeroPadding2D', ((2, 0),), {'data_format': cf}),
        ('Conv2D', (C, 3), {'dilation_rate': 2, 'activation': 'linear', 'data_format': cf}),
    ))
    net.set_weights(ws)
    return net


def test_neuralnet_rollout_matches_reference_loop(golden_dir):
    g = _load(golden_dir, 'rollout_neuralnet.npz')
    for key in g['cases']:
        key = str(key)
        td, steps, ss, ktd = (int(p[len(pre):]) for p, pre in zip(key.split('_')[1:], ('td', 's', 'ss', 'k')))
        net = _small_seq(td, [g['w_td%d_%d' % (td, k)] for k in range(4)])
        fn = lambda p: net.forward(np.asarray(p, np.float64)).astype(np.float32)
        y = OR.neuralnet_predict_timeseries(fn, g['x0_td%d' % td], steps, time_dim=td, step_sequence=bool(ss),
                                            keep_time_dim=bool(ktd))
        assert y.shape == g[key].shape, key
        assert y.dtype == np.float32
        np.testing.assert_array_equal(y, g[key], err_msg=key)


def test_functional_rollout_matches_reference_loop(golden_dir):
    g = _load(golden_dir, 'rollout_functional.npz')
    for key in g['cases']:
        key = str(key)
        td, n, steps, ktd = (int(p[len(pre):]) for p, pre in zip(key.split('_')[1:], ('td', 'n', 's', 'k')))
        net = _small_seq(td, [g['w_td%d_%d' % (td, k)] for k in range(4)])

        def fn(p):
            outs = [net.forward(np.asarray(p, np.float64))]
            for _ in range(1, n):
                outs.append(net.forward(outs[-1]))
            outs = [o.astype(np.float32) for o in outs]
            return outs[0] if n == 1 else outs
        y = OR.functional_predict_timeseries(fn, g['x0_td%d' % td], steps, n_steps=n, time_dim=td,
                                             keep_time_dim=bool(ktd))
        assert y.shape == g[key].shape, key
        np.testing.assert_array_equal(y, g[key], err_msg=key)


@pytest.mark.parametrize('tag', ['small', 'full'])
def test_net_a_rollout_matches_reference_torch_twin(golden_dir, tag):
    """End to end (periodic pad + zero pad + dilated conv + tanh + feedback loop) vs DLWPTorchNN run on CPU."""
    g = _load(golden_dir, 'torchnn_net_a.npz')
    x0 = g[tag + '_x0']
    net = OL.OSequential(OL.net_a_layers(x0.shape[1:]))
    net.set_weights([g[tag + '_k1'], g[tag + '_b1'], g[tag + '_k2'], g[tag + '_b2']])
    sub = int(g[tag + '_sub'])
    steps = int(g[tag + '_steps'])
    y64 = OR.neuralnet_predict_timeseries(lambda p: net.forward(p), x0.astype(np.float64), steps, dtype=np.float64)
    ref = g[tag + '_y']
    got = y64[:, :, :, ::sub, ::sub]
    assert got.shape == ref.shape
    rel = np.abs(got - ref).max() / np.abs(got).max()
    assert rel < 2e-6, rel  # fp32 reference vs fp64 oracle over 10 feedback steps


def test_conv_closed_form_equals_pad_then_conv():
    rng = np.random.RandomState(0)
    x = rng.standard_normal((2, 5, 9, 12))
    for (kh, kw, d, ph, pw, mh, mw) in [(3, 3, 2, (2, 2), (2, 2), 'zero', 'periodic'),
                                        (5, 5, 1, (2, 2), (2, 2), 'zero', 'periodic'),
                                        (3, 3, 1, (1, 1), (1, 1), 'periodic', 'periodic'),
                                        (3, 5, 1, (0, 2), (3, 1), 'zero', 'zero'),
                                        (3, 3, 1, (2, 0), (0, 2), 'periodic', 'zero')]:
        k = rng.standard_normal((kh, kw, 5, 4))
        b = rng.standard_normal(4)
        xp = x
        xp = OO.periodic_pad2d(xp, ((0, 0), pw)) if mw == 'periodic' else OO.zero_pad2d(xp, ((0, 0), pw))
        xp = OO.periodic_pad2d(xp, (ph, (0, 0))) if mh == 'periodic' else OO.zero_pad2d(xp, (ph, (0, 0)))
        y1 = OO.conv2d_valid(xp, k, b, (d, d))
        y2 = OO.pad_conv2d_closed_form(x, k, b, (d, d), ph, pw, mh, mw)
        np.testing.assert_allclose(y1, y2, rtol=0, atol=1e-12)


def test_conv_matches_torch_conv2d():
    """Keras Conv2D cannot run offline; its cross-correlation semantics are cross-checked against torch (independent)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.RandomState(1)
    x = rng.standard_normal((2, 6, 11, 14))
    for kh, kw, d in [(3, 3, 1), (3, 3, 2), (5, 5, 1), (1, 3, 1)]:
        k = rng.standard_normal((kh, kw, 6, 7))
        b = rng.standard_normal(7)
        y = OO.conv2d_valid(x, k, b, (d, d))
        yt = F.conv2d(torch.from_numpy(x), torch.from_numpy(np.transpose(k, (3, 2, 0, 1)).copy()),
                      torch.from_numpy(b), dilation=d).numpy()
        np.testing.assert_allclose(y, yt, rtol=0, atol=1e-11)


def test_torch_tier1_equals_numpy_tier0():
    import torch
    net = OL.OFunctionalNet((4, 16, 24), skip_connections=True, integration_steps=2)
    OL.init_weights(net.conv_layers, seed=3, bias_scale=0.1)
    x = np.random.RandomState(2).standard_normal((2, 4, 16, 24))
    y64 = net.forward(x)
    with torch.no_grad():
        y32 = net.forward(torch.from_numpy(x.astype(np.float32)))
    for a, b in zip(y64, y32):
        assert np.abs(a - b.numpy()).max() / np.abs(a).max() < 1e-5
    netb = OL.OFunctionalNet((4, 16, 24), skip_connections=False)
    OL.init_weights(netb.conv_layers, seed=4)
    yb = netb.forward(x)
    assert yb.shape == (2, 4, 16, 24)


def test_pool_upsample_slice_semantics():
    x = np.arange(2 * 2 * 5 * 6, dtype=np.float64).reshape(2, 2, 5, 6)
    p = OO.max_pool2d(x, 2)
    assert p.shape == (2, 2, 2, 3)          # floor: the odd 5th row is dropped
    assert p[0, 0, 0, 0] == x[0, 0, :2, :2].max()
    u = OO.upsample2d(p, 2)
    assert u.shape == (2, 2, 4, 6) and u[1, 1, 3, 5] == p[1, 1, 1, 2] and u[0, 0, 0, 1] == p[0, 0, 0, 0]
    assert OO.slice_channels(x, 1, 2).shape == (2, 1, 5, 6)
    with pytest.raises(ValueError):
        OO.slice_channels(x, 0, 1, axis=-1)


def test_rollout_argument_errors():
    with pytest.raises(ValueError):
        OR.neuralnet_predict_timeseries(lambda p: p, np.zeros((1, 2, 3, 4), np.float32), 0)
    with pytest.raises(ValueError):
        OR.functional_predict_timeseries(lambda p: p, np.zeros((1, 2, 3, 4), np.float32), -1)


def test_periodic_padding3d_and_fill_padding_match_reference_call(golden_dir):
    """PeriodicPadding3D.call (custom.py:277-306) and FillPadding2D.call (custom.py:359-402), run in place."""
    g = _load(golden_dir, 'padding3d_fill2d.npz')
    for k in range(int(g['n3'])):
        pad = tuple(tuple(int(v) for v in row) for row in g['pad3_%d' % k])
        for fmt in ('channels_first', 'channels_last'):
            np.testing.assert_array_equal(OO.periodic_pad3d(g['x5_' + fmt], pad, fmt), g['y3_%d_%s' % (k, fmt)])
    for k in range(int(g['n2'])):
        pad = tuple(tuple(int(v) for v in row) for row in g['pad2_%d' % k])
        for fmt in ('channels_first', 'channels_last'):
            np.testing.assert_array_equal(OO.fill_pad2d(g['x_' + fmt], pad, fmt), g['yfill_%d_%s' % (k, fmt)])


def small_recurrent_layers(time_dim, nvar=2, H=6, W=8):
    """The net of tests/golden/make_golden.py:_small_recurrent (ConvLSTM2D front block of examples/train.py:144-157)."""
    cf = 'channels_first'
    cs = (time_dim, nvar, H, W)
    return (('PeriodicPadding3D', ((0, 0, 2),), {'data_format': cf, 'input_shape': cs}),
            ('ZeroPadding3D', ((0, 2, 0),), {'data_format': cf}),
            ('ConvLSTM2D', (2 * nvar, 3), {'dilation_rate': 2, 'padding': 'valid', 'data_format': cf,
                                           'activation': 'tanh', 'return_sequences': True}),
            ('Reshape', ((2 * time_dim * nvar, H, W),), None),
            ('PeriodicPadding2D', ((0, 1),), {'data_format': cf}),
            ('ZeroPadding2D', ((1, 0),), {'data_format': cf}),
            ('Conv2D', (time_dim * nvar, 3), {'activation': 'linear', 'data_format': cf}),
            ('Reshape', (cs,), None))


def test_recurrent_rollout_matches_reference_loop(golden_dir):
    """is_recurrent=True branch of models.py:270-301 (5-D predictors) around the ConvLSTM2D-fronted oracle net."""
    g = _load(golden_dir, 'rollout_recurrent.npz')
    for key in g['cases']:
        key = str(key)
        td, steps, ss, ktd = (int(p[len(pre):]) for p, pre in zip(key.split('_')[1:], ('td', 's', 'ss', 'k')))
        net = OL.OSequential(small_recurrent_layers(td))
        net.set_weights([g['w_td%d_%d' % (td, k)] for k in range(5)])
        fn = lambda p: net.forward(np.asarray(p, np.float64)).astype(np.float32)
        y = OR.neuralnet_predict_timeseries(fn, g['x0_td%d' % td], steps, time_dim=td, is_recurrent=True,
                                            step_sequence=bool(ss), keep_time_dim=bool(ktd))
        assert y.shape == g[key].shape, key
        np.testing.assert_array_equal(y, g[key], err_msg=key)


def test_conv_lstm_first_step_and_gate_algebra():
    """ConvLSTM2D restatement: with a zero recurrent kernel every step is the stateless first step with the carried cell
    state; the first step equals the closed form o * tanh(i * tanh(z_c)) of the four gate convolutions."""
    rng = np.random.RandomState(3)
    x = rng.standard_normal((2, 3, 2, 7, 9))
    k = 0.3 * rng.standard_normal((3, 3, 2, 8))
    u = 0.3 * rng.standard_normal((3, 3, 2, 8))
    b = 0.1 * rng.standard_normal(8)
    y = OO.conv_lstm2d(x, k, u, b, padding='same')
    assert y.shape == (2, 3, 2, 7, 9)
    z = OO.conv2d_valid(np.pad(x[:, 0], [(0, 0), (0, 0), (1, 1), (1, 1)]), k, b)
    i, c, o = OO.hard_sigmoid(z[:, :2]), np.tanh(z[:, 4:6]), OO.hard_sigmoid(z[:, 6:])
    np.testing.assert_allclose(y[:, 0], o * np.tanh(i * c), atol=1e-14)
    y0 = OO.conv_lstm2d(x, k, 0 * u, b, padding='same', return_sequences=False)
    z2 = [OO.conv2d_valid(np.pad(x[:, t], [(0, 0), (0, 0), (1, 1), (1, 1)]), k, b) for t in range(3)]
    cst = 0
    for zt in z2:
        cst = OO.hard_sigmoid(zt[:, 2:4]) * cst + OO.hard_sigmoid(zt[:, :2]) * np.tanh(zt[:, 4:6])
    np.testing.assert_allclose(y0, OO.hard_sigmoid(z2[-1][:, 6:]) * np.tanh(cst), atol=1e-14)


def test_latitude_weighted_loss_matches_reference_function(golden_dir):
    """custom.py:956-991 executed from the reference module with an eagerly assigning `K.zeros` (the documented intent; see
    DESIGN.md section 7 on graph-mode TF1): cosine and mid-latitude weights broadcast over (H, W), and the no-latitude case.
    The (H, W) weight map the training kernel uses (`training._loss_weight_map`) comes from the same object."""
    from dlwp_b200.custom import latitude_weighted_loss
    from dlwp_b200.keras.losses import mean_squared_error
    g = _load(golden_dir, 'lat_loss.npz')
    for weighting in ('cosine', 'midlatitude'):
        fn = latitude_weighted_loss(mean_squared_error, g['lats'], (3, 8, 10), axis=-2, weighting=weighting)
        np.testing.assert_allclose(fn(g['y_true'], g['y_pred']), g['loss_' + weighting], rtol=0, atol=2e-7)
        assert np.asarray(fn.weights).shape == (8, 10)
    fn = latitude_weighted_loss(mean_squared_error, None, (3, 8, 10))
    np.testing.assert_allclose(fn(g['y_true'], g['y_pred']), g['loss_none'], rtol=0, atol=2e-7)


def test_anomaly_correlation_loss_matches_reference_function(golden_dir):
    """DLWP/custom.py:994-1088 executed from the reference module on a numpy backend (make_golden.py:gen_acc_loss): the
    product's host-side loss objects reproduce it for every regularize_mean, with and without a climatology.  The device
    loss of the training path is checked against these objects in tests/test_training_gpu.py."""
    from dlwp_b200.custom import anomaly_correlation, anomaly_correlation_loss
    g = _load(golden_dir, 'acc_loss.npz')
    yt, yp, mean = g['y_true'], g['y_pred'], g['mean']
    for reg in (None, 'mse', 'mae', 'global', 'spatial'):
        for use_mean in (False, True):
            fn = anomaly_correlation_loss(mean=mean if use_mean else None, regularize_mean=reg)
            want = g['loss_%s_%d' % (reg, use_mean)]
            got = np.asarray(fn(yt, yp))
            assert got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(anomaly_correlation(yt, yp, regularize_mean=reg), g['metric_%s' % reg], rtol=1e-12)
    fwd = anomaly_correlation_loss(regularize_mean=None, reverse=False)
    np.testing.assert_allclose(fwd(yt, yp), g['loss_none_forward'], rtol=1e-12)


def _estimator_cases(golden_dir):
    g = _load(golden_dir, 'estimator.npz')
    for key in [str(c) for c in g['cases']]:
        t_in, t_out, interval, sol, steps, impute, first, keep_time = [int(v) for v in g[key + '/spec']]
        yield g, key, t_in, t_out, interval, bool(sol), steps, bool(impute), bool(first), bool(keep_time)


def test_estimator_loop_matches_reference_loop(golden_dir):
    """oracle/estimator.py vs the reference's own TimeSeriesEstimator.predict (extensions.py:136-303) executed on a numpy
    stand-in for xarray (make_golden.py:gen_estimator): forecast values including the NaN pattern of samples that run out
    of data, for equal / fewer / more output time steps, insolation, impute, interval 2, keep_time_dim."""
    from oracle import estimator as OE
    for g, key, t_in, t_out, interval, sol, steps, impute, first, keep_time in _estimator_cases(golden_dir):
        w = g[key + '/w']
        fn = lambda x: np.tanh(np.einsum('nchw,co->nohw', np.asarray(x, np.float32), w)).astype(np.float32)
        times = g['times'].astype('datetime64[s]')
        S = g[key + '/p'].shape[0]
        res, es, keep = OE.estimator_predict(fn, g[key + '/p'], steps, t_in, t_out, list(g[key + '/in_varlev']),
                                             list(g[key + '/varlev']), times[:S], times[1] - times[0], g['lat'], g['lon'],
                                             interval, sol, impute, first)
        got = OE.estimator_series(res, steps, es, keep, keep_time, first)
        want = g[key + '/result']
        assert got.shape == want.shape, key
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg=key)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6, equal_nan=True, err_msg=key)


def test_product_estimator_host_loop_matches_reference_loop(golden_dir):
    """dlwp_b200.model.TimeSeriesEstimator (host loop; the device loop is checked against it in tests/test_estimator_gpu.py)
    vs the same golden: values, dimension names, f_hour / time / varlev coordinates."""
    from dlwp_b200.model import DLWPNeuralNet, TimeSeriesEstimator
    for g, key, t_in, t_out, interval, sol, steps, impute, first, keep_time in _estimator_cases(golden_dir):
        w = g[key + '/w']
        times = g['times'].astype('datetime64[s]')
        in_vl = [v for v in g[key + '/in_varlev'] if v != 'SOL']

        class Gen(object):
            _input_sel, _output_sel = {'varlev': in_vl}, {'varlev': list(g[key + '/varlev'])}
            _input_time_steps, _output_time_steps, _interval, _add_insolation = t_in, t_out, interval, sol
            _n_sample = g[key + '/p'].shape[0]
            sample_times, lat, lon = times[:g[key + '/p'].shape[0]], g['lat'], g['lon']
            convolution_shape = g[key + '/p'].shape[1:]

            def generate(self, samples, scale_and_impute=True):
                return g[key + '/p'].copy(), g[key + '/t'].copy()

        dlwp = DLWPNeuralNet(is_convolutional=True, time_dim=t_in, scaler_type=None, scale_targets=False)
        dlwp.predict = lambda x, **kw: np.tanh(np.einsum('nchw,co->nohw', np.asarray(x, np.float32), w)).astype(np.float32)
        est = TimeSeriesEstimator(dlwp, Gen())
        est._device_ok = lambda: False
        out = est.predict(steps, impute=impute, keep_time_dim=keep_time, prefer_first_times=first)
        want = g[key + '/result']
        got = np.asarray(out.values)
        assert got.shape == want.shape and tuple(out.dims) == tuple(g[key + '/dims']), key
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6, equal_nan=True, err_msg=key)
        coords = out.coords if isinstance(out.coords, dict) else {k: v.values for k, v in out.coords.items()}
        np.testing.assert_array_equal(np.asarray(coords['f_hour']).astype('timedelta64[s]').astype(np.int64), g[key + '/f_hour'])
        np.testing.assert_array_equal(np.asarray(coords['time']).astype('datetime64[s]').astype(np.int64), g[key + '/time'])
        assert list(coords['varlev']) == list(g[key + '/varlev'])


def test_product_estimator_functional_sequence_matches_reference(golden_dir):
    """extensions.py:204-208: a DLWPFunctional that predicts a sequence (_n_steps = 2) -- the estimator defers to the model's
    own predict_timeseries; values and coordinates vs the reference's estimator + the reference's DLWPFunctional."""
    from dlwp_b200.model import DLWPFunctional, TimeSeriesEstimator
    g = _load(golden_dir, 'estimator.npz')
    key = 'functional_sequence'
    w = g[key + '/w'].astype(np.float64)
    times = g['times'].astype('datetime64[s]')
    S = g[key + '/p'].shape[0]

    class Model(object):
        outputs = [None, None]

        def predict(self, x, **kwargs):
            a = np.tanh(np.einsum('nchw,co->nohw', np.asarray(x, np.float64), w))
            return [a.astype(np.float32), np.tanh(np.einsum('nchw,co->nohw', a, w)).astype(np.float32)]

    class Gen(object):
        _input_sel, _output_sel = {'varlev': list(g['names'])}, {'varlev': list(g['names'])}
        _input_time_steps, _output_time_steps, _interval, _add_insolation = 1, 1, 1, False
        _n_sample, sample_times, lat, lon = S, times[:S], g['lat'], g['lon']
        convolution_shape = g[key + '/p'].shape[1:]

        def generate(self, samples, scale_and_impute=True):
            return g[key + '/p'].copy(), g[key + '/t'].copy()

    fun = DLWPFunctional(is_convolutional=True, is_recurrent=False, time_dim=1)
    fun.model, fun._n_steps = Model(), 2
    out = TimeSeriesEstimator(fun, Gen()).predict(5)
    got, want = np.asarray(out.values), g[key + '/result']
    assert got.shape == want.shape and tuple(out.dims) == tuple(g[key + '/dims'])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6, equal_nan=True)
    coords = out.coords if isinstance(out.coords, dict) else {k: v.values for k, v in out.coords.items()}
    np.testing.assert_array_equal(np.asarray(coords['f_hour']).astype('timedelta64[s]').astype(np.int64), g[key + '/f_hour'])
    np.testing.assert_array_equal(np.asarray(coords['time']).astype('datetime64[s]').astype(np.int64), g[key + '/time'])


def series_generator_cases(golden_dir):
    """(golden, key, ArraySeriesGenerator, spec) for every case of tests/golden/series_generator.npz."""
    from dlwp_b200.model import ArraySeriesGenerator
    g = _load(golden_dir, 'series_generator.npz')
    for key in [str(c) for c in g['cases']]:
        t_in, t_out, seq, interval, sol, batch, n_sample, n_batches = [int(v) for v in g[key + '/spec']]
        series = ArraySeriesGenerator(g['data'], g['times'].astype('datetime64[s]'), g['lat'], g['lon'], list(g['names']),
                                      list(g[key + '/in_sel']), list(g[key + '/out_sel']), t_in, t_out, interval, bool(sol))
        yield g, key, series, (t_in, t_out, seq, interval, sol, batch, n_sample, n_batches)


def test_array_series_generator_matches_reference_generator(golden_dir):
    """The reference's own SeriesDataGenerator (generators.py:323-640) run on the xarray stand-in
    (make_golden.py:gen_series_generator) vs the in-memory ArraySeriesGenerator: predictors incl. the insolation channel
    (float32 insolation: 1e-6), first target array, shapes and sample count (sequence = None geometry)."""
    for g, key, series, (t_in, t_out, seq, interval, sol, batch, n_sample, n_batches) in series_generator_cases(golden_dir):
        assert tuple(series.convolution_shape) == tuple(g[key + '/shapes'][:3])
        assert tuple(series.output_convolution_shape) == tuple(g[key + '/shapes'][3:])
        p, t = series.generate([])
        if not seq:
            assert series._n_sample == n_sample
        np.testing.assert_allclose(p[:n_sample], g[key + '/p'], rtol=0, atol=1e-6, err_msg=key)
        np.testing.assert_array_equal(t[:n_sample], g[key + '/t0'], err_msg=key)


def test_product_series_data_generator_matches_reference_generator(golden_dir):
    """dlwp_b200.model.SeriesDataGenerator on the xarray stand-in vs the reference's own class on the same dataset
    (generators.py:323-640 -> series_generator.npz): sample count, batch count, every shape property, all samples, one batch,
    every target of a sequence; `as_array_series()` hands the same series to the GPU assembly; and the TimeSeriesEstimator
    accepts it as its generator."""
    import sys
    sys.path.insert(0, golden_dir)
    import fake_xarray
    from dlwp_b200.model import DLWPNeuralNet, SeriesDataGenerator, TimeSeriesEstimator
    g = _load(golden_dir, 'series_generator.npz')
    times = g['times'].astype('datetime64[s]').astype('datetime64[ns]')
    pred = fake_xarray.DataArray(g['data'], coords=[times, g['names'], g['lat'], g['lon']],
                                 dims=['sample', 'varlev', 'lat', 'lon'])
    ds = fake_xarray.Dataset({'sample': times, 'varlev': g['names'], 'lat': g['lat'], 'lon': g['lon']}, predictors=pred)
    model = DLWPNeuralNet(is_convolutional=True, is_recurrent=False, time_dim=1, scaler_type=None, scale_targets=False)
    for key in [str(c) for c in g['cases']]:
        t_in, t_out, seq, interval, sol, batch, n_sample, n_batches = [int(v) for v in g[key + '/spec']]
        plain = key == 'plain'
        gen = SeriesDataGenerator(model, ds, input_sel=None if plain else {'varlev': list(g[key + '/in_sel'])},
                                  output_sel=None if plain else {'varlev': list(g[key + '/out_sel'])},
                                  input_time_steps=t_in, output_time_steps=t_out, sequence=seq or None, interval=interval,
                                  add_insolation=bool(sol), batch_size=batch, shuffle=False, remove_nan=False)
        assert (gen._n_sample, len(gen)) == (n_sample, n_batches), key
        assert tuple(gen.convolution_shape) + tuple(gen.output_convolution_shape) == tuple(g[key + '/shapes']), key
        p, t = gen.generate([], scale_and_impute=False)
        xb, yb = gen[1]
        np.testing.assert_allclose(p, g[key + '/p'], rtol=0, atol=1e-6, err_msg=key)
        np.testing.assert_allclose(xb, g[key + '/xb'], rtol=0, atol=1e-6, err_msg=key)
        for k, (tt, yy) in enumerate(zip(t if seq else [t], yb if seq else [yb])):
            np.testing.assert_array_equal(tt, g[key + '/t%d' % k], err_msg=key)
            np.testing.assert_array_equal(yy, g[key + '/yb%d' % k], err_msg=key)
        arr = gen.as_array_series()
        np.testing.assert_allclose(arr.generate([])[0][:n_sample], g[key + '/p'], rtol=0, atol=1e-6, err_msg=key)
        if not seq:
            est = TimeSeriesEstimator(model, gen)
            assert list(est._output_sel['varlev']) == [str(v) for v in g[key + '/out_sel']]
            assert est._input_time_steps == t_in and len(est.generator.sample_times) == n_sample
    # recurrent models keep the time axis (generators.py:451-461)
    rec = DLWPNeuralNet(is_convolutional=True, is_recurrent=True, time_dim=2, scaler_type=None, scale_targets=False)
    gen = SeriesDataGenerator(rec, ds, input_time_steps=2, output_time_steps=2, add_insolation=True, batch_size=5)
    assert gen.convolution_shape == (2, 5, 4, 6) and gen.output_convolution_shape == (2, 4, 4, 6)
    assert gen.shape_2d == (10, 4, 6) and gen.output_shape_2d == (8, 4, 6) and gen.shape == (2, 4, 4, 6)
    assert gen.n_features == 2 * 5 * 24 and gen.dense_shape == (2, 5 * 24) and gen[0][0].shape == (5, 2, 5, 4, 6)
    with pytest.raises(ValueError):
        SeriesDataGenerator(rec, object())


def test_insolation_matches_reference_function(golden_dir):
    """DLWP/util.py:300-352 run from the reference's source text (tests/golden/make_golden.py:gen_insolation)."""
    from oracle import estimator as OE
    from dlwp_b200 import util
    g = _load(golden_dir, 'insolation.npz')
    dates = g['dates'].astype('datetime64[s]')
    for fn in (OE.insolation, util.insolation):
        np.testing.assert_array_equal(fn(dates, g['lat'], g['lon']), g['sol'])
        lon2, lat2 = np.meshgrid(g['lon'], g['lat'])
        np.testing.assert_array_equal(fn(dates[:3], lat2, lon2, S=2.), g['sol2'])
        with pytest.raises(ValueError):
            fn(dates, lat2, g['lon'])
